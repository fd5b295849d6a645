#!/bin/bash
# build K1 variants ON THE GPU BOX and time each: usage tools/variants_k1.sh "<nvcc flags 1>" "<nvcc flags 2>" ...
for f in "$@"; do
  HL_NVCC_FLAGS="$f" python headland_trajectory_planning_b200/build_ext.py --force > /dev/null 2>&1 || { echo "build failed: $f"; continue; }
  echo "=== $f"
  python tools/profile_k1.py 8388608 paths 2>&1 | tail -1 | cut -c55-130
  python tools/profile_k1.py 16777216 random 2>&1 | tail -1 | cut -c55-130
done
HL_NVCC_FLAGS="" python headland_trajectory_planning_b200/build_ext.py --force > /dev/null 2>&1
