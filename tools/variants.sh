#!/bin/bash
# build K4 variants on the GPU box and time each: usage tools/variants.sh "<flags1>" "<flags2>" ...
for f in "$@"; do
  HL_NVCC_FLAGS="$f" python headland_trajectory_planning_b200/build_ext.py --force > /dev/null 2>&1 || { echo "build failed: $f"; continue; }
  echo "=== $f"
  REPS=4 timeout 180 python tools/profile_k4.py 4096 2>&1 | tail -2 | cut -c1-420
done
HL_NVCC_FLAGS="" python headland_trajectory_planning_b200/build_ext.py --force > /dev/null 2>&1
