#!/bin/bash
# build K6 variants on the GPU box (only hl_grid.cu is recompiled) and time each: tools/variants_df.sh "<flags1>" ...
cd "$(dirname "$0")/.."
P=headland_trajectory_planning_b200
for f in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --fmad=true $f -c $P/csrc/hl_grid.cu -o $P/csrc/hl_grid.o > /dev/null 2>&1 || { echo "build failed $f"; continue; }
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $P/libheadland_b200.so $P/csrc/*.o -lcudart
  echo "=== $f"; python tools/profile_df.py 2>&1 | tail -1; python tools/profile_df.py 2>&1 | tail -1
done
