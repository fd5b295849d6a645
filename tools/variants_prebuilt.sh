#!/bin/bash
# time K4 variants that were BUILT IN THE BUILD CONTAINER (tools/_bin/lib_<name>.so, git-ignored but shipped with the
# snapshot) so that no GPU-box time goes into nvcc:   tools/variants_prebuilt.sh v0 v1 ...
cd "$(dirname "$0")/.."
P=headland_trajectory_planning_b200
cp $P/libheadland_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp tools/_bin/lib_$v.so $P/libheadland_b200.so || continue
  echo "=== $v"
  REPS=${REPS:-7} timeout 180 python tools/profile_k4.py 4096 2>&1 | tail -2 | cut -c1-420
done
cp /tmp/lib_orig.so $P/libheadland_b200.so
