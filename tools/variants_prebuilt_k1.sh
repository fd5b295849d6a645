#!/bin/bash
# time K1 variants built in the build container (tools/_bin/lib_<name>.so):   tools/variants_prebuilt_k1.sh k0 k1 ...
cd "$(dirname "$0")/.."
P=headland_trajectory_planning_b200
cp $P/libheadland_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp tools/_bin/lib_$v.so $P/libheadland_b200.so || continue
  echo "=== $v"
  python tools/profile_k1.py 16777216 random 2>&1 | tail -1 | cut -c55-140
  python tools/profile_k1.py 8388608 paths 2>&1 | tail -1 | cut -c55-140
done
cp /tmp/lib_orig.so $P/libheadland_b200.so
