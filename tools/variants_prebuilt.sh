#!/bin/bash
# time variants that were BUILT IN THE BUILD CONTAINER (tools/_bin/lib_<name>.so, git-ignored but shipped with the
# snapshot) so that no GPU-box time goes into nvcc:   [TOOL="tools/profile_df.py"] tools/variants_prebuilt.sh v0 v1 ...
# TOOL defaults to the K4 sweep timer; T=1 also runs the named pytest file with each variant in place.
cd "$(dirname "$0")/.."
P=headland_trajectory_planning_b200
TOOL=${TOOL:-"tools/profile_k4.py 4096"}
cp $P/libheadland_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp tools/_bin/lib_$v.so $P/libheadland_b200.so || continue
  echo "=== $v"
  for r in $(seq ${RUNS:-1}); do REPS=${REPS:-7} timeout 180 python $TOOL 2>&1 | tail -${TAIL:-2} | cut -c1-420; done
  if [ -n "$NCU_SUM" ]; then   # summed duration of the kernels matching $NCU_SUM over one run of the tool (host-independent)
    REPS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$NCU_SUM --csv --log-file /tmp/ncu_$v.csv python $TOOL > /dev/null 2>&1
    python - /tmp/ncu_$v.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]; vi, ui = h.index("Metric Value"), h.index("Metric Unit")
t = sum(float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui], 1e-6) for r in rows[1:])
print(f"ncu: {len(rows) - 1} launches, {t:.3f} ms summed = {t / 3:.3f} ms per field (3 fields)")
PY
  fi
  if [ -n "$TESTS" ]; then timeout 300 python -m pytest $TESTS -m gpu -x -q 2>&1 | tail -2; fi
done
cp /tmp/lib_orig.so $P/libheadland_b200.so
