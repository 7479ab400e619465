#!/bin/bash
# One gpurun call: GPU test suite of the current build, then same-box A/B at cfg2 of (new, base = previous commit's
# library, new with NOVA_B200_FIXED_N=0), then the Chamfer kernel variants.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/ -x -q -m gpu --tb=short > gpurun_out/ab_test.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/ab_test.log
timeout 300 python scripts/probe_chamfer_packed.py > gpurun_out/probe_chamfer_packed.json 2> gpurun_out/probe_chamfer_packed.err; echo "chamfer probe exit $?"; cat gpurun_out/probe_chamfer_packed.json
BASE="$PWD/nova_pointcloud_b200/lib/libnova_b200_base.so"
for wl in "${@:-cfg2}"; do
for t in new base nofix new base nofix; do
  unset NOVA_B200_LIB NOVA_B200_FIXED_N
  if [ "$t" = base ]; then export NOVA_B200_LIB="$BASE"; fi
  if [ "$t" = nofix ]; then export NOVA_B200_FIXED_N=0; fi
  timeout 600 python bench.py --workload "$wl" --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-north-star --no-compile-bar > gpurun_out/ab_${wl}_$t.json 2> gpurun_out/ab_${wl}_$t.err
  echo "$wl $t exit $?"
  python - "$wl" "$t" <<'PY'
import json, sys
wl, t = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/ab_{wl}_{t}.json").read().strip().splitlines()[-1])
    print("   ms/step", round(d["ms_per_step"], 2), "clouds/s", round(d["value"], 1), "step_frac", round(d["step_roofline"]["frac"], 3),
          {k: round(v["ms_per_step"], 2) for k, v in d["kernel_shares"].items()}, "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("   unreadable", e); print(open(f"gpurun_out/ab_{wl}_{t}.err").read()[-800:])
PY
done
done
