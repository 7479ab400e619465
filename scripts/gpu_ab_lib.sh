#!/bin/bash
# same-box A/B of two builds of the library: nova_pointcloud_b200/lib/libnova_b200_base.so (built from the previous
# commit) against the current one, after the parity tests of the current build.  Usage: gpu_ab_lib.sh [workload ...]
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_gemm.py tests/test_gpu_gemm_2cta.py tests/test_gpu_pipeline.py -x -q -m gpu --tb=short > gpurun_out/ab_test.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/ab_test.log
BASE="$PWD/nova_pointcloud_b200/lib/libnova_b200_base.so"
for wl in "${@:-cfg2}"; do
for t in new base new base; do
  if [ "$t" = base ]; then export NOVA_B200_LIB="$BASE"; else unset NOVA_B200_LIB; fi
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
