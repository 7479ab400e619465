#!/bin/bash
# fused block tail (gate GEMM epilogue) vs the resid kernel: parity tests, then bench on both (same box)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_gemm.py tests/test_gpu_gemm_2cta.py -x -q -m gpu --tb=short > gpurun_out/tail_test.log 2>&1; echo "tests exit $?"; tail -12 gpurun_out/tail_test.log
for t in 1 0; do
  NOVA_B200_FUSE_TAIL=$t timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-north-star --no-compile-bar > gpurun_out/tail_$t.json 2> gpurun_out/tail_$t.err
  echo "FUSE_TAIL=$t exit $?"
  python - "$t" <<'PY'
import json, sys
t = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/tail_{t}.json").read().strip().splitlines()[-1])
    print("   ms/step", round(d["ms_per_step"], 2), "clouds/s", round(d["value"], 1), "step_frac", round(d["step_roofline"]["frac"], 3),
          {k: round(v["ms_per_step"], 2) for k, v in d["kernel_shares"].items()}, "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("   unreadable", e); print(open(f"gpurun_out/tail_{t}.err").read()[-800:])
PY
done
