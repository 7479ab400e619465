#!/bin/bash
# next-tile L2 prefetch in the AdaLN / tail epilogues: parity tests, K-sweep probe vs cuBLAS, then same-box A/B bench
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_gemm.py tests/test_gpu_gemm_2cta.py -x -q -m gpu --tb=short > gpurun_out/pf_test.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/pf_test.log
timeout 600 python scripts/probe_gemm_k.py > gpurun_out/probe_gemm_k.log 2>&1; echo "probe exit $?"; cat gpurun_out/probe_gemm_k.log | tail -10
for t in 1 0 1 0; do
  NOVA_B200_EPI_PREFETCH=$t timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-north-star --no-compile-bar > gpurun_out/pf_$t.json 2> gpurun_out/pf_$t.err
  echo "EPI_PREFETCH=$t exit $?"
  python - "$t" <<'PY'
import json, sys
t = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/pf_{t}.json").read().strip().splitlines()[-1])
    print("   ms/step", round(d["ms_per_step"], 2), "clouds/s", round(d["value"], 1), "step_frac", round(d["step_roofline"]["frac"], 3),
          {k: round(v["ms_per_step"], 2) for k, v in d["kernel_shares"].items()}, "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("   unreadable", e); print(open(f"gpurun_out/pf_{t}.err").read()[-800:])
PY
done
