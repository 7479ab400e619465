#!/bin/bash
# One gpurun call: every GPU test file in its own process (a trapped kernel poisons only its own
# CUDA context), then smoke and a short bench.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in test_gpu_gemm test_gpu_gemm_2cta test_gpu_chamfer test_gpu_head test_gpu_pipeline; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --tb=short -s > gpurun_out/$f.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/$f.log
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
NOVA_B200_STREAMS=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1stream.log 2> gpurun_out/bench_1stream.err; echo "bench 1stream exit $?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload cfg3-2048 > gpurun_out/bench_cfg3_2048.log 2> gpurun_out/bench_cfg3_2048.err; echo "bench cfg3-2048 exit $?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
for f in ("gpurun_out/bench.log", "gpurun_out/bench_1stream.log", "gpurun_out/bench_cfg3_2048.log"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 2), "step_frac", round(d["step_roofline"]["frac"], 3),
              "dominant TF/s", round(d["roofline"]["achieved"], 1), {k: round(v["ms_per_step"], 2) for k, v in d["kernel_shares"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
