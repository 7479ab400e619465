#!/bin/bash
# A/B of library variants on one B200: quick parity tests, then bench.py under each env setting.
# usage: scripts/gpu_ab.sh "<workloads>" "ENV1=a ENV2=b" "ENV1=c" ...   (each later arg = one variant's env)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
WLS="$1"; shift
for f in test_gpu_gemm_2cta test_gpu_head; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --tb=short > gpurun_out/ab_$f.log 2>&1
  echo "$f exit $? : $(tail -1 gpurun_out/ab_$f.log)"
done
i=0
for envs in "$@"; do
  for wl in $WLS; do
    tag="v${i}_${wl}"
    env $envs timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/ab_$tag.log 2> gpurun_out/ab_$tag.err
    echo "variant $i [$envs] $wl exit $?"
    python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{tag}.log").read().strip().splitlines()[-1])
    print("   ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "step_frac", round(d["step_roofline"]["frac"], 3),
          "ada TF/s", round(d["roofline"]["achieved"], 1), {k: round(v["ms_per_step"], 2) for k, v in d["kernel_shares"].items()},
          "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("   unreadable", e)
PY
  done
  i=$((i+1))
done
