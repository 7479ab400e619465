#!/bin/bash
# One gpurun call: the whole GPU test suite (one process, as the round-end driver runs it), smoke, and the default
# bench with both arms.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench reference exit $?"
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_ref.log", "gpurun_out/bench.log"):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, d.get("impl", "nova"), round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 2),
              {k: round(v["value"], 1) for k, v in d.items() if isinstance(v, dict) and "value" in v and k != "roofline"})
    except Exception as e:
        print(f, "unreadable", e)
PY
