#!/bin/bash
# geometry tests + timings + the default bench (no CPU baseline): prints the set-by-set legs
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_geometry.py -x -q 2>&1 | tail -3
timeout 120 python scripts/profile_geometry.py 2>&1 | tail -1 | tee gpurun_out/plain_geometry.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_nocpu.log 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_nocpu.log").read().strip().splitlines()[-1])
print(d["value"], d["set_by_set"]["value"], d["set_by_set_20"], d["geometry"]["softmax_interp_quarter_ms"])
PY
