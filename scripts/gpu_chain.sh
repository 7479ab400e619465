#!/bin/bash
# chain-kernel check: parity tests, stage timeline, then per-set timings with the chain kernel on and off (same box)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu --tb=short > gpurun_out/chain_test.log 2>&1; echo "chain tests exit $?"; tail -15 gpurun_out/chain_test.log
for r in 64 128; do
  NOVA_B200_CHAIN_CLUSTER_ROWS=$r timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu --tb=short -k bit_identical > gpurun_out/chain_test_r$r.log 2>&1; echo "chain tests rows=$r exit $?"; tail -3 gpurun_out/chain_test_r$r.log
done
timeout 300 python scripts/profile_chain_timeline.py > gpurun_out/chain_timeline.json 2> gpurun_out/chain_timeline.err; echo "timeline $?"
for c in ${CHAIN_VARIANTS:-1 0}; do
  echo "NOVA_B200_CHAIN=$c"
  NOVA_B200_CHAIN=$c timeout 300 python scripts/profile_sets.py 2> gpurun_out/chain_sets_$c.err | tail -1 > gpurun_out/chain_sets_$c.json
  python -c "
import json
d=json.loads(open('gpurun_out/chain_sets_$c.json').read())
print([(r['rows'],r['device_ms']) for r in d['per_set']], d['pass_wall_ms'])" || tail -5 gpurun_out/chain_sets_$c.err
done
