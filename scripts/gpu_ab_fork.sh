set -u
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q -m gpu --tb=short 2>&1 | tail -3
for fr in 640 1024; do
  echo "FORK_ADA_ROWS=$fr"
  NOVA_B200_FORK_ADA_ROWS=$fr timeout 300 python scripts/profile_sets.py 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print([(r['rows'],r['device_ms']) for r in d['per_set']], d['pass_wall_ms'])"
done
