for D in 768 1024 1536; do
for c in 8192 0; do echo "D=$D CHAIN_ROWS=$c"; PROFILE_SETS_D=$D NOVA_B200_CHAIN_ROWS=$c PROFILE_SETS_N=24,32,51,56,64,72,80,102 python scripts/profile_sets.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print([(r['rows'],r['device_ms']) for r in d['per_set']], d['pass_wall_ms'])"; done; done
