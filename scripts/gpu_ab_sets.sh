#!/bin/bash
# same-box A/B of two library builds on the set-by-set pattern: lib/libnova_base.so vs lib/libnova_b200.so
set -u
for l in ${AB_LIBS:-base b200 base b200}; do
  echo "lib=$l"
  NOVA_B200_LIB=$PWD/nova_pointcloud_b200/lib/libnova_$l.so timeout 300 python scripts/profile_sets.py 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print([(r['rows'],r['device_ms']) for r in d['per_set']], d['pass_wall_ms'])"
done
