#!/bin/bash
# same-box comparison of library builds nova_pointcloud_b200/lib/libnova_<name>.so on the set-by-set pattern
# (per-set device time of a fused 25-step call and the 64-set pass).  Usage: gpu_ab_variants.sh name [name ...]
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for rep in 1 2; do
for v in "$@"; do
  export NOVA_B200_LIB="$PWD/nova_pointcloud_b200/lib/libnova_$v.so"
  PROFILE_SETS_D=${PROFILE_SETS_D:-768} PROFILE_SETS_N=${PROFILE_SETS_N:-1,8,16,28,32,40,51} timeout 300 python scripts/profile_sets.py 2> gpurun_out/var_$v.err | tail -1 > gpurun_out/var_${v}_$rep.json
  python - "$v" "$rep" <<'PY'
import json, sys
v, rep = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/var_{v}_{rep}.json").read())
    print(v, rep, [(r["rows"], r["device_ms"]) for r in d["per_set"]], "pass", d["pass_wall_ms"])
except Exception as e:
    print(v, "unreadable", e); print(open(f"gpurun_out/var_{v}.err").read()[-600:])
PY
done
done
