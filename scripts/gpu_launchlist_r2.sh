#!/bin/bash
# round 2: ncu launch list (gpu__time_duration) of one sampling pass of the bench command, after a plain run that exited 0
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
export NOVA_B200_GRAPH=0   # individual launches, not graph replays
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-north-star --no-compile-bar > gpurun_out/plain_bench.log 2>&1 || { echo plain run failed; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 715 -c 720 --csv --log-file gpurun_out/r2_launches_cfg2.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-north-star --no-compile-bar > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit $?"
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open("gpurun_out/r2_launches_cfg2.csv")) if len(r)>5]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
agg=collections.defaultdict(lambda:[0,0.0]); tot=0.0
for r in rows[1:]:
    try: v=float(r[iv].replace(",",""))
    except ValueError: continue
    k=re.sub(r"\(CUtensorMap.*","",r[ik]).replace("void ","").replace("nova::","")
    agg[k[:80]][0]+=1; agg[k[:80]][1]+=v; tot+=v
print("launches", sum(c for c,_ in agg.values()), "total ms", round(tot/1e6,2))
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1]):
    print(f"{t/1e3/c:8.1f} us avg  {c:4d}  share {t/tot:.3f}  {k}")
PY
