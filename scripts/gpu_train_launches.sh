set -u
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_launches.csv python scripts/profile_train.py 768 65536 > gpurun_out/ncu_train.log 2>&1; echo exit $?
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open("gpurun_out/r2_train_launches.csv")) if len(r)>5]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
data=[(r[ik], float(r[iv].replace(",",""))) for r in rows[1:] if r[iv].replace(",","").replace(".","").isdigit()]
n=len(data)//5  # 5 identical steps (2 warm-up + 3 timed)
last=data[-n:]
agg=collections.defaultdict(lambda:[0,0.0])
for k,v in last:
    k=re.sub(r"\(.*","",k); k=k.replace("void ","").replace("nova::","")
    agg[k][0]+=1; agg[k][1]+=v
tot=sum(v for _,v in last)
print("launches per step", n, "sum us", round(tot/1e3,1))
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:22]:
    print(f"{t/1e3:9.1f} us  {c:4d}  {k[:110]}")
PY
