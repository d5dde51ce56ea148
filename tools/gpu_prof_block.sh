#!/bin/bash
mkdir -p gpurun_out
AMGXB_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_block.csv python tools/bench_block_dist.py dDFI 128 80 AMG > gpurun_out/prof_block.log 2>&1
tail -2 gpurun_out/prof_block.log
python - <<'PY'
import csv, collections
rows=[]
with open("gpurun_out/launches_block.csv") as f:
    lines=[l for l in f if not l.startswith("==")]
r=csv.DictReader(lines)
tot=collections.defaultdict(lambda:[0,0.0])
for row in r:
    try:
        name=row["Kernel Name"]; v=float(row["Metric Value"].replace(",",""))
    except Exception: continue
    unit=row.get("Metric Unit","")
    if unit in ("ns","nsecond"): v/=1000.0
    elif unit in ("ms","msecond"): v*=1000.0
    tot[name][0]+=1; tot[name][1]+=v
allt=sum(v[1] for v in tot.values())
print("total us", allt)
for k,v in sorted(tot.items(), key=lambda kv:-kv[1][1])[:14]:
    print(f"{v[1]:12.1f} us {100*v[1]/allt:5.1f}% n={v[0]:6d} avg={v[1]/v[0]:8.2f} {k[:110]}")
PY
