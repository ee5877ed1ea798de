#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/r2_train_tf32_launches.csv python bench.py --workload train --train-matmul tf32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_train_ncu.log 2>&1; echo rc=$?
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2_train_tf32_launches.csv')))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
hdr=rows[hi]; kn=hdr.index("Kernel Name"); mv=hdr.index("Metric Value"); mu=hdr.index("Metric Unit")
agg=collections.OrderedDict(); n=0
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    v=float(r[mv].replace(",","")); u=r[mu]
    us = v/1e3 if u.startswith("n") else (v if u.startswith("u") else v*1e3)
    k=r[kn][:80]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=us; n+=1
tot=sum(a[1] for a in agg.values()); print("launches",n,"total us",tot)
for k,(c,us) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]: print("%-82s %4d %10.1f us %5.1f%%"%(k,c,us,100*us/tot))
PY
