#!/bin/bash
# Launch list (ncu gpu__time_duration) of a short K3-only bench run, aggregated per kernel.  usage: bash scripts/gpu_launches.sh TAG [ENV=VAL ...]
TAG=${1:-ll}; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --hyp ${HYP:-3.4e7} --no-cpu --e2e-steps 1 --spectra 0"
env "$@" RB2_X=0 timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 && env "$@" RB2_X=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    k=r[4].split("(")[0][-44:]; a=agg.setdefault(k,[0,0.0,0.0]); a[0]+=1; a[1]+=float(r[-1])/1e3; a[2]=max(a[2],float(r[-1])/1e3)
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]: print("%-46s n=%3d total %9.1f us avg %8.1f us max %8.1f" % (k,a[0],a[1],a[1]/a[0],a[2]))
PY
