#!/bin/bash
# usage (GPU box): bash scripts/gpu_prof.sh TAG KERNEL_REGEX  -> launch list + one --set full capture (after a plain run exits 0)
TAG=${1:-x}; K=${2:-ransac_score}
CMD="python bench.py --steps 2 --warmup 1 --hyp 2e7 --no-cpu --e2e-steps 1"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo plain run failed; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    k=r[4].split("(")[0][-40:]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[-1])/1e3
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-42s n=%3d total %9.1f us avg %8.1f us" % (k,a[0],a[1],a[1]/a[0]))
PY
