#!/bin/bash
# GPU box: all GPU tests, then a short bench per named knob setting, then a launch list.
# usage: bash scripts/gpu_sweep.sh TAG "name1:ENV=1 ENV2=2" "name2:..."
TAG=${1:-sweep}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_exit=$?
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_$TAG.log | tail -40
grep "COEFF_GAP" gpurun_out/pytest_$TAG.log
B="python bench.py --steps 5 --warmup 3 --no-cpu --spectra 0"
for spec in "default:RB2_X=0" "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  env $envs timeout 600 $B > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo $name rc=$?
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s value %.4g  e2e %.4g  ms/step %.3f  kernel_ms %.3f frac %.4f e2e_ms %.3f" % (f, d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
python scripts/prof_default_fit.py > gpurun_out/default_fit_$TAG.log 2>&1; cat gpurun_out/default_fit_$TAG.log
CMD="python bench.py --steps 2 --warmup 1 --hyp 2e7 --no-cpu --e2e-steps 1 --spectra 0"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    k=r[4].split("(")[0][-44:]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[-1])/1e3
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-46s n=%3d total %9.1f us avg %8.1f us" % (k,a[0],a[1],a[1]/a[0]))
PY
