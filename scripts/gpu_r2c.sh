#!/bin/bash
# Round-2 phase C on the GPU box.  usage: bash scripts/gpu_r2c.sh TAG
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_exit=$?
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_$TAG.log | tail -40
grep "COEFF_GAP" gpurun_out/pytest_$TAG.log
python scripts/prof_k4.py > gpurun_out/k4_$TAG.log 2>&1; cat gpurun_out/k4_$TAG.log
RB2_LIB=$PWD/rascal_b200/_variants/librascal_b200_k4prof.so python scripts/prof_k4.py c3_deg4 2>&1 | grep -E "K4 profile|refit_kernel" | sort | uniq -c | sort -rn | head -8
B="python bench.py --steps 5 --warmup 3 --no-cpu --spectra 0"
run() { name=$1; shift; env "$@" timeout 600 $B > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo $name rc=$?; }
run default RB2_X=0
run A4 RB2_A_CTAS_PER_SM=4
run A3_lanes4 RB2_PIPE_LANES=4
run A2_lanes4 RB2_PIPE_LANES=4 RB2_A_CTAS_PER_SM=2
run A3_c1x2 RB2_C1_CTAS_PER_SM=2
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s value %.4g  e2e %.4g  ms/step %.3f  kernel_ms %.3f frac %.4f e2e_ms %.3f" % (f, d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
python scripts/prof_e2e.py 1e7 > gpurun_out/e2e_prof_$TAG.log 2>&1; head -45 gpurun_out/e2e_prof_$TAG.log
python scripts/prof_default_fit.py > gpurun_out/default_fit_$TAG.log 2>&1; cat gpurun_out/default_fit_$TAG.log
CMD="python bench.py --steps 2 --warmup 1 --hyp 2e7 --no-cpu --e2e-steps 1 --spectra 0"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'ransac_(draw_fit|match)_kernel' -s 6 -c 2 -f -o gpurun_out/prof_k3_$TAG $CMD > gpurun_out/ncu_full_k3_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'refit_kernel|hough_rank_kernel|ransac_prepare_kernel' -c 3 -f -o gpurun_out/prof_fixed_$TAG python scripts/prof_e2e.py 1e6 > gpurun_out/ncu_full_fixed_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    k=r[4].split("(")[0][-44:]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[-1])/1e3
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-46s n=%3d total %9.1f us avg %8.1f us" % (k,a[0],a[1],a[1]/a[0]))
PY
