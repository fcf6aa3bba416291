#!/bin/bash
# Round-2 phase B: GPU tests of the changed kernels + knob sweep of the K3 stage grids.  usage: bash scripts/gpu_r2b.sh TAG
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_exit=$?
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_$TAG.log | tail -40
grep "COEFF_GAP" gpurun_out/pytest_$TAG.log
B="python bench.py --steps 5 --warmup 3 --no-cpu --spectra 0"
run() { name=$1; shift; env "$@" timeout 600 $B > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo $name rc=$?; }
run default RB2_X=0
for a in 2 3 4 6; do run A$a RB2_A_CTAS_PER_SM=$a; done
run A3_small RB2_A_CTAS_PER_SM=3 RB2_B_CTAS_PER_SM=2 RB2_C1_CTAS_PER_SM=1 RB2_C2_CTAS_PER_SM=2
run A4_small RB2_A_CTAS_PER_SM=4 RB2_B_CTAS_PER_SM=2 RB2_C1_CTAS_PER_SM=1 RB2_C2_CTAS_PER_SM=2
run A8_small RB2_B_CTAS_PER_SM=2 RB2_C1_CTAS_PER_SM=1 RB2_C2_CTAS_PER_SM=2
run A3_mid RB2_A_CTAS_PER_SM=3 RB2_B_CTAS_PER_SM=4 RB2_C1_CTAS_PER_SM=2 RB2_C2_CTAS_PER_SM=4
run lanes1 RB2_PIPE_LANES=1
run lanes2_A3s RB2_PIPE_LANES=2 RB2_A_CTAS_PER_SM=3 RB2_B_CTAS_PER_SM=2 RB2_C1_CTAS_PER_SM=1 RB2_C2_CTAS_PER_SM=2
run lanes4_A3s RB2_PIPE_LANES=4 RB2_A_CTAS_PER_SM=3 RB2_B_CTAS_PER_SM=2 RB2_C1_CTAS_PER_SM=1 RB2_C2_CTAS_PER_SM=2
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s value %.4g  e2e %.4g  ms/step %.3f  kernel_ms %.3f frac %.4f" % (f, d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
CMD="python bench.py --steps 2 --warmup 1 --hyp 2e7 --no-cpu --e2e-steps 1 --spectra 0"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'ransac_(draw_fit|monotonic|match|cost|exact|record)_kernel|refit_kernel' -s 12 -c 8 -f -o gpurun_out/prof_k3_$TAG $CMD > gpurun_out/ncu_full_k3_$TAG.log 2>&1
tail -n 1 gpurun_out/ncu_full_k3_$TAG.log
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    k=r[4].split("(")[0][-44:]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[-1])/1e3
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-46s n=%3d total %9.1f us avg %8.1f us" % (k,a[0],a[1],a[1]/a[0]))
PY
