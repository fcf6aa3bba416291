#!/bin/bash
# Round-2 evidence on the GPU box: tests, bench lines, launch list, ncu captures.  usage: bash scripts/gpu_r2r.sh TAG [notest]
TAG=${1:-r2r}
mkdir -p gpurun_out
if [ "$2" != "notest" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_exit=$?
  grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_$TAG.log | tail -20
fi
timeout 900 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo bench rc=$?
timeout 600 python bench.py --hyp 1.25e7 --steps 8 --no-cpu --spectra 0 > gpurun_out/bench_${TAG}_1.25e7.json 2> gpurun_out/bench_${TAG}_1.25e7.err
timeout 900 python bench.py --config c2 > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo c2 rc=$?
timeout 900 python bench.py --config c5 > gpurun_out/bench_${TAG}_c5.json 2> gpurun_out/bench_${TAG}_c5.err; echo c5 rc=$?
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference_arm.json 2> gpurun_out/bench_${TAG}_reference_arm.err; echo ref rc=$?
python scripts/prof_step_host.py 1.25e7 > gpurun_out/step_host_$TAG.log 2>&1; cat gpurun_out/step_host_$TAG.log
python scripts/prof_k4.py > gpurun_out/k4_$TAG.log 2>&1; cat gpurun_out/k4_$TAG.log
python scripts/prof_e2e.py 1.25e7 > gpurun_out/e2e_prof_$TAG.log 2>&1; head -8 gpurun_out/e2e_prof_$TAG.log
CMD="python bench.py --steps 2 --warmup 1 --hyp 1e8 --no-cpu --e2e-steps 1 --spectra 64"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'ransac_(draw_fit|monotonic|match|cost|exact|record)_kernel' -s 12 -c 6 -f -o gpurun_out/prof_k3_$TAG $CMD > gpurun_out/ncu_full_k3_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'refit_kernel|hough_rank|ransac_prepare_kernel|hough_bin_kernel|vote_kernel|hough_interval' -c 9 -f -o gpurun_out/prof_fixed_$TAG python scripts/prof_e2e.py 1e6 > gpurun_out/ncu_full_fixed_$TAG.log 2>&1
python - <<PY
import csv, json
rows=[r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    k=r[4].split("(")[0][-44:]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[-1])/1e3
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-46s n=%3d total %9.1f us avg %8.1f us" % (k,a[0],a[1],a[1]/a[0]))
for n in ("n1", "1.25e7", "c2", "c5"):
    try:
        d=json.loads(open("gpurun_out/bench_${TAG}_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.4g e2e %.4g ms/step %.3f kernel_ms %.3f frac %.4f e2e_ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"]))
    except Exception as e:
        print(n, "unreadable", e)
PY
