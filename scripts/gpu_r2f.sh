#!/bin/bash
# Final evidence of the round-2 tree in ONE short call, most important first (a call cut short keeps what is written).
# usage: bash scripts/gpu_r2f.sh TAG [VARIANT_SO]
TAG=${1:-r2f}
V=$2
mkdir -p gpurun_out
T0=$SECONDS
timeout 420 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_exit=$? t=$((SECONDS-T0))
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_$TAG.log | tail -8
timeout 300 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo bench rc=$? t=$((SECONDS-T0))
CMD="python bench.py --steps 2 --warmup 1 --hyp 1e8 --no-cpu --e2e-steps 1 --spectra 64"
timeout 200 $CMD > gpurun_out/plain_$TAG.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo launches t=$((SECONDS-T0))
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'ransac_(draw_fit|monotonic|match|cost|exact|record)_kernel' -s 12 -c 6 -f -o gpurun_out/prof_k3_$TAG $CMD > gpurun_out/ncu_full_k3_$TAG.log 2>&1
echo ncu_full t=$((SECONDS-T0))
timeout 120 python bench.py --hyp 1.25e7 --steps 8 --no-cpu --spectra 0 > gpurun_out/bench_${TAG}_1.25e7.json 2> /dev/null
if [ -n "$V" ]; then bash scripts/gpu_cmp.sh $V 2 > gpurun_out/cmp_$TAG.log 2>&1; cat gpurun_out/cmp_$TAG.log; fi
timeout 200 python bench.py --config c2 > gpurun_out/bench_${TAG}_c2.json 2> /dev/null; echo c2 rc=$? t=$((SECONDS-T0))
timeout 200 python bench.py --config c5 > gpurun_out/bench_${TAG}_c5.json 2> /dev/null; echo c5 rc=$? t=$((SECONDS-T0))
python - <<PY
import csv, json
try:
    rows=[r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r)>10 and r[0].isdigit()]
    agg={}
    for r in rows:
        k=r[4].split("(")[0][-44:]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[-1])/1e3
    for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:16]: print("%-46s n=%3d total %9.1f us avg %8.1f us" % (k,a[0],a[1],a[1]/a[0]))
except Exception as e:
    print("launch list unreadable", e)
for n in ("n1", "1.25e7", "c2", "c5"):
    try:
        d=json.loads(open("gpurun_out/bench_${TAG}_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.4g e2e %.4g ms/step %.3f kernel_ms %.3f frac %.4f e2e_ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"]), d.get("sequential", {}).get("ms_per_step"))
    except Exception as e:
        print(n, "unreadable", e)
PY
echo total t=$((SECONDS-T0))
