#!/bin/bash
# Final single-GPU evidence of a tree: bench lines (all legs, per-rank share, c2, c5), launch list, ncu of the K3 stage kernels.
# usage: bash scripts/gpu_final.sh TAG
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo bench rc=$?
timeout 600 python bench.py --hyp 1.25e7 --steps 8 --no-cpu --spectra 0 > gpurun_out/bench_${TAG}_1.25e7.json 2> /dev/null
timeout 900 python bench.py --config c2 > gpurun_out/bench_${TAG}_c2.json 2> /dev/null; echo c2 rc=$?
timeout 900 python bench.py --config c5 > gpurun_out/bench_${TAG}_c5.json 2> /dev/null; echo c5 rc=$?
python scripts/prof_step_host.py 1.25e7 > gpurun_out/step_host_$TAG.log 2>&1; cat gpurun_out/step_host_$TAG.log
CMD="python bench.py --steps 2 --warmup 1 --hyp 1e8 --no-cpu --e2e-steps 1 --spectra 64"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'ransac_(draw_fit|monotonic|match|cost|exact|record)_kernel' -s 12 -c 6 -f -o gpurun_out/prof_k3_$TAG $CMD > gpurun_out/ncu_full_k3_$TAG.log 2>&1
python - <<PY
import json
for n in ("n1", "1.25e7", "c2", "c5"):
    try:
        d=json.loads(open("gpurun_out/bench_${TAG}_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.4g e2e %.4g ms/step %.3f kernel_ms %.3f frac %.4f e2e_ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"]), d.get("sequential", {}).get("ms_per_step"))
    except Exception as e:
        print(n, "unreadable", e)
PY
