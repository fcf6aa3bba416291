#!/bin/bash
# Full evidence run on the GPU box: tests, bench (with cpu_baseline), launch list, ncu --set full of the
# K3 stage kernels and of the K1/K2/K4/K5 kernels.  usage: bash scripts/gpu_evidence.sh TAG
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1; echo pytest_exit=$?; tail -2 gpurun_out/pytest_$TAG.log
timeout 1200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo bench_rc=$?
CMD="python bench.py --steps 2 --warmup 1 --hyp 2e7 --no-cpu --e2e-steps 1 --spectra 64"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
# the five K3 stage kernels of one steady-state chunk (skip the first chunk's launches)
timeout 600 $CMD > gpurun_out/plain2_$TAG.log 2>&1 && timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'ransac_(draw_fit|monotonic|match|cost|record)_kernel' -s 10 -c 5 -f -o gpurun_out/prof_k3_$TAG $CMD > gpurun_out/ncu_full_k3_$TAG.log 2>&1
# K1 / K2 / K4 / K5 kernels of the many-spectra leg (64 spectra per launch)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'hough_interval_kernel|hough_bin_kernel|hough_rank_kernel|vote_kernel|refit_kernel|match_peaks_kernel' -s 12 -c 6 -f -o gpurun_out/prof_k1245_$TAG $CMD > gpurun_out/ncu_full_k1245_$TAG.log 2>&1
# K6 (peak finding / refinement) on 512 rendered arcs: launch list and full capture
K6="python scripts/prof_peaks.py 512"
timeout 300 $K6 > gpurun_out/k6_plain_$TAG.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_k6_$TAG.csv $K6 > gpurun_out/ncu_launch_k6_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'find_peaks_kernel|refine_(windows|queue|fit|collect)_kernel' -c 5 -f -o gpurun_out/prof_k6_$TAG $K6 > gpurun_out/ncu_full_k6_$TAG.log 2>&1
tail -n 1 gpurun_out/ncu_full_k3_$TAG.log; tail -n 1 gpurun_out/ncu_full_k1245_$TAG.log
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("value %.4g  e2e %.4g  ms/step %.3f  kernel_ms %.3f frac %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"]))
print(d.get("spectra_per_sec")); print(d.get("cpu_baseline")); print(d["clocks"])
PY
