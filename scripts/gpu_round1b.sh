# full evidence run: tests, bench, launch list, ncu full of the K3 stage kernels
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu_final.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu_final.log; tail -2 gpurun_out/pytest_gpu_final.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench_rc=$?
CMD="python bench.py --steps 2 --warmup 1 --hyp 2e7 --no-cpu --e2e-steps 1"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launch.log 2>&1
CMD2="python bench.py --steps 1 --warmup 1 --hyp 8e6 --no-cpu --e2e-steps 1"
timeout 600 $CMD2 > gpurun_out/plain2.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'ransac_score_kernel|ransac_draw_fit_kernel|ransac_monotonic_kernel' -s 3 -c 3 -o gpurun_out/prof_k3_final $CMD2 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
