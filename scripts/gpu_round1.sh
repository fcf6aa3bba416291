mkdir -p gpurun_out; timeout 600 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/pytest_gpu_3.log 2>&1; echo exit=$? >> gpurun_out/pytest_gpu_3.log
timeout 900 python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo bench_exit=$?
CMD="python bench.py --steps 2 --warmup 1 --hyp 2e7 --no-cpu --e2e-steps 1"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ransac_kernel -s 1 -c 1 -o gpurun_out/prof_ransac $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu_3.log; cat gpurun_out/bench_a.json | cut -c1-600
