mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --hyp 8e6 --no-cpu --e2e-steps 1"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'ransac_score_kernel|ransac_draw_fit_kernel|ransac_monotonic_kernel' -s 3 -c 3 -o gpurun_out/prof_k3 $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
