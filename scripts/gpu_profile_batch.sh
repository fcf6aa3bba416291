mkdir -p gpurun_out
CMD="python scripts/bench_batch.py 512"
timeout 600 $CMD > gpurun_out/batch_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/batch_launches.csv $CMD > gpurun_out/ncu_bl.log 2>&1
timeout 600 $CMD > gpurun_out/batch_plain2.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'hough_bin_kernel|vote_kernel|hough_rank_kernel|hough_interval_kernel' -s 8 -c 4 -o gpurun_out/prof_batch $CMD > gpurun_out/ncu_bf.log 2>&1
tail -2 gpurun_out/batch_plain.log; tail -3 gpurun_out/ncu_bf.log
