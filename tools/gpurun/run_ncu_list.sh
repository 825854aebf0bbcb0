ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log | cut -c1-200; wc -l gpurun_out/launches_r1.csv
