# round-2 evidence: ncu --set full of the window-stage kernels (1,024 and 128 curves) and the launch list of bench.py
NO_ORACLE=1 ITERS=3 python tools/time_stage.py 1024 4400 > gpurun_out/plain_ev.log 2>&1 &&
NO_ORACLE=1 ITERS=3 ncu --set full --clock-control none --import-source on -k "regex:k_corr_shapes|k_epilogue2" -s 6 -c 2 -f -o gpurun_out/window_r2 python tools/time_stage.py 1024 4400 > gpurun_out/ncu_ev1.log 2>&1
NO_ORACLE=1 ITERS=3 ncu --set full --clock-control none --import-source on -k "regex:k_corr_shapes|k_epilogue2" -s 6 -c 2 -f -o gpurun_out/window_r2_b128 python tools/time_stage.py 128 4400 > gpurun_out/ncu_ev2.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sustained > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sustained > gpurun_out/ncu_ev3.log 2>&1
tail -2 gpurun_out/ncu_ev1.log gpurun_out/ncu_ev2.log gpurun_out/ncu_ev3.log | cut -c1-200
wc -l gpurun_out/launches_r2.csv
