python tools/time_weighted.py 512 > gpurun_out/plain_wsep.log 2>&1 && cat gpurun_out/plain_wsep.log | tail -2 &&
ncu --set full --clock-control none --import-source on -k regex:k_window_wsep -c 1 -s 2 -f -o gpurun_out/wsep_r2 python tools/time_weighted.py 512 > gpurun_out/ncu_wsep.log 2>&1
tail -2 gpurun_out/ncu_wsep.log
