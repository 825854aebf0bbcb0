ncu --set full --clock-control none --import-source on -k regex:k_window_wsep -c 1 -s 2 -o gpurun_out/wsep_r1b -f python tools/time_weighted.py 512 > gpurun_out/ncu_wsep.log 2>&1
tail -2 gpurun_out/ncu_wsep.log
