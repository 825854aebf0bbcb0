BFL_NO_WSEP=1 ncu --set full --clock-control none --import-source on -k regex:k_window_weighted -c 1 -s 2 -o gpurun_out/weighted_r1 -f python tools/time_weighted.py 512 > gpurun_out/ncu_weighted.log 2>&1
tail -2 gpurun_out/ncu_weighted.log
