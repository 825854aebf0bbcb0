ncu --set full --clock-control none --import-source on -k "regex:k_corr_shapes|k_epilogue" -s 6 -c 2 -o gpurun_out/window_r1 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-150
