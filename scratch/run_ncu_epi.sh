ncu --set full --clock-control none --import-source on -k regex:k_epilogue -c 1 -s 3 -o gpurun_out/epi_r1c -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_epi.log 2>&1
tail -2 gpurun_out/ncu_epi.log
