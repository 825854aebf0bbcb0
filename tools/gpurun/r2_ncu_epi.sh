# ncu --set full of the window-stage kernels (plain run first, as the recipe requires)
B=${1:-1024}
NO_ORACLE=1 ITERS=3 python tools/time_stage.py $B 4400 > gpurun_out/plain_epi.log 2>&1 &&
NO_ORACLE=1 ITERS=3 ncu --set full --clock-control none --import-source on -k "regex:k_corr_shapes|k_epilogue" -s 6 -c 2 -f -o gpurun_out/${2:-win_r2} python tools/time_stage.py $B 4400 > gpurun_out/ncu_epi2.log 2>&1
tail -3 gpurun_out/ncu_epi2.log
