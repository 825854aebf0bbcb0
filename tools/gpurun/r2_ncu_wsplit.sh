# ncu of the split weighted form: --set full of one k_wsolve launch (256 curves -> sub-batches of 128)

timeout 300 python tools/time_wsplit.py 256 > gpurun_out/plain_wsplit.log 2>&1 && tail -1 gpurun_out/plain_wsplit.log &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wsolve" -c 1 -s 2 -f -o gpurun_out/wsolve_r2 python tools/time_wsplit.py 256 > gpurun_out/ncu_wsplit.log 2>&1
tail -2 gpurun_out/ncu_wsplit.log
