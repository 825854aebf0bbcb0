# ncu of the split weighted form: --set full of one k_wcorr and one k_wsolve launch (256 curves -> sub-batches of 128),
# after the same command ran clean without the profiler; then the new split-vs-one-kernel test
timeout 300 python tools/time_wsplit.py 256 1024 > gpurun_out/plain_wsplit.log 2>&1 && tail -2 gpurun_out/plain_wsplit.log &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wcorr|k_wsolve" -c 2 -s 2 -f -o gpurun_out/wsplit_r2 python tools/time_wsplit.py 256 > gpurun_out/ncu_wsplit.log 2>&1
tail -1 gpurun_out/ncu_wsplit.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split_form_in_chunks" 2>&1 | tail -3
