# weighted path, split form (bfl_wsplit.cu): parity tests, then timing against the one-kernel form
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "weighted or kepler or varying" > gpurun_out/wsplit_tests.log 2>&1; tail -3 gpurun_out/wsplit_tests.log
BFL_NO_WSPLIT=1 timeout 300 python tools/time_weighted.py 512 2>&1 | tail -1
timeout 300 python tools/time_weighted.py 512 2>&1 | tail -2
timeout 300 python tools/time_wsplit.py 256 1024 2>&1 | tail -2
