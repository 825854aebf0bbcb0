BFL_WSPLIT_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_detector_no_noise_models" > gpurun_out/dbg.log 2>&1
grep "wsplit\]" gpurun_out/dbg.log | head; tail -3 gpurun_out/dbg.log
BFL_WSPLIT_DEBUG=1 timeout 300 python tools/time_weighted.py 1024 2>&1 | grep "wsplit\]\|units" | head -20
