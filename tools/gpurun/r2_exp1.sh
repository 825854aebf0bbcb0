# round 2, experiment 1: new mixed-precision epilogue vs round-1 epilogue, strong-scaling shard size, PCIe rates
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
python tools/pcie_probe.py
python tools/time_stage.py 1024 4400 epi2
BFL_EPI=1 NO_ORACLE=1 python tools/time_stage.py 1024 4400 epi1
python - <<'PY'
import numpy as np
a=np.load("gpurun_out/stage_epi1.npy"); b=np.load("gpurun_out/stage_epi2.npy")
print("epi1 vs epi2 max scaled diff %.3e" % np.max(np.abs(a-b)/np.maximum(1,np.abs(a))))
PY
NO_ORACLE=1 python tools/time_stage.py 128 4400
BFL_EPI=1 NO_ORACLE=1 python tools/time_stage.py 128 4400
NO_ORACLE=1 BFL_KT=2 python tools/time_stage.py 128 4400
NO_ORACLE=1 python tools/time_stage.py 256 4400
NO_ORACLE=1 python tools/time_stage.py 512 4400
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
