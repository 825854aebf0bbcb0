python tools/time_stage.py 1024 4400 epi2
python - <<'PY'
import numpy as np
a=np.load("gpurun_out/stage_epi1.npy"); b=np.load("gpurun_out/stage_epi2.npy")
print("epi1 vs epi2 max scaled diff %.3e" % np.max(np.abs(a-b)/np.maximum(1,np.abs(a))))
PY
NO_ORACLE=1 python tools/time_stage.py 128 4400
NO_ORACLE=1 python tools/time_stage.py 256 4400
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
