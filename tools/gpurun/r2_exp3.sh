python tools/time_stage.py 1024 4400 epi2
NO_ORACLE=1 python tools/time_stage.py 128 4400
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/gpurun/r2_ncu_epi.sh 1024 win_r2c
