# launch list of the weighted call (split form): per-launch times of one 256-curve bfl_detector_odds through the host API
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_wsplit_r2.csv python tools/time_wsplit.py 256 > gpurun_out/ncu_list_wsplit.log 2>&1
tail -1 gpurun_out/ncu_list_wsplit.log; wc -l gpurun_out/launches_wsplit_r2.csv
