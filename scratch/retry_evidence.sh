for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scratch/run_evidence1.sh' 2>&1)
  echo "$out" | grep -v "^\[gpurun\] sending" | tail -16
  if echo "$out" | grep -q "status=ok\|status=fail"; then break; fi
  sleep 200
done
