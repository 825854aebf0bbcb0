export BFL_OVERLAP=1
for cfg in "0 0 4" "0 0 8" "2 2 8" "2 1 8" "1 3 8" "2 3 8" "3 2 8"; do set -- $cfg; echo "capC $1 capE $2 chunks $3"; BFL_OV_CAPC=$1 BFL_OV_CAPE=$2 BFL_CHUNKS=$3 python tools/time_window.py 2>&1 | tail -1 | cut -c1-60; done
