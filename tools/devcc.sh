#!/bin/bash
# developer helper: compile bfl_kernels.cu to an object with resource usage, dump the SASS of one kernel
#   tools/devcc.sh <mangled-name-regex> [extra nvcc flags]
pat="$1"; shift
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xptxas -v -I include "$@" \
  -c bayesflare_b200/csrc/bfl_kernels.cu -o /tmp/k.o 2>&1 | grep -A2 "$pat" | grep -v "^--" 
