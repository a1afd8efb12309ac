#!/bin/bash
# Round 2, GPU call I (1 GPU): dynamic instruction profile (ncu --set full with source) of the current kernel at 24 query
# warps per SM on the C2 shape (u8, CS = 4, L = 176) and on C4's shape at 10^7 points (float, D = 96, L = 36), + timings.
mkdir -p gpurun_out
for sh in "sift1m 176" "deep10m 36"; do set -- $sh
  echo "== $1 L=$2: timing"; timeout 600 python profiles/prof_search.py $2 inmemory 4 $1 2>&1 | grep -E "^run|Error" | tail -4
  echo "== $1 L=$2: ncu"
  timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:bang_search_kernel -s 1 -c 1 -o gpurun_out/r2i_$1 -f python profiles/prof_search.py $2 inmemory 3 $1 > gpurun_out/r2i_ncu_$1.log 2>&1; tail -2 gpurun_out/r2i_ncu_$1.log
done
ls -la gpurun_out/
