#!/bin/bash
# GPU call K: ncu --set full (with source) of the current build on one shape.  usage: r2_call_k.sh <shape> <L> <tag>
mkdir -p gpurun_out
timeout 600 python profiles/prof_search.py $2 inmemory 4 $1 2>&1 | grep -E "^run|Error" | tail -3
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:bang_search_kernel -s 1 -c 1 -o gpurun_out/$3_$1 -f python profiles/prof_search.py $2 inmemory 3 $1 > gpurun_out/$3_ncu_$1.log 2>&1; tail -1 gpurun_out/$3_ncu_$1.log
