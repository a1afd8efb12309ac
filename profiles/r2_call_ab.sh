#!/bin/bash
# Round 2, GPU call AB (1 GPU): DRAM bytes per launch of the final search kernel on the default workload at the two worklist
# lengths a fresh index build picks for recall 0.90 (L = 36 / 40), for profiles/traffic.json.
mkdir -p gpurun_out
timeout 170 python bench.py --prepare 2> gpurun_out/r2ab_prepare.err; grep "L90" gpurun_out/r2ab_prepare.err | cut -c1-120 | tail -1
timeout 60 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base mangled -k regex:bang_search_kernel --csv -c 16 --log-file gpurun_out/r2ab_c4_L36_L40.csv python bench.py --L 36 --L95 40 --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2ab_c4.log 2>&1
grep -c bang_search gpurun_out/r2ab_c4_L36_L40.csv
