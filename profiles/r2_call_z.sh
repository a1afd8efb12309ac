#!/bin/bash
# Round 2, GPU call Z (1 GPU): smoke() as the driver runs it; DRAM bytes of the search kernel at the other worklist length a
# fresh index build can pick for recall 0.90 on C4 (L = 40) and on C2, for profiles/traffic.json.
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4
echo "== prepare C4"; timeout 900 python bench.py --prepare 2> gpurun_out/r2z_prepare.err; grep "L90" gpurun_out/r2z_prepare.err | cut -c1-200 | tail -1
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base mangled -k regex:bang_search_kernel --csv"
timeout 600 ncu $M -c 16 --log-file gpurun_out/r2z_c4_L40.csv python bench.py --L 40 --L95 48 --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2z_c4.log 2>&1
grep -c bang_search gpurun_out/r2z_c4_L40.csv
timeout 300 ncu $M -c 16 --log-file gpurun_out/r2z_c2.csv python bench.py --workload sift1m --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2z_c2.log 2>&1
grep -c bang_search gpurun_out/r2z_c2.csv; tail -2 gpurun_out/r2z_c2.log | cut -c1-300
