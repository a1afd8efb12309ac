#!/bin/bash
# Round 2, GPU call N (1 GPU): evidence for the round's last kernel (commit 10d3dff): the GPU suite, both bench arms as
# the driver runs them (reference first), the ncu launch list of the bench command, the DRAM traffic of the search kernel
# on the default workload, and one `ncu --set full` capture of the C2 kernel at the default 24 warps.
mkdir -p gpurun_out
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8
echo "== bench --impl reference (default workload; prepares the C4 cache)"; S=$SECONDS
timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_ref.json 2> gpurun_out/r2n_ref.err
echo "exit $? wall $((SECONDS-S)) s"; cut -c1-500 gpurun_out/r2n_ref.json
echo "== bench N = 1 (full line with other_configs)"; S=$SECONDS
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2n_b200_1gpu.json 2> gpurun_out/r2n_b200_1gpu.err
echo "exit $? wall $((SECONDS-S)) s"; grep -E "Error|Traceback" gpurun_out/r2n_b200_1gpu.err | head; cut -c1-1500 gpurun_out/r2n_b200_1gpu.json
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2n_launch_run.log 2>&1
tail -2 gpurun_out/r2n_launch_run.log | cut -c1-300; wc -l gpurun_out/r2n_launches.csv
echo "== ncu DRAM traffic of the search kernel, default workload"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:bang_search_kernelIfLi1ELi3E -s 6 -c 2 --csv --log-file gpurun_out/r2n_traffic.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2n_traffic_run.log 2>&1
tail -8 gpurun_out/r2n_traffic.csv | cut -c1-300
echo "== ncu --set full, sift1m (C2), default warps"
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:bang_search_kernelIhLi1ELi4E -s 1 -c 1 -o gpurun_out/r2n_c2 -f python profiles/prof_search.py 176 inmemory 3 > gpurun_out/r2n_ncu.log 2>&1; tail -3 gpurun_out/r2n_ncu.log
ls -la gpurun_out | tail -12
