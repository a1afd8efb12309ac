#!/bin/bash
# Round 2, GPU call H (2 GPUs): the default bench exactly as the driver launches it at N = 2 (torchrun, weak + strong
# figures), the reference arm at N = 2, then on one GPU: the ncu launch list of the bench command, the DRAM traffic of the
# search kernel on the default workload, and the GPU suite.
mkdir -p gpurun_out
echo "== prepare (C4 cache)"; S=$SECONDS
timeout 900 python bench.py --prepare 2> gpurun_out/r2h_prepare.err; echo "wall $((SECONDS-S)) s"; grep "\[bench\] prepared" gpurun_out/r2h_prepare.err | cut -c1-300
echo "== reference arm, --gpus 2 under torchrun"; S=$SECONDS
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2h_ref_2gpu.json 2> gpurun_out/r2h_ref_2gpu.err
echo "exit $? wall $((SECONDS-S)) s"; cut -c1-400 gpurun_out/r2h_ref_2gpu.json
echo "== bench --gpus 2 under torchrun"; S=$SECONDS
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2h_b200_2gpu.json 2> gpurun_out/r2h_b200_2gpu.err
echo "exit $? wall $((SECONDS-S)) s"; grep -E "Error|error|Traceback" gpurun_out/r2h_b200_2gpu.err | head -5; cut -c1-2500 gpurun_out/r2h_b200_2gpu.json
echo "== bench --gpus 2 --scaling strong"; S=$SECONDS
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong > gpurun_out/r2h_b200_2gpu_strong.json 2> gpurun_out/r2h_b200_2gpu_strong.err
echo "exit $? wall $((SECONDS-S)) s"; cut -c1-700 gpurun_out/r2h_b200_2gpu_strong.json
echo "== bench N = 1 (full line with other_configs)"; S=$SECONDS
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_b200_1gpu.json 2> gpurun_out/r2h_b200_1gpu.err
echo "exit $? wall $((SECONDS-S)) s"; cut -c1-1200 gpurun_out/r2h_b200_1gpu.json
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2h_launch_run.log 2>&1
tail -2 gpurun_out/r2h_launch_run.log | cut -c1-300; wc -l gpurun_out/r2h_launches.csv
echo "== ncu DRAM traffic of the search kernel, default workload"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:bang_search_kernelIfLi1ELi3E -s 6 -c 2 --csv --log-file gpurun_out/r2h_traffic.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2h_traffic_run.log 2>&1
tail -8 gpurun_out/r2h_traffic.csv | cut -c1-300
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12
