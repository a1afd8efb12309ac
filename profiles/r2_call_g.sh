#!/bin/bash
# Round 2, GPU call G (N GPUs): the sharded configuration through bench.py.
#   gpurun --gpus 2 --timeout 1500 -- 'bash profiles/r2_call_g.sh sift256m 2 > gpurun_out/r2g_256m.log 2>&1; tail -30 gpurun_out/r2g_256m.log'
#   gpurun --gpus 8 --timeout 1500 -- 'bash profiles/r2_call_g.sh sift1b 8   > gpurun_out/r2g_1b.log 2>&1;   tail -30 gpurun_out/r2g_1b.log'
W=${1:-sift256m}; G=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8
S=$SECONDS
timeout 1400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29571 bench.py \
  --workload $W --gpus $G --steps 5 --warmup 2 > gpurun_out/r2g_${W}_${G}gpu.json 2> gpurun_out/r2g_${W}_${G}gpu.err
echo "exit $? wall $((SECONDS-S)) s"
grep -E "\[c5\]|\[bench c5\]|Error|error|Traceback" gpurun_out/r2g_${W}_${G}gpu.err | cut -c1-400 | tail -40
cut -c1-4000 gpurun_out/r2g_${W}_${G}gpu.json
