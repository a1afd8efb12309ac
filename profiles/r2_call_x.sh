#!/bin/bash
# Round 2, GPU call X (2 GPUs): the sharded configuration (C5 path: sharded builder, device-resident hand-over, VMM shards, NVLink
# P2P row reads in the kernel) through bench.py at 32 M points, on the round's final kernel.
mkdir -p gpurun_out
S=$SECONDS
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py \
  --workload sift256m --points 32000000 --gpus 2 --steps 5 --warmup 2 > gpurun_out/r2x_sharded_32m_2gpu.json 2> gpurun_out/r2x_sharded_32m_2gpu.err
echo "exit $? wall $((SECONDS-S)) s"
grep -E "\[c5\]|\[bench c5\]|Error|error|Traceback" gpurun_out/r2x_sharded_32m_2gpu.err | cut -c1-300 | tail -12
grep "^{" gpurun_out/r2x_sharded_32m_2gpu.json | cut -c1-3000
