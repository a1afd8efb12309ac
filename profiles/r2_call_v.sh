#!/bin/bash
# Round 2, GPU call V (2 GPUs): the 2-GPU parity tests (sharded rows read over NVLink, both mapping schemes), then the default
# bench exactly as the driver launches it at N = 2 (torchrun; reference arm first; the line carries weak and strong scaling).
mkdir -p gpurun_out
nvidia-smi -L
echo "== 2-GPU parity tests"
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "two_gpus or sharded or p2p" 2>&1 | tail -4
echo "== reference arm, --gpus 2 under torchrun (prepares the C4 cache)"; S=$SECONDS
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2v_ref_2gpu.json 2> gpurun_out/r2v_ref_2gpu.err
echo "exit $? wall $((SECONDS-S)) s"; cut -c1-300 gpurun_out/r2v_ref_2gpu.json
echo "== bench --gpus 2 under torchrun"; S=$SECONDS
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2v_b200_2gpu.json 2> gpurun_out/r2v_b200_2gpu.err
echo "exit $? wall $((SECONDS-S)) s"; grep -E "Error|error|Traceback" gpurun_out/r2v_b200_2gpu.err | head -5; cut -c1-2200 gpurun_out/r2v_b200_2gpu.json
