#!/bin/bash
# Round 2, GPU call D (2 GPUs): the sharded path.  P2P parity (pytest + tests/p2p_check.py), the peer-access probe
# (cudaMalloc+peer / CUDA IPC / VMM mappings vs footprint), then the slow-mode reproducer of profiles/r1_c5.md on the
# v5 kernel: 9 M points on 2 GPUs (13.3 ms in round 1 vs ~7 expected), with the shard-mapping and ownership variants.
#   gpurun --gpus 2 --timeout 1500 -- 'bash profiles/r2_call_d.sh > gpurun_out/r2d.log 2>&1; tail -60 gpurun_out/r2d.log'
mkdir -p gpurun_out
P=$PWD/bang-billion-scale-ann_b200
nvidia-smi topo -m | head -8
echo "== P2P parity"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "p2p" 2>&1 | tail -3
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$T 29511 tests/p2p_check.py 2>&1 | grep -E "rank|Error|error" | head -8
echo "== probe"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o /tmp/p2p_probe profiles/p2p_probe.cu -lcuda 2> gpurun_out/r2d_probe_build.log
S=64,732,1648,2930,5860,23438,46000
( for mode in local peer vmm; do timeout 120 /tmp/p2p_probe $mode $S; done
  for mib in 732 1648 5860 23438; do timeout 120 /tmp/p2p_probe ipc $mib; done
  P2P_PROBE_LD=1 timeout 120 /tmp/p2p_probe peer $S ) > gpurun_out/r2d_probe.jsonl 2> gpurun_out/r2d_probe.err
cat gpurun_out/r2d_probe.jsonl | cut -c1-260
short() { grep -E '^\{' | python -c "
import json,sys
for ln in sys.stdin:
    j=json.loads(ln)
    print('   N %d L %d own %s: kernel ms per rank %s  QPS %.0f  e2e %.0f  recall %.2f  hops %.0f cand %.0f' % (j['N'], j['L'], j['ownership'], j['kernel_ms_per_rank'], j['value'], j['e2e'], j['recall_at_10'], j['hops_per_query'], j['candidates_per_query']))
"; }
echo "== 9 M / 2 GPUs, L = 176: default (IPC, id mod G)"
$T 29551 profiles/c5_run.py 9e6 10000 176 2>gpurun_out/r2d_err.log | short
echo "== same, rows through the VMM API"
BANG_B200_SHARD_VMM=1 $T 29552 profiles/c5_run.py 9e6 10000 176 2>>gpurun_out/r2d_err.log | short
echo "== same, plain (coherent) loads for graph rows"
[ -f $P/libbang_b200_plain.so ] && BANG_B200_LIB=$P/libbang_b200_plain.so $T 29553 profiles/c5_run.py 9e6 10000 176 2>>gpurun_out/r2d_err.log | short
echo "== same, partition-owned rows + query routing"
C5_OWNERSHIP=partition $T 29554 profiles/c5_run.py 9e6 10000 176 2>>gpurun_out/r2d_err.log | short
echo "== 9 M on ONE GPU (no peer rows) for reference"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29555 profiles/c5_run.py 9e6 10000 176 2>>gpurun_out/r2d_err.log | short
echo "== 32 M / 2 GPUs, L = 176 and 256"
$T 29556 profiles/c5_run.py 32e6 10000 176,256 2>>gpurun_out/r2d_err.log | short
tail -5 gpurun_out/r2d_err.log
