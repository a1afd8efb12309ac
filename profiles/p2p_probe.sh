#!/bin/bash
# Next-round first step (needs 2 GPUs): gpurun --gpus 2 --timeout 600 -- 'bash profiles/p2p_probe.sh > gpurun_out/p2p_probe.jsonl 2> gpurun_out/p2p_probe.err'
# Footprints = the per-GPU row buffers of the measured C5 runs (MiB): 4M/2 732, 9M/2 1648 (slow), 64M/8 2930 (slow),
# 32M/2 5860 (fast), 256M/4 23438 (slow), plus 64 (L2-resident) and 46000 (1e9 points / 8 GPUs).
set -e
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o /tmp/p2p_probe profiles/p2p_probe.cu -lcuda
S=64,732,1648,2930,5860,23438,46000
for mode in local peer vmm; do /tmp/p2p_probe $mode $S; done
for mib in 64 732 1648 2930 5860 23438 46000; do
  /tmp/p2p_probe ipc $mib
  /tmp/p2p_probe ipc $mib 16 2000 4 1     # owner heap fragmented first
done
/tmp/p2p_probe peer $S 16 2000 4 1
/tmp/p2p_probe peer 1648,5860 4           # concurrency dependence
/tmp/p2p_probe ipc 1648 4
# the search kernel's load form (ld.global.nc.L1::no_allocate + evict_first hint) instead of a plain load
P2P_PROBE_LD=1 /tmp/p2p_probe peer $S
for mib in 732 1648 5860 23438; do P2P_PROBE_LD=1 /tmp/p2p_probe ipc $mib; done
# then the A/B on the real path (2 GPUs, 9 M points: 13.3 ms in the slow mode, ~7 ms expected):
#   T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 profiles/c5_run.py 9e6 10000 176"
#   $T ; BANG_B200_SHARD_VMM=1 $T
#   C5_OWNERSHIP=partition $T        # rows owned by their partition's GPU + query routing (profiles/locality_sim.py: ~86-97 % local hops)
