"""Sharded-graph (C5 design) vs replicated timing on G GPUs of one box, C2-shape index.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29541 profiles/p2p_bench.py [L]
Every rank searches its own 10 000-query batch.  sharded: graph rows id % G on each GPU, peers read with P2P loads."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bang_b200
from bang_b200 import api, builder, formats, recall, sharding
L = int(sys.argv[1]) if len(sys.argv) > 1 else 176
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
prefix = "/tmp/bang_p2p/u8_1m"
if rank == 0 and not os.path.exists(prefix + "_gt.bin"):
    builder.make_fixture_auto(prefix, 1_000_000, 128, "uint8", 10000 * world, 32, device=torch.device("cuda", 0))
dist.barrier()
q_all = formats.read_bin(prefix + "_query.bin", np.uint8)
gi, gd = formats.read_truthset(prefix + "_gt.bin")
sl = sharding.rank_batch(10000, rank)
q = np.ascontiguousarray(q_all[sl])
for label, shard in (("replicated", False), ("sharded", True)):
    s = api.BANGSearch("uint8", "inmemory", device=local)
    if shard:
        s.set_sharding(rank, world)
    assert s.bang_load(prefix), s.last_error
    if shard:
        sharding.exchange_shards(s, rank, world)
    s.set_dists_layout(1); s.bang_set_searchparams(10, L); s.bang_alloc(len(q))
    ms = []
    for r in range(4):
        s.bang_init(len(q)); dist.barrier(); ids, d = s.bang_query(q); ms.append(s.last_timing().kernel_ms)
    worst = sharding.max_over_ranks([min(ms[1:])])[0]
    rec = recall.calculate_recall(gi[sl], gd[sl], ids, 10)
    if rank == 0:
        print(f"{label:10s} G={world} L={L}: kernel {worst:.3f} ms (max over ranks) -> {world * len(q) / worst * 1e3:,.0f} QPS total, "
              f"rows in HBM per GPU {s.info().device_bytes / 2**20:.0f} MiB, recall {rec:.2f}", flush=True)
    dist.barrier()
    s.bang_free(); s.bang_unload()
dist.destroy_process_group()
