"""world_size-2 gloo run of the multi-GPU plumbing (zen_b200/shard.py) on CPU:
streams are sharded without overlap, no data-path collective, time is the MAX
over ranks.  The per-rank "step" is the oracle on a tiny stream, standing in
for the kernel, so the whole thing runs without a GPU."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oraclebind as ob
    from zen_b200 import shard
    from zen_b200.synth import synth_audio
    r, w, _ = shard.rank_world()
    assert (r, w) == (rank, world)
    streams_per_rank, hop, n_hops = 3, 256, 12
    seeds = [shard.stream_seed(1000, r, streams_per_rank, s) for s in range(streams_per_rank)]
    checks = []
    for sd in seeds:
        o = ob.OracleHPR(ob.GEOM_GPU, 44100.0, hop, 2.5, ob.OUT_P, ob.CAUSAL, True)
        out = o.run(synth_audio(n_hops * hop, seed=sd), want=(False, True, False))
        checks.append(float(np.abs(out[1]).sum()))
    dist.barrier()
    my_time = 1.0 + rank          # rank 1 is "slower"
    t_max = shard.max_over_ranks(dist, my_time)
    gathered = [None] * world
    dist.all_gather_object(gathered, (seeds, checks, shard.shard_range(10, r, w)))
    q.put((rank, t_max, gathered))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from zen_b200 import shard
    for rank, t_max, gathered in res:
        assert t_max == 2.0                                  # MAX over ranks, not this rank's own time
        all_seeds = [s for g in gathered for s in g[0]]
        assert len(all_seeds) == len(set(all_seeds)) == 6    # disjoint shards
        assert all(c > 0 for g in gathered for c in g[1])
        ranges = [g[2] for g in gathered]
        assert ranges == [(0, 5), (5, 10)]                   # strong-scaling split covers every stream once
    assert shard.aggregate_throughput(100.0, 2, 2.0) == 100.0
