"""Raw host<->device link bandwidth at N ranks (the floor under bench.py's e2e number).

    python tools/link_bw.py                      # one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/link_bw.py

Every rank copies a pinned host buffer to its GPU and another one back, first one direction at a time, then both
directions concurrently on two streams (what the host pipeline of zen_hpr_batch_process_host does), all ranks at
once.  Rank 0 prints one JSON line: per-rank and whole-node GB/s, plus what the box says about its topology."""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

from zen_b200 import shard


def main():
    rank, world, local_rank = shard.rank_world()
    torch.cuda.set_device(local_rank)
    numa = shard.bind_to_gpu_numa(local_rank) if os.environ.get("ZEN_NO_NUMA_BIND") is None else {"bound": False}
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mb = int(os.environ.get("LINK_MB", "1024"))
    reps = int(os.environ.get("LINK_REPS", "8"))
    h_in = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_in = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    d_out = torch.ones(mb << 20, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt = shard.max_over_ranks(dist if world > 1 else None, dt, dev)
        return reps * (mb << 20) / dt / 1e9

    run(True, True)
    res = {"h2d_only_gbs_per_gpu": run(True, False), "d2h_only_gbs_per_gpu": run(False, True), "both_gbs_per_direction_per_gpu": run(True, True)}
    if rank == 0:
        res.update({"n_gpus": world, "buffer_mb": mb, "reps": reps, "numa": numa, "cpus": os.cpu_count(),
                    "node_gbs_per_direction_concurrent": res["both_gbs_per_direction_per_gpu"] * world})
        try:
            res["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[-3000:]
            res["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
            res["mem_gb"] = round(os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 1e9, 1)
        except Exception as exc:  # noqa: BLE001
            res["topo_error"] = repr(exc)[:100]
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
