"""Multi-GPU plumbing of the batched path: independent streams are sharded over
ranks (one process per GPU), no collective touches the data path; the only
communication is a barrier and a MAX-reduction of the measured time.
Backend-agnostic (nccl on GPUs, gloo in the CPU tests)."""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def stream_seed(base, rank, streams_per_rank, s):
    """global, collision-free seed of local stream s on `rank` (weak scaling: every rank owns streams_per_rank streams)"""
    return base + rank * streams_per_rank + s


def shard_range(n_total, rank, world):
    """contiguous block of a fixed total (strong scaling); blocks differ by at most one stream"""
    q, r = divmod(n_total, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def max_over_ranks(dist, value, device=None):
    """max of a python float over all ranks (identity when not distributed)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank, world, seconds_max):
    """whole-job throughput: units all ranks processed / slowest rank's time"""
    return world * units_per_rank / seconds_max


def gpu_numa_node(local_rank):
    """NUMA node of CUDA device `local_rank` (its PCI function's numa_node in sysfs), or -1 when unknown"""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        return int(open(path).read().strip())
    except Exception:  # noqa: BLE001
        return -1


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPUs of its GPU's NUMA node, so that the pinned host buffers it allocates afterwards
    (first touch) and the threads that drive the copies sit next to the PCIe root the GPU hangs off.  No-op on
    single-node hosts or when sysfs says nothing.  Returns a dict describing what was done (for the bench line)."""
    info = {"numa_nodes": 0, "gpu_node": -1, "bound": False}
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
        info["numa_nodes"] = len(nodes)
        node = gpu_numa_node(local_rank)
        info["gpu_node"] = node
        if len(nodes) < 2 or node < 0:
            return info
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
            info["cpus"] = len(allowed)
    except Exception as exc:  # noqa: BLE001
        info["error"] = repr(exc)[:100]
    return info
