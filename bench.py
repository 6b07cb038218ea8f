#!/usr/bin/env python
"""bench.py — headline benchmark of the HPR hot path (BASELINE.json metric:
"HPR audio-sec/sec batched at 1/2/4/8 B200; p50 per-hop latency @1024 hop").

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path on the host cores

Workload (BASELINE.json configs[4]): a batch of 4096 independent synthetic 60 s
mono 44.1 kHz streams per GPU, real-time HPR (HPRRealtime<GPU> semantics: causal,
copy-border, percussive output, hard mask) at hop 1024, beta 2.5.  One step = one
pass of the hot path over the whole batch.  Under torchrun every rank owns one
GPU and its own 4096 streams (no collective on the data path; weak scaling).

Prints ONE JSON line (rank 0).  See DESIGN.md for the definition of every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 44100
HOP = 1024
BETA = 2.5
SECONDS = 60
N_STREAMS = 4096
BYTES_PER_HOP = 4 * HOP + 4 * HOP  # algorithmic HBM traffic: one hop in, one percussive hop out


def fakert_hops(n_samples, hop):
    """zen/fakert.h:15-34 get_chunk_limits: full hops strictly before size - hop"""
    return 0 if n_samples <= hop else -(-(n_samples - hop) // hop)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ synthetic ---

def synth_batch_device(torch, n_streams, n_samples, device, seed0, chunk=64):
    """tones + decaying noise bursts + noise floor per stream (zen_b200/synth.py recipe), generated on the device
    so that the 43 GB batch never has to be uploaded"""
    x = torch.empty((n_streams, n_samples), dtype=torch.float32, device=device)
    t = torch.arange(n_samples, device=device, dtype=torch.float64) / FS
    idx = torch.arange(n_samples, device=device, dtype=torch.int32)
    tau = 0.005 * FS
    for s0 in range(0, n_streams, chunk):
        ns = min(chunk, n_streams - s0)
        rng = np.random.default_rng(seed0 + s0)
        g = torch.Generator(device=device)
        g.manual_seed(seed0 + s0)
        f1 = torch.from_numpy(rng.uniform(110.0, 880.0, ns)).to(device)
        acc = torch.zeros((ns, n_samples), dtype=torch.float32, device=device)
        for k in range(1, 5):
            ph = torch.remainder(f1[:, None] * k * t[None, :], 1.0)
            acc += (0.4 / k) * torch.sin(2.0 * np.pi * ph).float()
        last = torch.zeros((ns, n_samples), dtype=torch.int32, device=device)
        for r in range(ns):
            gaps = rng.uniform(0.25, 0.5, int(n_samples / FS / 0.25) + 2)
            pos = (np.cumsum(gaps) * FS).astype(np.int64)
            pos = pos[pos < n_samples]
            last[r, torch.from_numpy(pos).to(device)] = torch.from_numpy(pos.astype(np.int32)).to(device)
        last = torch.cummax(last, dim=1).values
        env = torch.where(last > 0, torch.exp(-(idx[None, :] - last).float() / tau), torch.zeros((), device=device))
        acc += 0.5 * env * torch.randn((ns, n_samples), generator=g, device=device)
        acc += 0.01 * torch.randn((ns, n_samples), generator=g, device=device)
        x[s0:s0 + ns] = acc.clamp_(-1.0, 1.0)
        del acc, last, env
    return x


# ------------------------------------------------------------- CPU baseline ---

def cpu_workers(kind, cores, n_hops, seed0):
    """one independent stream per host core, all cores at once (the reference is single-threaded per stream)"""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "oracle", "cpu_worker.py"), kind, str(HOP), str(BETA),
                               str(n_hops), str(seed0 + c)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
             for c in range(cores)]
    res = []
    for p in procs:
        out, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("cpu worker failed: " + err[-500:])
        res.append(json.loads(out.strip().splitlines()[-1]))
    wall = max(r["seconds"] for r in res)
    audio = sum(r["audio_s"] for r in res)
    return audio / wall, wall, audio


def cpu_kind():
    return "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libzen_ref.so")) else "port"


def cpu_sample_text(kind, cores, n_hops):
    what = ("reference libzen/hps.cu CPU dataflow (HPRRealtime semantics), IPP calls served by oracle/ref/ippstub"
            if kind == "reference" else "oracle/hpr_oracle.c restatement of the reference CPU path")
    return "%d streams x %d hops of hop %d (one stream per core), %s" % (cores, n_hops, HOP, what)


# ----------------------------------------------------------------- reference arm ---

def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    if kind == "port":
        from oracle import oraclebind
        oraclebind.lib()
    n_hops = args.ref_hops
    for w in range(args.warmup):
        cpu_workers(kind, cores, max(8, n_hops // 8), 7000 + 100 * w)
    t_tot, audio_tot = 0.0, 0.0
    for k in range(args.steps):
        _, wall, audio = cpu_workers(kind, cores, n_hops, 9000 + 100 * k)
        t_tot += wall
        audio_tot += audio
    value = audio_tot / t_tot
    line = {
        "impl": "reference", "metric": "HPR audio-sec/sec batched", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, N_STREAMS, SECONDS),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": kind,
                         "sample": cpu_sample_text(kind, cores, n_hops) + " per step"},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_streams, seconds):
    return {"workload": "BASELINE.json configs[4]: %d independent synthetic %d s mono 44.1 kHz streams per GPU, real-time HPR "
                        "(causal, copy-border, percussive out, hard mask), hop %d, beta %.1f" % (n_streams, seconds, HOP, BETA),
            "streams_per_gpu": n_streams, "hops_per_stream": fakert_hops(seconds * FS, HOP), "hop": HOP, "nfft": 4 * HOP,
            "fs": FS, "l2": ("inputs exceed L2 (%.1f GB in + %.1f GB out per GPU per step)" % (2 * (n_streams * seconds * FS * 4 / 1e9,))
                             if n_streams * seconds * FS * 4 >= (512 << 20) else "small run: L2 flushed between steps"),
            "parallelism": "streams sharded over GPUs, no collective"}


# ----------------------------------------------------------------------- our arm ---

def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from zen_b200 import _lib, hps, shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; zen_b200 has no CPU fallback")
    _lib.lib()  # fail loudly if the CUDA library is missing
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_streams, seconds = args.streams, args.seconds
    n_hops = fakert_hops(seconds * FS, HOP)
    n = n_hops * HOP
    x = synth_batch_device(torch, n_streams, n, dev, seed0=shard.stream_seed(1000, rank, n_streams, 0))
    out_p = torch.empty_like(x)
    b = hps.HPRBatch(float(FS), HOP, BETA, hps.OUTPUT_PERCUSSIVE)
    small = n_streams * n * 4 < (512 << 20)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if small else None

    def step():
        if flush is not None:
            flush.zero_()
        b.process(x, [None, out_p, None])

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out_p[:4]).all()) and float(out_p[:4].abs().max()) > 0, "kernel produced no output"
    # full-size sanity: a few streams of the batch must equal, bit for bit, the same stream fed hop by hop through
    # HPRRealtime-style per-hop launches (the path the parity tests pin against the oracle)
    verified = []
    for sidx in sorted(set([0, n_streams // 2, n_streams - 1])):
        hh = hps.HPR(float(FS), HOP, BETA, hps.OUTPUT_PERCUSSIVE, 0, True)
        n_chk = min(n_hops, 400)
        ref_p = hh.run(x[sidx, : n_chk * HOP].cpu().numpy(), n_chk)[1]
        hh.close()
        assert np.array_equal(out_p[sidx, : n_chk * HOP].cpu().numpy(), ref_p), "batched kernel != per-hop path on stream %d" % sidx
        verified.append(int(sidx))

    sampler = ClockSampler(local_rank)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        if flush is not None:
            flush.zero_()
        ev[2 * k].record()
        b.process(x, [None, out_p, None])
        ev[2 * k + 1].record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    kern_ms = [ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1]) if flush is None else sum(kern_ms)
    launches = args.steps * b.last_launches
    if world > 1:
        dist.barrier()
    total_ms_max = shard.max_over_ranks(dist if world > 1 else None, total_ms, dev)
    audio_per_step = n_streams * n / FS
    value = shard.aggregate_throughput(audio_per_step * args.steps, world, total_ms_max * 1e-3)

    # ---- end to end through the C ABI with host buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    try:
        # pinned host buffers for the whole batch when the host has room for them (2 x 43 GB per rank at full size);
        # otherwise the largest prefix of the streams that fits, and the dict says so
        e2e_streams = n_streams
        try:
            import psutil
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
            budget = 0.6 * psutil.virtual_memory().available / max(1, local_world)
            e2e_streams = int(max(1, min(n_streams, budget // (2 * n * 4))))
        except Exception:  # noqa: BLE001
            pass
        h_in = hps.PinnedArray(e2e_streams, n)
        h_out = hps.PinnedArray(e2e_streams, n)
        _lib.check(_lib.lib().zen_copy_to_host(h_in.ptr, x.data_ptr(), h_in.nbytes), "zen_copy_to_host")
        torch.cuda.synchronize()
        del out_p
        audio_per_e2e_step = e2e_streams * n / FS
        e2e_steps = max(1, min(args.steps, 3))
        b.process_host(h_in.array, [None, h_out.array, None])  # warm-up (allocates the staging buffers)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            b.process_host(h_in.array, [None, h_out.array, None])
        dt = time.perf_counter() - t0
        launches_e2e = e2e_steps * b.last_launches
        dt_max = shard.max_over_ranks(dist if world > 1 else None, dt, dev)
        e2e = {"value": shard.aggregate_throughput(audio_per_e2e_step * e2e_steps, world, dt_max), "unit": "audio-s/s",
               "h2d_bytes_per_step": int(e2e_streams) * n * 4, "d2h_bytes_per_step": int(e2e_streams) * n * 4,
               "streams_per_gpu": int(e2e_streams),
               "steps": e2e_steps, "kernel_launches_per_step": launches_e2e // e2e_steps,
               "api": "zen_hpr_batch_process_host (pinned host buffers in and out)"}
        assert bool(np.isfinite(h_out.array[:2]).all()) and float(np.abs(h_out.array[:2]).max()) > 0
        h_in.close()
        h_out.close()
    except Exception as exc:  # noqa: BLE001
        e2e = {"value": None, "unit": "audio-s/s", "error": repr(exc)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    kern_avg_ms = float(np.mean(kern_ms))
    achieved = n_streams * n_hops * BYTES_PER_HOP / (kern_avg_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            # measured on the full 4096 x 2583-hop launch; scaled to this run's launch size
            traffic = tj.get("dram_bytes_per_launch") * (n_streams * n_hops * BYTES_PER_HOP) / tj.get("algorithmic_bytes_per_launch")
        except Exception:  # noqa: BLE001
            traffic = None

    line = {
        "metric": "HPR audio-sec/sec batched", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n_streams, seconds),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "kernel": "hpr_tile_kernel<4096,256>", "kernel_ms": kern_avg_ms,
                     "algorithmic_bytes_per_launch": n_streams * n_hops * BYTES_PER_HOP,
                     "note": "the fused kernel is ALU/issue-bound on the median selection, not HBM-bound (DESIGN.md)"},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "wall_s_timed_region": t_wall,
        "verified_streams_bit_exact_vs_per_hop_path": verified,
    }

    # ---- per-hop latency of one real-time stream (BASELINE.json configs[1])
    if not args.no_latency:
        try:
            line["latency"] = measure_latency(args)
        except Exception as exc:  # noqa: BLE001
            line["latency"] = {"error": repr(exc)[:200]}

    # ---- CPU baseline on the host cores (rank 0, N == 1 only)
    if world == 1 and not args.no_cpu:
        try:
            cores = os.cpu_count() or 1
            kind = cpu_kind()
            v, wall, audio = cpu_workers(kind, cores, args.cpu_hops, 5000)
            line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": kind,
                                    "sample": cpu_sample_text(kind, cores, args.cpu_hops), "wall_s": wall}
        except Exception as exc:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "audio-s/s", "error": repr(exc)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_latency(args):
    """zen fakert's timed region per hop (zen/fakert.h:221-247) on one causal stream, hop 1024:
    host copy-in -> process -> copy_percussive -> host copy-out, timed inside the library with std::chrono."""
    import ctypes
    from zen_b200 import _lib
    from zen_b200.synth import MIXED_WAV_SAMPLES, synth_audio
    L = _lib.lib()
    n_h = args.latency_hops
    mixed = synth_audio(MIXED_WAV_SAMPLES, seed=1)
    a = np.tile(mixed, (n_h * HOP) // mixed.size + 1)[: n_h * HOP].copy()
    res = {}
    first = None
    same = True
    for name, fused in (("two_call", 0), ("fused_call", 1), ("resident_kernel", 2), ("resident_two_call", 3)):
        perc = np.zeros(n_h * HOP, dtype=np.float32)
        us = np.zeros(n_h, dtype=np.float64)
        _lib.check(L.zen_fakert_run(float(FS), HOP, BETA, 0, a.ctypes.data, n_h, 1000, fused, perc.ctypes.data, us.ctypes.data),
                   "zen_fakert_run")
        res[name] = {"p50_us": float(np.median(us)), "p99_us": float(np.percentile(us, 99)), "mean_us": float(us.mean())}
        if first is None:
            first = perc
        else:
            same = same and bool(np.array_equal(first, perc))
    res["outputs_bit_identical_across_call_styles"] = same
    res["n_hops"] = n_h
    res["region"] = "zen/fakert.h:221-247 (host copy-in, process_next_hop, copy_percussive, host copy-out)"
    res["two_call_api"] = "HPRRealtime::process_next_hop + copy_percussive (2 launches)"
    res["fused_call_api"] = "zen_hpr_process_hop_io (1 launch)"
    res["resident_kernel_api"] = ("zen_hpr_realtime_begin + zen_hpr_process_hop_io (persistent 4-CTA cluster kernel, hop pushed / output returned as tagged "
                                  "16-byte groups in mapped memory, 0 launches and 0 fences per hop)")
    res["resident_two_call_api"] = "zen_hpr_realtime_begin, then the reference's own pair HPRRealtime::process_next_hop + copy_percussive"
    res["p50_us"] = res["resident_kernel"]["p50_us"]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=N_STREAMS)
    ap.add_argument("--seconds", type=int, default=SECONDS)
    ap.add_argument("--cpu-hops", type=int, default=600, help="hops per core of the cpu_baseline sample")
    ap.add_argument("--ref-hops", type=int, default=1500, help="hops per core and step of the reference arm")
    ap.add_argument("--latency-hops", type=int, default=2000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    from zen_b200 import shard
    rank, world, local_rank = shard.rank_world()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
