"""Turn one closing capture (tools/gpu_final.sh -> gpurun_out/final/) into the tracked files under profiles/.
Run here (ncu and cuobjdump are installed; no GPU needed):  python tools/make_profiles.py r02"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src = os.path.join(ROOT, "gpurun_out", sys.argv[2] if len(sys.argv) > 2 else "final")
dst = os.path.join(ROOT, "profiles")


def cp(a, b):
    if os.path.exists(os.path.join(src, a)):
        shutil.copy(os.path.join(src, a), os.path.join(dst, "%s_%s" % (tag, b)))


for a, b in [("bench.json", "bench_n1.json"), ("rt_latency.json", "rt_latency.json"), ("rt_phases_hop1024.txt", "rt_phases_hop1024.txt"),
             ("batch_sweep.json", "batch_sweep.json"), ("other_configs.json", "other_configs.json"), ("mfilt_bench.json", "mfilt_bench_vs_npp.json"),
             ("box_bench.json", "box_bench.json"), ("fft_bench.json", "fft_bench_vs_cufft.json"),
             ("link_bw_n1.json", "link_bw_n1.json"), ("long_parity.json", "long_parity.json"), ("gpu_tests.log", "gpu_tests.txt"),
             ("hps_bench_sweep.json", "hps_bench_sweep.json"), ("stream_bench.json", "stream_bench.json"), ("tma_poll_microbench.json", "tma_poll_microbench.json"),
             ("handoff_litmus_c8.json", "handoff_litmus_c8.json"), ("pcm_bench.json", "pcm_bench.json")]:
    cp(a, b)

# ---- ncu --set full captures -> text summaries + the json bench.py reads
def summary(rep, out):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, one launch; summarised by tools/ncu_summary.py\n" + txt)
    return txt


def raw(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(r.splitlines()))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


full = os.path.join(src, "tile_kernel_full.ncu-rep")
small = os.path.join(src, "tile_kernel_296x30.ncu-rep")
if os.path.exists(small):
    summary(small, os.path.join(dst, "%s_tile_kernel_296x30.txt" % tag))
if os.path.exists(full):
    summary(full, os.path.join(dst, "%s_tile_kernel_full.txt" % tag))
    v, u = raw(full)

    def gb(k):
        x = float(v[k])
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u[k]]
    n_streams, n_hops, hop = 4096, 2583, 1024
    hops = n_streams * n_hops
    inst = float(v["smsp__inst_executed.sum"])
    j = {
        "ncu": "profiles/%s_tile_kernel_full.txt (ncu --set full of one full-size launch, 4096 streams x 2583 hops)" % tag,
        "kernel": v.get("Kernel Name", "hpr_tile_fast_kernel"),
        "gpu_time_ms_under_ncu": float(v["gpu__time_duration.sum"]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[u["gpu__time_duration.sum"]],
        "dram_bytes_read": gb("dram__bytes_read.sum"),
        "dram_bytes_write": gb("dram__bytes_write.sum"),
        "dram_bytes_per_launch": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"),
        "algorithmic_bytes_per_launch": hops * hop * 4 * 2,
        "warp_instructions": inst,
        "warp_instr_per_hop": inst / hops,
        "issue_frac": float(v["smsp__issue_active.avg.pct_of_peak_sustained_active"]) / 100.0,
        "alu_pipe_frac": float(v["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]) / 100.0,
        "fma_pipe_frac": float(v["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]) / 100.0,
        "fma_pipe_inst_frac": float(v["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]) / 100.0,
        "lsu_wavefront_frac": float(v.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "nan")),
        "warps_active_frac": float(v["sm__warps_active.avg.pct_of_peak_sustained_active"]) / 100.0,
        "smem_bank_conflicts": float(v["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]),
        "registers_per_thread": int(v["launch__registers_per_thread"]),
        "grid": int(v["launch__grid_size"]), "block": int(v["launch__block_size"]),
    }
    stalls = {}
    for k in v:
        m = re.match(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio", k) or re.match(r"smsp__average_warp_latency_issue_stalled_(\w+)\.ratio", k)
        if m:
            try:
                stalls[m.group(1)] = float(v[k])
            except ValueError:
                pass
    if stalls:
        j["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    with open(os.path.join(dst, "tile_kernel_ncu.json"), "w") as f:
        json.dump(j, f, indent=1)
    print(json.dumps(j, indent=1))

# ---- launch list of a short bench run
ll = os.path.join(src, "launch_list_bench.csv")
if os.path.exists(ll):
    rows = [r for r in csv.reader(open(ll, errors="replace")) if len(r) > 5]
    hdr = next((r for r in rows if "Kernel Name" in r), None)
    agg = collections.OrderedDict()
    if hdr:
        ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        for r in rows:
            if r is hdr or len(r) <= vi or r[mi] != "gpu__time_duration.sum":
                continue
            t = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}[r[ui]]
            name = re.sub(r"\(.*", "", r[ki])
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += t
        tot = sum(a[1] for a in agg.values())
        with open(os.path.join(dst, "%s_launch_list_bench.txt" % tag), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none of `python bench.py --steps 2 --warmup 1 ...` (see tools/gpu_final.sh);\n"
                    "# per-launch times are cold-cache and serialised: the SHARE is what matters.  kernel, launches, total ms, share\n")
            for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write("%-90s %6d %12.3f ms %6.2f %%\n" % (name, n, t, 100.0 * t / tot))
        print(open(os.path.join(dst, "%s_launch_list_bench.txt" % tag)).read())

# ---- SASS opcode histogram of the shipped kernel (static, from the in-tree library)
so = os.path.join(ROOT, "zen_b200", "lib", "libzen_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, hist = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        hist.setdefault(cur, collections.Counter())[m.group(1)] += 1
with open(os.path.join(dst, "%s_sass_histogram.txt" % tag), "w") as f:
    f.write("# static SASS opcode histogram (cuobjdump -sass zen_b200/lib/libzen_b200.so), the kernels of the headline path\n")
    for name, h in hist.items():
        d = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        d = d.replace("(int)", "").replace("(bool)0", "false").replace("(bool)1", "true")
        d = re.sub(r"\(zen_b200::HprDev.*", "", d)
        if not re.search(r"hpr_tile_fast_kernel<4096|hpr_rt_kernel<4096, *\d+, *true|hpr_tile_kernel<4096", d):
            continue
        tot = sum(h.values())
        f.write("\n%s\n  %d instructions: " % (d, tot) + ", ".join("%s %d" % kv for kv in h.most_common(28)) + "\n")
        packed = sum(n for k, n in h.items() if k in ("FFMA2", "FADD2", "FMUL2"))
        f.write("  packed fp32x2 (FFMA2 + FADD2 + FMUL2): %d; scalar FFMA/FADD/FMUL: %d; FSET/FSETP: %d; LDS/STS: %d; BAR: %d\n"
                % (packed, sum(h[k] for k in ("FFMA", "FADD", "FMUL")), h["FSET"] + h["FSETP"], h["LDS"] + h["STS"], h["BAR"]))
print(open(os.path.join(dst, "%s_sass_histogram.txt" % tag)).read()[:3000])
