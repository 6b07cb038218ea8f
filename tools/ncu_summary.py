"""Summarise an .ncu-rep (raw page + SASS hot regions) as text; run where ncu is installed (no GPU needed)."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 400
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
for k in want:
    if k in hdr:
        i = hdr.index(k)
        print("%-70s %s %s" % (k, vals[i], units[i]))
for i, h in enumerate(hdr):
    if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct"):
        try:
            if float(vals[i]) >= 3.0:
                print("%-70s %s" % (h.replace("smsp__warp_issue_stalled_", "stall "), vals[i]))
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
ci, si, sa = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
data = [(int(r[ci]), int(r[sa]), r[si].strip()) for r in rows[2:] if len(r) > ci and r[ci].isdigit()]
tot = sum(d[0] for d in data)
ts = sum(d[1] for d in data)
print("SASS instructions:", len(data), "executed warp-instructions:", tot)
for b in range(0, len(data), B):
    seg = data[b:b + B]
    n = sum(d[0] for d in seg)
    s = sum(d[1] for d in seg)
    if n < tot * 0.004:
        continue
    ops = collections.Counter()
    for d in seg:
        op = d[2].split()[0] if not d[2].startswith("@") else d[2].split()[1]
        ops[op.split(".")[0]] += d[0]
    print("sass %5d-%5d instr %5.1f%% stall-samples %5.1f%%  %s" % (
        b, b + B, 100 * n / tot, 100 * s / ts, ", ".join("%s %.0f%%" % (o, 100 * c / max(n, 1)) for o, c in ops.most_common(6))))
