"""Summarise one ncu report of a warp-specialised kernel: headline metrics and stall samples grouped by how often each SASS
instruction executed (= which warp role it belongs to).  python tools/ncu_roles.py report.ncu-rep"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_pipe_fp16.sum", "sm__inst_executed_pipe_fp16.avg.pct_of_peak_sustained_active"]
for h, v in zip(hdr, vals):
    if h in want:
        print(f"{h:80s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
groups = {}
for r in rows[2:]:
    ex = int(r[ix["Instructions Executed"]] or 0)
    n = int(r[ix["# Samples"]] or 0)
    g = groups.setdefault(ex, [0, 0, Counter()])
    g[0] += n
    g[1] += 1
    for h in stall:
        g[2][h] += int(r[ix[h]] or 0)
tot = sum(g[0] for g in groups.values())
print("total samples", tot)
for ex, (n, cnt, c) in sorted(groups.items(), key=lambda kv: -kv[1][0])[:12]:
    print(f"exec={ex:9d} static instrs={cnt:4d} samples={n:6d} ({100 * n / tot:4.1f}%)", [(k[6:], v) for k, v in c.most_common(6)])
