"""Print the key metrics of every kernel in an .ncu-rep (reads `ncu --page raw --csv`)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct", "dram__throughput.avg.pct", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__average_warp", "smsp__average_warps_issue_stalled"] + extra
idx = [i for i, h in enumerate(hdr) if any(h.startswith(w) for w in WANT)]
for r in rows[2:]:
    print("====", r[4][:90], "grid", r[hdr.index("Grid Size")] if "Grid Size" in hdr else "", "block", r[hdr.index("Block Size")] if "Block Size" in hdr else "")
    for i in idx:
        v = r[i]
        try:
            f = float(v.replace(",", ""))
            if "stall" in hdr[i] and f < 0.3:
                continue
            v = f"{f:,.3f}"
        except ValueError:
            pass
        print(f"   {hdr[i]:90s} {units[i]:12s} {v}")
