"""Turn `ncu -i X.ncu-rep --page raw --csv` (one row per profiled launch) into a compact markdown table:
duration, DRAM read / write bytes and throughput, registers, achieved occupancy, issue-slot, FMA / tensor pipe
utilisation.  Usage: python tools/ncu_table.py raw.csv [labels.txt] > table.md

labels.txt (optional): one label per profiled launch, in launch order (tools/prof_kernels.py prints them)."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr, units, data = rows[0], rows[1], rows[2:]
labels = [l.strip() for l in open(sys.argv[2])] if len(sys.argv) > 2 else []


def col(suffix):
    """Column of a metric: exact name first, else a section-prefixed copy that has data in the first row."""
    exact = [i for i, h in enumerate(hdr) if h == suffix]
    pref = [i for i, h in enumerate(hdr) if h != suffix and h.endswith(suffix)]
    for i in exact + pref:
        if data and data[0][i] not in ("", "no data", "n/a"):
            return i
    return (exact + pref + [None])[0]


def scale(i, v):
    """ncu prints each column in its own unit; normalise times to us and bytes to MB."""
    u = units[i].lower()
    f = float(v.replace(",", ""))
    if u in ("ns", "nsecond"):
        return f / 1e3
    if u in ("us", "usecond"):
        return f
    if u in ("ms", "msecond"):
        return f * 1e3
    if u in ("s", "second"):
        return f * 1e6
    if u == "byte":
        return f / 1e6
    if u == "kbyte":
        return f / 1e3
    if u == "mbyte":
        return f
    if u == "gbyte":
        return f * 1e3
    return f


C = {
    "t": col("gpu__time_duration.sum"),
    "rd": col("dram__bytes_read.sum"),
    "wr": col("dram__bytes_write.sum"),
    "regs": col("launch__registers_per_thread"),
    "occ": col("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "issue": col("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fma": col("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "tensor": col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    "dram": col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "lts": col("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    "grid": col("Grid Size"),
    "block": col("Block Size"),
    "name": col("Kernel Name"),
}


def get(r, k, fmt="{:.1f}"):
    i = C[k]
    if i is None or r[i] == "":
        return "-"
    try:
        return fmt.format(scale(i, r[i]))
    except ValueError:
        return r[i]


print("| # | kernel | what | grid x block | regs | us | DRAM rd MB | DRAM wr MB | DRAM GB/s | dram % | L2 % | occ % | issue % | fma % | tensor % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for n, r in enumerate(data):
    name = re.sub(r"\(.*", "", r[C["name"]]).replace("<unnamed>::", "").replace("void ", "")
    t = scale(C["t"], r[C["t"]])
    rd, wr = scale(C["rd"], r[C["rd"]]), scale(C["wr"], r[C["wr"]])
    gbs = (rd + wr) / t * 1e3 if t > 0 else 0.0  # MB / us = TB/s -> GB/s
    grid = r[C["grid"]].replace(" ", "") if C["grid"] is not None else ""
    block = r[C["block"]].replace(" ", "") if C["block"] is not None else ""
    lab = labels[n] if n < len(labels) else ""
    print(f"| {n} | `{name[:60]}` | {lab} | {grid} x {block} | {get(r, 'regs', '{:.0f}')} | {t:.1f} | {rd:.1f} | {wr:.1f} | {gbs:.0f} | "
          f"{get(r, 'dram')} | {get(r, 'lts')} | {get(r, 'occ')} | {get(r, 'issue')} | {get(r, 'fma')} | {get(r, 'tensor')} |")
