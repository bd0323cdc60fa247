"""Summarise an `ncu --page source --csv` dump: instruction mix by opcode and the hottest SASS regions.
usage: python tools/ncu_source_hot.py file.csv [n_regions]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
S, IE, WS = ci["Source"], ci["Instructions Executed"], ci["Warp Stall Sampling (All Samples)"]
agg, st = collections.Counter(), collections.Counter()
tot = tots = 0
seq = []
for r in rows[2:]:
    try:
        n, s = int(r[IE]), int(r[WS])
    except (ValueError, IndexError):
        continue
    toks = r[S].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    agg[op] += n
    st[op] += s
    tot += n
    tots += s
    seq.append((n, s, r[S].strip()))
print(f"total warp instructions {tot}, stall samples {tots}")
for k, v in agg.most_common(22):
    print(f"  {k:10s} inst {v:>12d} ({100 * v / tot:5.1f}%)  stalls {st[k]:>7d} ({100 * st[k] / max(tots, 1):5.1f}%)")
# regions: runs of instructions with the same execution count
print("regions (consecutive instructions, same execution count):")
i = 0
regs = []
while i < len(seq):
    j = i
    while j < len(seq) and seq[j][0] == seq[i][0]:
        j += 1
    regs.append((i, j, seq[i][0], sum(s[1] for s in seq[i:j])))
    i = j
for a, b, n, s in sorted(regs, key=lambda r: -r[3])[: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print(f"  [{a:5d},{b:5d}) x{n:>9d} inst {n * (b - a):>12d} stalls {s:>7d}   {seq[a][2][:60]}")
