"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into a per-kernel share table (markdown)."""
import csv
import re
import sys
from collections import defaultdict

path, first, last = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e3))
rows = [r for r in rows if first <= r[0] < last]
agg = defaultdict(lambda: [0, 0.0])
for _, name, us in rows:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("<unnamed>::", "").replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(f"launches {len(rows)} (IDs {first}..{last}), total {tot:.0f} us\n")
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"| `{name[:95]}` | {n} | {us:.0f} | {100 * us / tot:.1f}% | {us / n:.1f} |")
