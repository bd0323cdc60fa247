"""Static counts of the Blackwell-specific SASS instructions per kernel of the built library (markdown table on stdout).
usage: python tools/sass_evidence.py [regex of kernel names to keep]"""
import collections
import re
import subprocess
import sys

LIB = "linnaeus_b200/liblinnaeus_b200.so"
keep = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cols = ["UTCHMMA", "HMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "FFMA2", "MUFU.TANH", "MUFU.EX2", "MUFU.COS"]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name)
        cur = counts.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    for c in cols:
        if re.search(r"(?<![A-Z])" + re.escape(c) + r"\b", line):
            cur[c] += 1
print("| kernel | " + " | ".join(cols) + " |\n|---|" + "---|" * len(cols))
for name, c in sorted(counts.items()):
    if keep and not keep.search(name):
        continue
    if not any(c[k] for k in cols):
        continue
    print(f"| `{name}` | " + " | ".join(str(c[k]) for k in cols) + " |")
