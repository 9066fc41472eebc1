#!/usr/bin/env python
"""Per source line: samples by stall reason, from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_stall_lines.py [file-substring] [topN]"""
import csv
import sys
from collections import defaultdict

pat = sys.argv[1] if len(sys.argv) > 1 else ""
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = csv.reader(sys.stdin)
cur = None
hdr = None
acc = defaultdict(lambda: defaultdict(float))
tot = defaultdict(float)
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        idx = {n: i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n}
        i_s = hdr.index("# Samples")
        i_i = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or cur is None or pat not in cur:
        continue
    try:
        key = (cur, int(r[0]))
    except ValueError:
        continue
    for n, i in idx.items():
        try:
            v = float(r[i] or 0)
        except ValueError:
            v = 0
        acc[key][n] += v
        tot[n] += v
    def num(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    acc[key]["#"] += num(r[i_s])
    acc[key]["instr"] += num(r[i_i])
alls = sum(a["#"] for a in acc.values())
print("samples", alls, {k: round(100 * v / max(alls, 1), 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]})
for key, a in sorted(acc.items(), key=lambda kv: -kv[1]["#"])[:top]:
    reasons = sorted(((v, n) for n, v in a.items() if n.startswith("stall_")), reverse=True)[:3]
    print("%5.1f%% %-26s instr %10d  %s" % (100 * a["#"] / max(alls, 1), "%s:%d" % key, a["instr"], ", ".join("%s %.0f%%" % (n[6:], 100 * v / max(a["#"], 1)) for v, n in reasons)))
