#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source file:line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_src_summary.py [topN]"""
import csv
import sys
from collections import defaultdict

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
cur_file = None
by_line = defaultdict(lambda: [0, 0, 0, ""])   # samples, instr executed, thread instr
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) == 2:
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_line, i_src = 0, 1
        i_samp = hdr.index("# Samples")
        i_inst = hdr.index("Instructions Executed")
        i_thr = hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        key = (cur_file, int(r[i_line]))
    except ValueError:
        continue
    e = by_line[key]
    def num(v):
        try:
            return int(float(v))
        except ValueError:   # "", "-", or a source line whose inline asm broke the CSV columns
            return 0
    e[0] += num(r[i_samp])
    e[1] += num(r[i_inst])
    e[2] += num(r[i_thr])
    e[3] = r[i_src][:110]
tot_s = sum(e[0] for e in by_line.values()) or 1
tot_i = sum(e[1] for e in by_line.values()) or 1
print("total samples %d, total warp instr %d" % (tot_s, tot_i))
byfile = defaultdict(lambda: [0, 0])
for (f, l), e in by_line.items():
    byfile[f][0] += e[0]
    byfile[f][1] += e[1]
for f, e in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("%-28s samples %5.1f%%  instr %5.1f%%" % (f, 100.0 * e[0] / tot_s, 100.0 * e[1] / tot_i))
print()
for (f, l), e in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% ins  lanes/instr %4.1f  %s:%d  %s" % (100.0 * e[0] / tot_s, 100.0 * e[1] / tot_i, e[2] / max(1, e[1]), f, l, e[3].strip()))
