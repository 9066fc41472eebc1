#!/usr/bin/env python
"""Executed warp instructions per frame pair, by code region of sel_stream_kernel.cu (ncu source-page CSV).
usage: python tools/ncu_regions.py src.csv <pairs in the profiled launch>"""
import csv
import sys
from collections import defaultdict

rows = csv.reader(open(sys.argv[1]))
pairs = float(sys.argv[2])
cur = hdr = None
acc = defaultdict(float)
smp = defaultdict(float)


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed")
        isamp = hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) or cur is None:
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    acc[(cur, ln)] += num(r[ii])
    smp[(cur, ln)] += num(r[isamp])
print("total per pair %.0f" % (sum(acc.values()) / pairs))
f = "sel_stream_kernel.cu"
src = open(__file__.rsplit("/", 2)[0] + "/beamform_b200/csrc/" + f).read().splitlines()


def find(s):
    return next(i + 1 for i, l in enumerate(src) if s in l)


def rng(a, b, d=None):
    d = acc if d is None else d
    return sum(v for (ff, l), v in d.items() if ff == f and a <= l <= b) / (pairs if d is acc else 1)


marks = [("helpers", 1), ("ss_fft1024_fwd", find("ss_fft1024_fwd(float2 (&v)[32]")), ("ss_solve_batch", find("void ss_solve_batch(")), ("kernel setup", find("sel_stream_kernel(const __grid_constant__")),
         ("transform prologue", find("= transform warps")), ("mic load/window", find("hops -> windowed packed frame pair")), ("inv assemble G", find("output spectra of the pair -> G")),
         ("fft call", find("the transform (one code copy for both roles)")), ("inv OLA", find("synthesis window, overlap-add (util.h")), ("mic unpack", find("microphone warps only from here")),
         ("mic bar1", find("named_bar_sync(1,")), ("mic gate", find("gate (mvdr.cpp:79-85)")), ("mic bar2", find("named_bar_sync(2,")), ("mic park+capture", find("xpark[k2 * 32 + lane] = x1[k2]") - 1),
         ("mic staging", find("unsigned emask = 0;")), ("epilogue", find("end of the stream: one more arrival per slot") - 1), ("solver loop", find("= solver warps")), ("kernel end", find("tcgen05.fence::before_thread_sync") + 20), ("end", len(src) + 1)]
marks = sorted(marks, key=lambda x: x[1])
tot_s = sum(smp.values())
for (n, a), (_, b) in zip(marks, marks[1:]):
    print("%-22s lines %3d-%3d: %6.0f instr/pair  %5.1f%% samples" % (n, a, b - 1, rng(a, b - 1), 100 * rng(a, b - 1, smp) / tot_s))
for ff in sorted(set(k[0] for k in acc) - {f}):
    print("%-28s %6.0f instr/pair  %5.1f%% samples" % (ff, sum(v for (f2, l), v in acc.items() if f2 == ff) / pairs, 100 * sum(v for (f2, l), v in smp.items() if f2 == ff) / tot_s))
