#!/usr/bin/env python
"""Latency of the drop-in callback: wall time of one bf_process_hop call (host buffers in, host buffers out: H2D, kernels,
D2H, synchronisation) per node, against the JACK period it has to fit in.  usage: python tools/hop_latency.py [out.json]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import beamform_b200 as bf  # noqa: E402
from beamform_b200.synth import synth_stream  # noqa: E402

CASES = [("das", "aira3", 512, {}), ("mvdr", "circ8", 512, {}), ("lcmv", "circ8", 512, dict(interferers=(80.0, -60.0, 150.0))),
         ("gss", "circ8", 512, dict(interferers=(80.0, -60.0, 150.0))), ("phase", "aira3", 512, {}), ("phasempf", "binaural", 512, {}),
         ("phasempf", "binaural", 2048, {}), ("mcra", "aira3", 512, {}), ("gsc", "aira3", 512, {}), ("ref", "aira3", 512, {})]
out = []
for algo, mics, hop, kw in CASES:
    cfg = bf.make_config(algo, mics=mics, hop=hop, **kw)
    b = bf.Beamformer(cfg, n_streams=1)
    n = 300
    x = synth_stream(bf.GEOMETRIES[mics], n * hop, seed=1)
    for t in range(50):
        b.process_hop(x[:, t * hop:(t + 1) * hop])
    ts = []
    for t in range(50, n):
        t0 = time.perf_counter()
        b.process_hop(x[:, t * hop:(t + 1) * hop])
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    rec = {"algo": algo, "mics": len(bf.GEOMETRIES[mics]), "hop": hop, "period_us": 1e6 * hop / 48000.0, "median_us": float(np.median(ts)),
           "p99_us": float(np.percentile(ts, 99)), "max_us": float(ts.max())}
    out.append(rec)
    print(json.dumps(rec))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
