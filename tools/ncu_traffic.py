#!/usr/bin/env python
"""profiles/traffic_<workload>.json from an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`
log of one bench.py run: DRAM bytes per launch of the dominant (longest) kernel, averaged over its launches.
usage: python tools/ncu_traffic.py <ncu.csv> <workload> [out_dir]"""
import csv
import json
import os
import sys
from collections import defaultdict

path, wl = sys.argv[1], sys.argv[2]
out_dir = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = rows[0]
iK, iM, iU, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
iID = hdr.index("ID")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
per = defaultdict(dict)
for r in rows[1:]:
    if not any(k in r[iK] for k in ("bf::", "das_pairs_kernel", "sel_pairs_kernel", "sel_stream_kernel", "frames_kernel", "phase_n_kernel", "mcra_pairs_kernel", "srp_", "save_prev_hop", "gss_reset", "zero_hops", "ref_kernel", "gsc_")):
        continue
    per[(r[iID], r[iK])][r[iM]] = float(r[iV].replace(",", "")) * scale.get(r[iU], 1.0)
# bench.py also launches a small selection-density probe (16 streams): only launches of at least half the kernel's longest
# duration count as "the bench shape"
tmax = defaultdict(float)
for (_, k), m in per.items():
    tmax[k] = max(tmax[k], m.get("gpu__time_duration.sum", 0.0))
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for (_, k), m in per.items():
    if m.get("gpu__time_duration.sum", 0.0) < 0.5 * tmax[k]:
        continue
    a = agg[k]
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("dram__bytes_read.sum", 0.0)
    a[3] += m.get("dram__bytes_write.sum", 0.0)
if not agg:
    sys.exit("no bf:: kernels in " + path)
k, a = max(agg.items(), key=lambda kv: kv[1][1])
out = {"workload": wl, "kernel": k, "launches": a[0], "dram_bytes_read_per_launch": a[2] / a[0], "dram_bytes_write_per_launch": a[3] / a[0],
       "dram_bytes_per_launch": (a[2] + a[3]) / a[0], "ncu_time_ms_per_launch": 1e3 * a[1] / a[0],
       "share_of_bf_kernel_time": a[1] / sum(v[1] for v in agg.values()),
       "dram_bytes_all_kernels_per_step": sum(v[2] + v[3] for v in agg.values()) / a[0],
       "kernels": {kk: {"launches": v[0], "ms_per_launch": 1e3 * v[1] / v[0], "dram_bytes_per_launch": (v[2] + v[3]) / v[0]} for kk, v in agg.items()},
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none (bench shape)"}
json.dump(out, open(os.path.join(out_dir, "traffic_%s.json" % wl), "w"), indent=1)
print(json.dumps(out))
