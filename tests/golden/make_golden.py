#!/usr/bin/env python
"""Generates tests/golden/ref_outputs.npz by running the reference's UNMODIFIED node sources
(oracle/_ref/<node>_ref, built from /root/reference by oracle/Makefile.ref against oracle/shim/) on the
seeded inputs of cases.py.  Run in the build container (where /root/reference exists):

    make -C oracle -f Makefile.ref && python tests/golden/make_golden.py

The archive holds, per case, the reference output (float32), the final interference list and a checksum of
the input (so a drifting generator is caught, not silently compared)."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from golden.cases import CASES, build_case  # noqa: E402
from ref_lib import run_ref  # noqa: E402


def main():
    out = {}
    for name in CASES:
        cfg, x, events = build_case(name)
        y, interf = run_ref(CASES[name]["algo"], cfg, x, events=events, want_interf=True)
        out[name + "/out"] = y
        out[name + "/interf"] = np.asarray(interf, dtype=np.float64)
        out[name + "/in_sha256"] = np.frombuffer(hashlib.sha256(x.tobytes()).digest(), dtype=np.uint8)
        print("%-32s hops=%4d  nan=%d  rms=%.4g" % (name, CASES[name]["hops"], int(np.isnan(y).sum()), float(np.sqrt(np.nanmean(y.astype(np.float64) ** 2)))))
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)


if __name__ == "__main__":
    main()
