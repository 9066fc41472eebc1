"""CPU: the oracle (oracle/bf_oracle.hpp) against golden vectors produced by the reference's own,
unmodified node sources (tests/golden/make_golden.py -> oracle/_ref/<node>_ref).  When the _ref binaries
are present (build container) they are also run live on a fresh input."""
import hashlib
import os

import numpy as np
import pytest

import beamform_b200 as bf
from beamform_b200.synth import synth_stream
from golden.cases import CASES, build_case
from oracle_lib import Oracle
import ref_lib

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_outputs.npz"))
# The oracle and the compiled reference differ only in the association order inside their (restated) FFT and
# small-matrix products: agreement is at double-rounding level, far below one float32 ulp of the output.
TOL = 1e-8


def rel_l2_finite(a, b):
    ok = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), ok), "NaN/inf samples must coincide (cold-start quirk B-10)"
    return float(np.linalg.norm(a[ok].astype(np.float64) - b[ok]) / max(np.linalg.norm(b[ok].astype(np.float64)), 1e-300))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
    cfg, x, events = build_case(name)
    assert hashlib.sha256(x.tobytes()).digest() == GOLD[name + "/in_sha256"].tobytes(), "synthetic generator drifted"
    o = Oracle(cfg)
    y = o.process(x, events=events)
    ref = GOLD[name + "/out"]
    assert y.shape == ref.shape
    assert rel_l2_finite(y, ref) <= TOL
    assert o.interferences == list(GOLD[name + "/interf"]), "interference list must be bit-exact"


def test_coldstart_case_really_contains_nans():
    ref = GOLD["mvdr_aira3_coldstart_nan/out"]
    assert np.isnan(ref).sum() >= 512


@pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("algo,mics,kw", [("das", "circ8", {}), ("mvdr", "aira3", {}), ("lcmv", "circ8", dict(interferers=(45.0, -120.0))),
                                          ("gss", "circ8", dict(interferers=(45.0,))), ("phase", "binaural", dict(mag_threshold=0.003)),
                                          ("phasempf", "aira3", dict(out_only_noise=True, smooth_size=20))])
def test_oracle_matches_live_reference_build(algo, mics, kw):
    cfg = bf.make_config(algo, mics=mics, initial_angle=-12.5, **kw)
    x = synth_stream(bf.GEOMETRIES[mics], 70 * 512, seed=977, gate_hz=1.3 if algo == "phasempf" else 0.0)
    events = [(11, "theta", 33.0), (30, "interf", 1, 50.0), (31, "interf", 9, -170.0)]
    y = Oracle(cfg).process(x, events=events)
    ref, interf = ref_lib.run_ref(algo, cfg, x, events=events, want_interf=True)
    assert rel_l2_finite(y, ref) <= TOL
    o = Oracle(cfg)
    o.process(x, events=events)
    assert o.interferences == interf
