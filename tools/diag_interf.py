import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import beamform_b200 as bf
from beamform_b200.synth import synth_stream
from oracle_lib import Oracle
H = 512
for algo, mics, K, kw in (("lcmv", "circ16", 8, dict(past_windows=40, freq_min=3000, freq_max=9000)), ("lcmv", "circ16", 9, dict(past_windows=40, freq_min=4000, freq_max=9000)),
                          ("lcmv", "circ16", 10, dict(past_windows=40, freq_min=5000, freq_max=9000)), ("lcmv", "circ16", 12, dict(past_windows=40, freq_min=6000, freq_max=9000)),
                          ("lcmv", "circ16", 8, dict(past_windows=10, freq_min=3000, freq_max=9000))):
    interf = tuple(-170.0 + 340.0 / K * k for k in range(K))
    cfg = bf.make_config(algo, mics=mics, initial_angle=10.0, interferers=interf, **kw)
    x = np.stack([synth_stream(bf.GEOMETRIES[mics], 61 * H, seed=900 + b, sources=((20.0, 0.1, 180.0, 60), (-70.0, 0.05, 233.0, 60))) for b in range(2)])
    ref, sel, _ = zip(*[Oracle(cfg).process(x[b], want_flags=True) for b in range(2)])
    ref = np.stack(ref)
    got = bf.Beamformer(cfg, n_streams=2).process(x)
    ok = np.isfinite(ref) & np.isfinite(got)
    print(algo, mics, K, kw, "finite eq", np.array_equal(np.isfinite(ref), np.isfinite(got)), "rel", np.linalg.norm(got[ok] - ref[ok]) / np.linalg.norm(ref[ok]), "ref rms", np.sqrt(np.mean(ref[ok] ** 2)), "selected", int(np.stack(sel).sum()), flush=True)
