import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import beamform_b200 as bf
from beamform_b200.synth import synth_batch
from oracle_lib import Oracle
for algo, kw in (("mvdr", {}), ("lcmv", dict(interferers=(80.0, -60.0, 150.0)))):
    cfg = bf.make_config(algo, mics="circ8", **kw)
    x = synth_batch(bf.GEOMETRIES["circ8"], 4, 41 * 512, seed=123)
    t0 = time.time()
    got = bf.Beamformer(cfg, n_streams=4).process(x)
    ref = np.stack([Oracle(cfg).process(x[b]) for b in range(4)])
    ok = np.isfinite(ref)
    print(algo, "finite match", np.array_equal(np.isfinite(got), ok), "rel", np.linalg.norm(got[ok] - ref[ok]) / np.linalg.norm(ref[ok]), "t", time.time() - t0, flush=True)
