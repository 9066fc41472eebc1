#!/usr/bin/env python
"""Closed-loop steering from the steered-response map.

The reference ships scripts (scripts/energy2theta.py, energy2theta-diff.py, SIR2theta.py) that search the look direction by gradient steps on
the beamformer's output energy and publish it on /theta.  On the GPU the search is a sweep: the steered-response power of every candidate
direction comes out of one call (bf_srp_batch_device, config C5), its arg-max is the direction, and bf_set_theta applies it at the next hop
boundary -- the same topic, the same convention (0 front, -90 left, 90 right, 180 back; README.md:21).

    follow(cfg, x, block_hops=8, thetas=...) -> (beamformed [L], theta per block)

`cfg` is the beamformer's bf_config (any node); the sweep uses a das handle on the same geometry."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import beamform_b200 as bf  # noqa: E402


def follow(cfg, x, block_hops=8, thetas=None, smooth=0.5):
    """x [M][L] float32 (host).  Every `block_hops` hops: sweep -> arg-max of the (exponentially smoothed) map -> set_theta -> beamform the block."""
    import torch
    M, L = x.shape
    H = cfg.hop
    T = L // H
    thetas = np.arange(-180.0, 180.0, 2.0, dtype=np.float32) if thetas is None else np.asarray(thetas, dtype=np.float32)
    sweep_cfg = bf.BfConfig.from_buffer_copy(cfg)
    sweep_cfg.algo = bf.ALGOS["das"]
    sweeper = bf.Beamformer(sweep_cfg, n_streams=1)
    beam = bf.Beamformer(cfg, n_streams=1)
    xd = torch.from_numpy(np.ascontiguousarray(x[None], dtype=np.float32)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    out, track, avg = [], [], None
    for t0 in range(0, T, block_hops):
        n = min(block_hops, T - t0)
        blk = xd[:, :, t0 * H:(t0 + n) * H].contiguous()
        maps = torch.empty((1, n, len(thetas)), dtype=torch.float32, device="cuda")
        sweeper.srp_device(blk.data_ptr(), thetas, maps.data_ptr(), n, stream_ptr=st)
        m = maps[0].sum(dim=0).cpu().numpy().astype(np.float64)
        avg = m if avg is None else smooth * avg + (1.0 - smooth) * m
        theta = float(thetas[int(np.argmax(avg))])
        beam.set_theta(theta)                      # takes effect at the next hop boundary, like a /theta message
        track.append(theta)
        y = torch.empty((1, n * H), dtype=torch.float32, device="cuda")
        beam.process_device(blk.data_ptr(), y.data_ptr(), n, stream_ptr=st)
        out.append(y[0].cpu().numpy())
    return np.concatenate(out), np.asarray(track)


if __name__ == "__main__":
    from beamform_b200.synth import synth_stream
    xy = bf.GEOMETRIES["circ8"]
    a = synth_stream(xy, 64 * 512, sources=((30.0, 0.1, 190.0, 30),), lead_in=0, seed=1)
    b = synth_stream(xy, 64 * 512, sources=((-70.0, 0.1, 190.0, 30),), lead_in=0, seed=2)
    y, track = follow(bf.make_config("das", mics="circ8"), np.concatenate([a, b], axis=1))
    print("look direction per block:", track)
