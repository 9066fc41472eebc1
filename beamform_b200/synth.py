"""Seeded synthetic multichannel signals for parity tests and the bench.

Far-field plane waves generated with the reference's own delay model (util.h:136-161):
x_m[n] = sum_s a_s * sum_h (1/h) sin(2 pi h f0_s (n/sr - tau_m(theta_s)) + phi_{s,h}) + sigma * N(0,1),
cast to float32.  The first `lead_in` samples are noise-only so adaptive histories are non-zero
but below the magnitude gate (avoids the cold-start NaNs of SURVEY B-10 unless a test asks for them).
"""
import numpy as np

V_SOUND = 343.0


def mic_delays(mic_xy, theta_deg):
    """tau_i of util.h:136-161: polar coordinates from RAW x,y, mic 0 forced to 0."""
    xy = np.asarray(mic_xy, dtype=np.float64)
    dist = np.sqrt(xy[:, 0] ** 2 + xy[:, 1] ** 2)
    ang = np.arctan2(xy[:, 1], xy[:, 0]) * 180.0 / np.pi
    d = ang - theta_deg
    d = np.where(d > 180, d - 360, np.where(d < -180, d + 360, d))
    tau = dist * np.cos(d * np.pi / 180.0) / (-V_SOUND)
    tau[0] = 0.0
    return tau


def synth_stream(mic_xy, n_samples, sr=48000, sources=((20.0, 0.1, 180.0, 24), (-70.0, 0.05, 233.0, 24)), sigma=1e-3,
                 seed=0xBEA4F0, lead_in=2048, gate_hz=0.0):
    """[M][n_samples] float32.  sources: (theta_deg, amplitude, f0_hz, n_harmonics).  gate_hz > 0 switches the
    sources on/off with that rate (speech pauses for the MCRA of phasempf)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    M = len(mic_xy)
    n = np.arange(n_samples, dtype=np.float64)
    x = sigma * rng.standard_normal((M, n_samples))
    env = np.ones(n_samples)
    env[:lead_in] = 0.0
    if gate_hz > 0:
        env *= (np.sin(2 * np.pi * gate_hz * n / sr) > -0.3).astype(np.float64)
    for (theta, amp, f0, nh) in sources:
        tau = mic_delays(mic_xy, theta)
        phi = rng.uniform(0, 2 * np.pi, size=nh)
        for m in range(M):
            t = n / sr - tau[m]
            s = np.zeros(n_samples)
            for h in range(1, nh + 1):
                if h * f0 < sr / 2:
                    s += (1.0 / h) * np.sin(2 * np.pi * h * f0 * t + phi[h - 1])
            x[m] += amp * env * s
    return x.astype(np.float32)


def synth_batch(mic_xy, n_streams, n_samples, sr=48000, seed=0xBEA4F0, randomize_theta=True, **kw):
    """[B][M][n_samples] float32, stream b seeded with seed + b (sources at per-stream random directions)."""
    out = np.empty((n_streams, len(mic_xy), n_samples), dtype=np.float32)
    for b in range(n_streams):
        rng = np.random.Generator(np.random.PCG64(seed + 7919 * (b + 1)))
        if randomize_theta:
            srcs = ((float(rng.uniform(-40, 40)), 0.1, float(rng.uniform(120, 260)), 24),
                    (float(rng.uniform(60, 300)) - 180.0 * 0, 0.05, float(rng.uniform(120, 260)), 24))
            kw2 = dict(kw)
            kw2["sources"] = srcs
        else:
            kw2 = kw
        out[b] = synth_stream(mic_xy, n_samples, sr=sr, seed=seed + b, **kw2)
    return out
