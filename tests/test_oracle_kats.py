"""CPU: known-answer tests of the oracle derived analytically from the cited reference formulas, one test per
Appendix-B quirk that is observable from outside, and the independent numpy cross-check of the restated
third-party arithmetic (FFT, LU inverse)."""
import numpy as np
import pytest

import beamform_b200 as bf
from beamform_b200.synth import mic_delays, synth_stream
from oracle_lib import Oracle
import np_restatement as npr

H = 512


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def test_sqrt_hann_cola_identity():
    # util.h:201-211 window used for analysis and synthesis: w^2[n] + w^2[n+N/2] == 1 (SURVEY B-18)
    for hop in (256, 512, 2048):
        w = Oracle(bf.make_config("das", mics="aira3", hop=hop)).window()
        assert np.allclose(w[:hop] ** 2 + w[hop:] ** 2, 1.0, atol=5e-16)
        assert w[0] == 0.0


def test_frequency_vector_quirks():
    o = Oracle(bf.make_config("das", mics="aira3"))
    f = o.freqs()
    N = 1024
    assert f[0] == 0.0 and f[1] == 48000 / N and f[N - 1] == -48000 / N
    assert f[N // 2 - 1] == 24000.0          # B-1: overwritten with sr/2 instead of (N/2-1)*sr/N
    assert f[N // 2] == 0.0                  # B-2: never written; pinned to 0.0
    assert f[N // 2 + 1] == -(N // 2 - 1) * 48000 / N


def test_delays_follow_reference_geometry():
    cfg = bf.make_config("das", mics="aira3", initial_angle=33.0)
    o = Oracle(cfg)
    assert np.allclose(o.delays(), mic_delays(bf.GEOMETRIES["aira3"], 33.0), atol=1e-18)
    assert o.delays()[0] == 0.0
    # B-6: polar coordinates come from the RAW positions even when mic 0 is off-origin
    xy = [(0.1, 0.05), (0.0, -0.18), (-0.156, -0.09)]
    o2 = Oracle(bf.make_config("das", mics=xy, initial_angle=0.0))
    assert np.allclose(o2.delays(), mic_delays(xy, 0.0), atol=1e-18)


def test_das_single_mic_is_identity_with_one_hop_latency():
    # M = 1: Y = X, so the WOLA chain must reproduce the input delayed by exactly one hop (util.h:301-302, B-22)
    cfg = bf.make_config("das", mics=[(0.0, 0.0)])
    rng = np.random.default_rng(3)
    x = (0.3 * rng.standard_normal((1, 20 * H))).astype(np.float32)
    y = Oracle(cfg).process(x)
    assert np.allclose(y[H:], x[0, :-H], atol=2e-7)
    assert np.allclose(y[:H], 0.0, atol=1e-12)


def test_das_on_axis_plane_wave_reproduces_mic0():
    # a source exactly in the look direction: every aligned channel equals mic 0, so DAS == mic 0 (one hop late).
    # Bin-centred tones; the sqrt-Hann window leaks each tone into neighbouring bins, whose steering phase differs
    # by 2*pi*(sr/N)*tau ~ 0.1 rad, so the identity holds to ~1e-3, not to rounding.
    xy = bf.GEOMETRIES["aira3"]
    theta, sr, N = 25.0, 48000, 1024
    tau = mic_delays(xy, theta)
    n = np.arange(24 * H)
    x = np.zeros((3, n.size))
    for k, a in ((32, 0.2), (77, 0.1), (300, 0.05)):
        f0 = k * sr / N
        for m in range(3):
            x[m] += a * np.sin(2 * np.pi * f0 * (n / sr - tau[m]) + 0.3 * k)
    x = x.astype(np.float32)
    y = Oracle(bf.make_config("das", mics="aira3", initial_angle=theta)).process(x)
    assert rel_l2(y[2 * H:], x[0, H:-H]) < 3e-3


def test_lcmv_without_interferers_is_mvdr_except_bin0():
    # K = 0 reduces lcmv to mvdr (lcmv.cpp:116 with C = d), and the only difference left is B-15: mvdr copies
    # X_0[0] through unscaled (mvdr.cpp:76-77) while lcmv treats bin 0 as out of band (lcmv.cpp:102, freq_min > 0).
    # The difference of the two outputs is therefore exactly the WOLA synthesis of the DC bin of microphone 0:
    # d_t[n] = a_t * w[n] with a_t = sum_n frame_t[n] w[n] / N.
    x = synth_stream(bf.GEOMETRIES["aira3"], 30 * H, seed=23)
    x = (x + 0.05).astype(np.float32)   # a DC offset makes bin 0 matter
    y_m = Oracle(bf.make_config("mvdr", mics="aira3", initial_angle=5.0)).process(x)
    y_l = Oracle(bf.make_config("lcmv", mics="aira3", initial_angle=5.0)).process(x)
    N = 2 * H
    w = np.sqrt(0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N))
    x0 = np.concatenate([np.zeros(H), x[0].astype(np.float64)])
    T = 30
    a = np.array([np.dot(x0[t * H:t * H + N], w) / N for t in range(T)])
    d = np.zeros(T * H)
    for t in range(T):
        d[t * H:(t + 1) * H] = a[t] * w[:H] + (a[t - 1] * w[H:] if t > 0 else 0.0)
    k = 3 * H   # the offset's leakage selects low bins on the first frames: cold-start NaNs (B-10) in both nodes alike
    assert np.array_equal(np.isnan(y_m), np.isnan(y_l)) and not np.isnan(y_m[k:]).any()
    assert rel_l2(y_m[k:].astype(np.float64) - y_l[k:], d[k:]) < 1e-5
    assert np.linalg.norm(d[k:]) > 0.1 * np.linalg.norm(y_l[k:])


def test_cold_start_selected_bin_gives_nan():
    # B-10: a bin that passes the gate on the very first frame inverts an all-zero covariance
    x = synth_stream(bf.GEOMETRIES["aira3"], 6 * H, seed=1, lead_in=0)
    y = Oracle(bf.make_config("mvdr", mics="aira3")).process(x)
    assert np.isnan(y[:2 * H]).any()


def test_gss_geometric_gradient_only_for_k0():
    # B-7: dJ2 carries the integer 1/(K+1): with lambda = 0 and the decorrelation term off (single source => E = 0)
    # W only moves when K == 0 and W A != I.  Steering a single mic pair: K = 0 keeps adapting, K = 1 is frozen at A^H.
    x = synth_stream(bf.GEOMETRIES["aira3"], 40 * H, seed=9)
    y0a = Oracle(bf.make_config("gss", mics="aira3", mu=0.001)).process(x)
    y0b = Oracle(bf.make_config("gss", mics="aira3", mu=0.002)).process(x)
    assert rel_l2(y0a, y0b) > 1e-6      # K = 0: the update (and therefore mu) matters


INTERF_SCRIPTS = [
    # (initial list, threshold, [(id, angle)...], expected final list)
    ([80.0, -60.0, 150.0], 1.0, [(2, -55.0)], [80.0, -55.0, 150.0]),                       # move
    ([80.0, -60.0, 150.0], 1.0, [(4, 120.0)], [80.0, -60.0, 150.0, 120.0]),                # id = K+1 adds
    ([80.0, -60.0, 150.0], 1.0, [(9, 120.0)], [80.0, -60.0, 150.0, 120.0]),                # id >> K adds too
    ([80.0, -60.0, 150.0], 1.0, [(4, 150.5)], [80.0, -60.0, 150.0]),                       # too close: not added
    ([80.0, -60.0, 150.0], 1.0, [(1, 149.5)], [-60.0, 150.0]),                             # moved onto a neighbour: removed
    ([80.0, -60.0, 150.0], 1.0, [(0, 10.0)], [80.0, -60.0, 150.0]),                        # id 0 invalid
    ([179.5], 1.0, [(2, -179.5)], [179.5, -179.5]),                                        # B-14: no wrap-around
    ([], 5.0, [(1, 30.0), (1, 31.0), (2, 33.0), (2, 40.0)], [31.0, 40.0]),
    ([10.0, 20.0], 5.0, [(2, 12.0)], [10.0]),
    ([10.0, 20.0, 30.0], 5.0, [(3, 11.0), (1, 21.0)], [20.0]),
]


@pytest.mark.parametrize("algo", ["lcmv", "gss"])
@pytest.mark.parametrize("initial,thr,script,expected", INTERF_SCRIPTS)
def test_interference_list_state_machine(algo, initial, thr, script, expected):
    # lcmv.cpp:258-309 == gss.cpp:288-339; float32 angles are widened to double (lcmv.cpp:262)
    o = Oracle(bf.make_config(algo, mics="aira3", interferers=tuple(initial), interf_angle_threshold=thr))
    for (i, a) in script:
        o.set_interference(i, a)
    assert o.interferences == [float(np.float32(a)) for a in expected]


def test_config_interferer_ingest_stops_at_first_out_of_range():
    # util.h:101-112 / B-14: |a| > 180 terminates the list; later entries are ignored
    o = Oracle(bf.make_config("lcmv", mics="aira3", interferers=(40.0, 181.0, -20.0)))
    assert o.interferences == [40.0]


def test_row0_zero_after_online_restructure():
    # B-8: allocate_interf_buffers zero-fills and update_weights(ini=false) skips row 0
    o = Oracle(bf.make_config("lcmv", mics="aira3", interferers=(80.0,)))
    assert np.all(o.weights()[:, 0, :] == 1.0)
    o.set_interference(2, -30.0)
    w = o.weights()
    assert w.shape[2] == 3 and np.all(w[:, 0, :] == 0.0) and np.all(np.abs(np.abs(w[:, 1:, :]) - 1.0) < 1e-12)


@pytest.mark.parametrize("algo,mics,kw", [("das", "aira3", {}), ("mvdr", "circ8", {}), ("lcmv", "circ8", dict(interferers=(80.0, -60.0, 150.0))),
                                          ("das", "binaural", dict(hop=2048)), ("mvdr", "aira3", dict(hop=256))])
def test_numpy_restatement_agrees(algo, mics, kw):
    # pocketfft + LAPACK against the oracle's own radix-2 FFT + LU: pins the restated third-party arithmetic
    hop = kw.get("hop", H)
    cfg = bf.make_config(algo, mics=mics, initial_angle=12.0, **kw)
    x = synth_stream(bf.GEOMETRIES[mics], 40 * hop, seed=17)
    y = Oracle(cfg).process(x)
    y_np = npr.run(algo, bf.GEOMETRIES[mics], x, hop=hop, theta=12.0, interferers=kw.get("interferers", ()))
    assert rel_l2(y, y_np) < 1e-7
