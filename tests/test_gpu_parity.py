"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerance (BASELINE.json north_star): float outputs within relative L2 error <= 1e-4 of the oracle
(>= 80 dB SNR); selected-bin sets and interference lists bit-exact.
"""
import numpy as np
import pytest

import beamform_b200 as bf
from beamform_b200.synth import synth_batch, synth_stream
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu
REL_L2_TOL = 1e-4
MASK_MISMATCH_FRAC = 2e-5   # phase-mask decisions (not on the north_star's bit-exact list): at most 1 in 50 000 may differ from the oracle
H = 512


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def oracle_batch(cfg, x, events=()):
    return np.stack([Oracle(cfg).process(x[b], events=events) for b in range(x.shape[0])])


@pytest.mark.parametrize("mics,theta,n_hops", [("aira3", 0.0, 41), ("aira3", 20.0, 40), ("circ8", -35.0, 17), ("binaural", 90.0, 8)])
def test_das_matches_oracle(mics, theta, n_hops):
    cfg = bf.make_config("das", mics=mics, initial_angle=theta)
    x = synth_batch(bf.GEOMETRIES[mics], 3, n_hops * H, seed=11)
    ref = oracle_batch(cfg, x)
    got = bf.Beamformer(cfg, n_streams=3).process(x)
    assert got.shape == ref.shape
    assert rel_l2(got, ref) <= REL_L2_TOL
    for b in range(3):
        assert rel_l2(got[b], ref[b]) <= REL_L2_TOL


def test_das_split_calls_and_theta_events_match_single_call():
    cfg = bf.make_config("das", mics="aira3", initial_angle=0.0)
    x = synth_batch(bf.GEOMETRIES["aira3"], 2, 30 * H, seed=5)
    events = [(7, "theta", 20.0), (8, "theta", -45.0), (20, "theta", 110.0)]
    ref = oracle_batch(cfg, x, events=events)
    one = bf.Beamformer(cfg, n_streams=2).process(x, events=events)
    assert rel_l2(one, ref) <= REL_L2_TOL
    # same stream fed in three calls with the setter API instead of scheduled events
    b = bf.Beamformer(cfg, n_streams=2)
    parts = [b.process(x[:, :, :7 * H])]
    b.set_theta(20.0)
    parts.append(b.process(x[:, :, 7 * H:8 * H]))
    b.set_theta(-45.0)
    parts.append(b.process(x[:, :, 8 * H:20 * H]))
    b.set_theta(110.0)
    parts.append(b.process(x[:, :, 20 * H:]))
    assert rel_l2(np.concatenate(parts, axis=1), ref) <= REL_L2_TOL


def test_das_hop_at_a_time_callback_matches_batch():
    cfg = bf.make_config("das", mics="aira3", initial_angle=20.0)
    x = synth_stream(bf.GEOMETRIES["aira3"], 12 * H, seed=3)
    ref = Oracle(cfg).process(x)
    b = bf.Beamformer(cfg, n_streams=1)
    got = np.concatenate([b.process_hop(x[:, t * H:(t + 1) * H]) for t in range(12)])
    assert rel_l2(got, ref) <= REL_L2_TOL


def test_das_streams_cut_across_workers_keep_their_overlap_add_tails():
    """das_pairs_kernel cuts the stream-major pair sequence into equal ranges per warp (148 x 8 workers): with many more pairs
    than workers most streams have their first and their last pair in different warps, and the second call starts from non-zero
    OLA tails (util.h:301-302).  Every stream is checked against the oracle, first hop of each call included."""
    cfg = bf.make_config("das", mics="aira3", initial_angle=20.0)
    B, T = 1500, 24
    x = synth_batch(bf.GEOMETRIES["aira3"], 6, 2 * T * H, seed=77)
    x = np.ascontiguousarray(np.tile(x, (B // 6, 1, 1)) * np.linspace(0.5, 1.5, B, dtype=np.float32)[:, None, None])   # DAS is linear: distinct streams cheaply
    ref6 = oracle_batch(cfg, x[:6] / np.linspace(0.5, 1.5, B, dtype=np.float32)[:6, None, None])
    b = bf.Beamformer(cfg, n_streams=B)
    got = np.concatenate([b.process(x[:, :, :T * H]), b.process(x[:, :, T * H:])], axis=1)
    for blk in (0, 1, 7, 100, 249):
        spot = oracle_batch(cfg, x[6 * blk:6 * blk + 6])
        for i in range(6):
            assert rel_l2(got[6 * blk + i], spot[i]) <= REL_L2_TOL
    scale = np.linspace(0.5, 1.5, B)[:, None]
    ref = np.tile(ref6, (B // 6, 1)) * scale
    per_stream = np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert per_stream.max() <= REL_L2_TOL, "stream %d: rel_l2 %g" % (int(per_stream.argmax()), per_stream.max())
    first_hops = np.concatenate([got[:, :H], got[:, T * H:(T + 1) * H]], axis=1), np.concatenate([ref[:, :H], ref[:, T * H:(T + 1) * H]], axis=1)
    fh = np.linalg.norm(first_hops[0] - first_hops[1], axis=1) / np.linalg.norm(first_hops[1], axis=1)
    assert fh.max() <= 1e-3, "first hop of a call (carries the previous call's tail): stream %d rel_l2 %g" % (int(fh.argmax()), fh.max())


def test_lcmv_drops_hops_while_the_interference_list_is_restructured():
    """lcmv.cpp:271-276: an interference add/remove sets READY=false for 30 ms: the callback emits zeros and does not feed the ring
    buffers.  Offline the number of lost hops is the driver knob dropped_hops_on_restructure."""
    cfg = bf.make_config("lcmv", mics="circ8", interferers=(80.0, -60.0), dropped_hops_on_restructure=3)
    x = synth_batch(bf.GEOMETRIES["circ8"], 2, 70 * H, seed=91)
    events = [(20, "interf", 3, 150.0), (35, "interf", 2, -55.0), (50, "interf", 1, -54.5)]   # add (restructure), move, too close => remove (restructure)
    ref = np.stack([Oracle(cfg).process(x[b], events=events, dropped_hops=3) for b in range(2)])
    b = bf.Beamformer(cfg, n_streams=2)
    got = b.process(x, events=events)
    assert np.all(got[:, 20 * H:23 * H] == 0.0) and np.all(got[:, 50 * H:53 * H] == 0.0), "dropped hops are silent"
    assert np.any(got[:, 35 * H:38 * H] != 0.0), "a plain move does not drop hops"
    o = Oracle(cfg)
    o.process(x[0], events=events, dropped_hops=3)
    assert b.interferences == o.interferences
    assert finite_rel_l2(got, ref) <= REL_L2_TOL


# ---------------------------------------------------------------------------------------------
# magnitude-gated nodes: mvdr / lcmv / gss (device-resident path, selection flags captured)
# ---------------------------------------------------------------------------------------------
def run_device(cfg, x, events=(), capture=True, H=H):
    import torch
    B, M, L = x.shape
    T = L // H
    b = bf.Beamformer(cfg, n_streams=B)
    xin = torch.from_numpy(x).cuda()
    out = torch.empty((B, L), dtype=torch.float32, device="cuda")
    flags = torch.zeros((B, T, 2 * H), dtype=torch.uint8, device="cuda") if capture else None
    if capture:
        b.set_capture(flags.data_ptr())
    b.process_device(xin.data_ptr(), out.data_ptr(), T, stream_ptr=torch.cuda.current_stream().cuda_stream, events=events)
    torch.cuda.synchronize()
    return out.cpu().numpy(), (flags.cpu().numpy() if capture else None), b


def oracle_with_flags(cfg, x, events=()):
    outs, sels, msks = [], [], []
    for bidx in range(x.shape[0]):
        o, s, m = Oracle(cfg).process(x[bidx], events=events, want_flags=True)
        outs.append(o), sels.append(s), msks.append(m)
    return np.stack(outs), np.stack(sels), np.stack(msks)


def finite_rel_l2(got, ref):
    ok = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), ok), "non-finite samples must coincide (cold-start NaNs, SURVEY B-10)"
    return rel_l2(got[ok], ref[ok])


@pytest.mark.parametrize("mics,theta", [("circ8", 0.0), ("aira3", 20.0)])
def test_mvdr_matches_oracle_and_selection_is_bit_exact(mics, theta):
    cfg = bf.make_config("mvdr", mics=mics, initial_angle=theta)
    x = synth_batch(bf.GEOMETRIES[mics], 3, 61 * H, seed=21)
    ref, sel, _ = oracle_with_flags(cfg, x)
    got, flags, _ = run_device(cfg, x)
    assert sel.sum() > 1000, "the test signal must exercise the gate"
    assert np.array_equal(flags & 1, sel), "selected-bin set must be bit-exact"
    err = finite_rel_l2(got, ref)
    print("mvdr", mics, "rel_l2", err, "selected fraction", sel.mean())
    assert err <= REL_L2_TOL


@pytest.mark.parametrize("thr", [0.000022, 0.00001])
def test_mvdr_dense_selection_matches_oracle(thr):
    """Gate opened far enough that a frame pair carries several hundred items (25 % / all in-band bins selected): the kernel
    switches to its lane-pair scheme and runs more than one batch per pair."""
    cfg = bf.make_config("mvdr", mics="circ8", freq_mag_threshold=thr)
    x = synth_batch(bf.GEOMETRIES["circ8"], 2, 41 * H, seed=23)
    ref, sel, _ = oracle_with_flags(cfg, x)
    got, flags, _ = run_device(cfg, x)
    assert sel[:, 4:, :513].mean() > 0.2
    assert np.array_equal(flags & 1, sel), "selected-bin set must be bit-exact"
    err = finite_rel_l2(got, ref)
    print("mvdr dense", thr, "rel_l2", err, "selected fraction", sel[:, :, :513].mean())
    assert err <= REL_L2_TOL


def test_lcmv_with_interference_events_matches_oracle():
    cfg = bf.make_config("lcmv", mics="circ8", initial_angle=0.0, interferers=(80.0, -60.0, 150.0))
    x = synth_batch(bf.GEOMETRIES["circ8"], 2, 100 * H, seed=31)
    # SURVEY.md §8d event script, compressed in time: theta move, interferer move, add, too-close => remove, invalid id
    events = [(20, "theta", 20.0), (40, "interf", 2, -55.0), (60, "interf", 4, 120.0), (80, "interf", 1, 119.5), (90, "interf", 0, 10.0)]
    ref, sel, _ = oracle_with_flags(cfg, x, events=events)
    got, flags, b = run_device(cfg, x, events=events)
    o = Oracle(cfg)
    o.process(x[0], events=events)
    assert b.interferences == o.interferences, "interference list must be bit-exact"
    assert np.array_equal(flags & 1, sel)
    err = finite_rel_l2(got, ref)
    print("lcmv rel_l2", err)
    assert err <= REL_L2_TOL


def test_gss_with_events_matches_oracle():
    cfg = bf.make_config("gss", mics="circ8", initial_angle=0.0, interferers=(80.0, -60.0, 150.0))
    x = synth_batch(bf.GEOMETRIES["circ8"], 2, 100 * H, seed=41)
    events = [(20, "theta", 20.0), (40, "interf", 2, -55.0), (60, "interf", 4, 120.0), (80, "interf", 1, 119.5), (90, "interf", 0, 10.0)]
    ref, sel, _ = oracle_with_flags(cfg, x, events=events)
    got, flags, b = run_device(cfg, x, events=events)
    assert np.array_equal(flags & 1, sel)
    err = finite_rel_l2(got, ref)
    print("gss rel_l2", err)
    assert err <= REL_L2_TOL


@pytest.mark.parametrize("mics,interf", [("circ8", (70.0,)), ("circ8", (70.0, -110.0)), ("aira3", (80.0, -60.0)), ("circ8", (40.0, 80.0, 120.0, 160.0, -150.0, -100.0)),
                                         ("circ8", (40.0, 80.0, 120.0, 160.0, -150.0, -100.0, -50.0))])
def test_gss_row_groups_every_constraint_count(mics, interf):
    """sel_pairs_kernel<gss> splits the rows of the separation matrix over groups of 4 lanes (1 or 2 rows per lane): 2, 3, 7 and 8
    rows here (1 and 4-5 rows are in the neighbouring tests), state carried across calls of odd lengths."""
    kw = dict(mu=1e-4) if len(interf) > 4 else {}   # many rows: a smaller step keeps the recursion well away from divergence
    cfg = bf.make_config("gss", mics=mics, initial_angle=5.0, interferers=interf, **kw)
    x = synth_batch(bf.GEOMETRIES[mics], 3, 61 * H, seed=47)
    ref, sel, _ = oracle_with_flags(cfg, x)
    b = bf.Beamformer(cfg, n_streams=3)
    cuts = [0, 9 * H, 10 * H, 33 * H, 61 * H]
    got = np.concatenate([b.process(x[:, :, a:c]) for a, c in zip(cuts[:-1], cuts[1:])], axis=1)
    err = finite_rel_l2(got, ref)
    print("gss", mics, len(interf) + 1, "rows rel_l2", err, "selected fraction", sel.mean())
    assert sel.sum() > 500
    assert err <= REL_L2_TOL


def test_gss_without_interferers_keeps_geometric_gradient():
    # K = 0: the integer 1/(K+1) is 1, so dJ2 is active (SURVEY B-7)
    cfg = bf.make_config("gss", mics="aira3", initial_angle=10.0)
    x = synth_batch(bf.GEOMETRIES["aira3"], 2, 50 * H, seed=43)
    ref, sel, _ = oracle_with_flags(cfg, x)
    got, flags, _ = run_device(cfg, x)
    assert np.array_equal(flags & 1, sel)
    assert finite_rel_l2(got, ref) <= REL_L2_TOL


# ---------------------------------------------------------------------------------------------
# phase-mask nodes: phase / phasempf
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mics,theta", [("binaural", 0.0), ("aira3", 20.0), ("circ8", -30.0)])
def test_phase_matches_oracle(mics, theta):
    # phase.launch sets keys phase.cpp never reads, so the getParam fall-backs apply (SURVEY B-11);
    # mag_threshold is lowered here so that the synthetic signal exercises both sides of the gate
    cfg = bf.make_config("phase", mics=mics, initial_angle=theta, mag_threshold=0.002)
    x = synth_batch(bf.GEOMETRIES[mics], 2, 60 * H, seed=51)
    ref, _, msk = oracle_with_flags(cfg, x)
    got, flags, _ = run_device(cfg, x)
    kept = (flags >> 1) & 1
    assert msk.sum() > 100 and (1 - msk).sum() > 100
    mism = int((kept != msk).sum())
    err = finite_rel_l2(got, ref)
    print("phase", mics, "rel_l2", err, "mask mismatches", mism, "of", msk.size)
    assert err <= REL_L2_TOL
    assert mism <= MASK_MISMATCH_FRAC * msk.size, "phase-mask decisions must agree with the oracle (FP32 guard band + FP64 re-decision)"


@pytest.mark.parametrize("mics,theta,kw", [("binaural", 0.0, {}), ("aira3", 15.0, {}), ("binaural", 0.0, dict(out_only_mcra=True)),
                                           ("binaural", 0.0, dict(out_only_noise=True, smooth_size=20))])
def test_phasempf_matches_oracle(mics, theta, kw):
    cfg = bf.make_config("phasempf", mics=mics, initial_angle=theta, **kw)
    # > 3*MCRA_L frames so the first-window logic and two minima resets are exercised; sources gated on/off
    x = np.stack([synth_stream(bf.GEOMETRIES[mics], 170 * H, seed=61 + b, gate_hz=1.3) for b in range(2)])
    ref, _, msk = oracle_with_flags(cfg, x)
    got, flags, _ = run_device(cfg, x)
    kept = (flags >> 1) & 1
    mism = int((kept != msk).sum())
    err = finite_rel_l2(got, ref)
    print("phasempf", mics, kw, "rel_l2", err, "mask mismatches", mism, "of", msk.size)
    assert err <= REL_L2_TOL
    assert mism <= MASK_MISMATCH_FRAC * msk.size, "phase-mask decisions must agree with the oracle"


def test_phasempf_split_calls_carry_state():
    cfg = bf.make_config("phasempf", mics="binaural")
    x = np.stack([synth_stream(bf.GEOMETRIES["binaural"], 121 * H, seed=71 + b, gate_hz=1.3) for b in range(2)])
    ref = oracle_batch(cfg, x)
    b = bf.Beamformer(cfg, n_streams=2)
    got = np.concatenate([b.process(x[:, :, :33 * H]), b.process(x[:, :, 33 * H:34 * H]), b.process(x[:, :, 34 * H:])], axis=1)
    assert rel_l2(got, ref) <= REL_L2_TOL


# ---------------------------------------------------------------------------------------------
# golden vectors: outputs of the reference's own (unmodified) node sources, tests/golden/make_golden.py
# ---------------------------------------------------------------------------------------------
import os as _os

from golden.cases import CASES as GOLDEN_CASES, build_case as golden_build_case

_GOLD = np.load(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "ref_outputs.npz"))


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_gpu_matches_reference_golden(name):
    """Every golden case (all six nodes; 512-, 1024- and 4096-point frames; theta / interference events; the
    cold-start NaN case) through the C ABI against the reference's output."""
    cfg, x, events = golden_build_case(name)
    ref = _GOLD[name + "/out"]
    b = bf.Beamformer(cfg, n_streams=1)
    got = b.process(x[None], events=events)[0]
    assert got.shape == ref.shape
    err = finite_rel_l2(got, ref)
    print(name, "rel_l2 vs reference", err)
    assert err <= REL_L2_TOL
    assert b.interferences == list(_GOLD[name + "/interf"]), "interference list must be bit-exact"


@pytest.mark.parametrize("algo,mics,hop,kw", [("das", "aira3", 256, {}), ("das", "circ8", 1024, {}), ("das", "binaural", 2048, {}),
                                              ("phase", "aira3", 256, dict(mag_threshold=0.002)), ("phase", "binaural", 2048, dict(mag_threshold=0.002)),
                                              ("phasempf", "binaural", 2048, {}), ("phasempf", "aira3", 1024, dict(out_only_mcra=True))])
def test_other_frame_sizes_match_oracle(algo, mics, hop, kw):
    # 512- to 4096-point frames (JACK periods 256..2048) run the frame-size-generic kernel
    cfg = bf.make_config(algo, mics=mics, hop=hop, initial_angle=15.0, **kw)
    n_hops = 61 if algo != "phasempf" else 130
    x = np.stack([synth_stream(bf.GEOMETRIES[mics], n_hops * hop, seed=300 + b, gate_hz=1.3 if algo == "phasempf" else 0.0) for b in range(2)])
    ref = oracle_batch(cfg, x)
    b = bf.Beamformer(cfg, n_streams=2)
    k = 20 * hop
    got = np.concatenate([b.process(x[:, :, :k]), b.process(x[:, :, k:k + hop]), b.process(x[:, :, k + hop:])], axis=1)   # state carried across calls
    err = rel_l2(got, ref)
    print(algo, mics, "hop", hop, "rel_l2", err)
    assert err <= REL_L2_TOL


@pytest.mark.parametrize("env", ["BF_PHASE_F32", "BF_PHASE_F64"])
@pytest.mark.parametrize("algo,mics,hop,kw", [("phase", "aira3", 512, dict(mag_threshold=0.002)), ("phase", "binaural", 2048, dict(mag_threshold=0.002)),
                                              ("phasempf", "binaural", 2048, {}), ("phasempf", "aira3", 512, {}), ("phasempf", "aira3", 1024, {})])
def test_phase_kernels_both_precisions(monkeypatch, env, algo, mics, hop, kw):
    """The phase-mask nodes have two kernels per frame size: FP32 spectra with exact re-decisions (phase_n_kernel, default up
    to 1024 points, BF_PHASE_F32 forces it for longer frames) and the double-spectra / CTA-per-stream kernels (default for
    longer frames, BF_PHASE_F64 forces them).  Both must match the oracle, masks included."""
    monkeypatch.setenv(env, "1")
    cfg = bf.make_config(algo, mics=mics, hop=hop, initial_angle=15.0, **kw)
    n_hops = 61 if algo != "phasempf" else 130
    x = np.stack([synth_stream(bf.GEOMETRIES[mics], n_hops * hop, seed=700 + b, gate_hz=1.3 if algo == "phasempf" else 0.0) for b in range(2)])
    ref, _, msk = oracle_with_flags(cfg, x)
    got, flags, _ = run_device(cfg, x, H=hop)
    kept = (flags >> 1) & 1
    mism = int((kept != msk).sum())
    err = rel_l2(got, ref)
    print(env, algo, mics, "hop", hop, "rel_l2", err, "mask mismatches", mism, "of", msk.size)
    assert err <= REL_L2_TOL
    assert mism <= MASK_MISMATCH_FRAC * msk.size


@pytest.mark.parametrize("algo,mics,hop,kw", [("phasempf", "binaural", 2048, {}), ("phase", "aira3", 512, dict(mag_threshold=0.002)),
                                              ("phasempf", "aira3", 256, {}), ("phase", "binaural", 1024, dict(mag_threshold=0.002))])
@pytest.mark.timeout(180, method="thread")
def test_phase_fp32_kernel_with_every_significant_bin_re_decided(monkeypatch, capfd, algo, mics, hop, kw):
    """Stress of phase_n_kernel's deferred exact decisions: with the error bound blown up 10 000x (BF_DEBUG=100003: kappa = 1e-2) most
    significant bins of every frame are doubtful (a value ending in 003 also prints the count), the list of 160 items overflows and a pair takes several collect / decide / apply
    rounds (the pseudo-bin and the bin it folds into wait for each other).  Every decision is then the exact one, so masks and output
    must still match the oracle; streams are cut into calls of odd lengths."""
    monkeypatch.setenv("BF_PHASE_F32", "1")
    monkeypatch.setenv("BF_DEBUG", "100003")
    cfg = bf.make_config(algo, mics=mics, hop=hop, initial_angle=15.0, **kw)
    n_hops = 23
    x = np.stack([synth_stream(bf.GEOMETRIES[mics], n_hops * hop, seed=760 + b, gate_hz=1.3 if algo == "phasempf" else 0.0) for b in range(3)])
    ref, _, msk = oracle_with_flags(cfg, x)
    b = bf.Beamformer(cfg, n_streams=3)
    cuts = [0, 7 * hop, 8 * hop, 15 * hop, n_hops * hop]
    got = np.concatenate([b.process(x[:, :, a:c]) for a, c in zip(cuts[:-1], cuts[1:])], axis=1)
    err = rel_l2(got, ref)
    got1, flags, _ = run_device(cfg, x, H=hop)
    kept = (flags >> 1) & 1
    mism = int((kept != msk).sum())
    print(algo, mics, "hop", hop, "rel_l2", err, "single call", rel_l2(got1, ref), "mask mismatches", mism, "of", msk.size)
    assert err <= REL_L2_TOL and rel_l2(got1, ref) <= REL_L2_TOL
    assert mism <= MASK_MISMATCH_FRAC * msk.size
    import re
    import torch
    torch.cuda.synchronize()
    counts = [(int(m.group(1)), int(m.group(2))) for m in re.finditer(r"(\d+) exact re-decisions over (\d+) pairs", capfd.readouterr().out)]
    assert counts, "the kernel under test must be phase_n_kernel"
    assert max(c / max(n, 1) for c, n in counts) > 160, "the stress must overflow the list of a round: %r" % (counts,)


@pytest.mark.parametrize("algo,mics,hop,interf,events", [
    ("mvdr", "circ8", 1024, (), ()), ("mvdr", "aira3", 256, (), ()), ("mvdr", "circ12", 512, (), ()), ("mvdr", "circ16", 512, (), ()),
    ("lcmv", "circ8", 256, (80.0, -60.0, 150.0), ((15, "theta", 20.0), (25, "interf", 2, -55.0), (35, "interf", 4, 120.0), (45, "interf", 1, 119.5))),
    ("lcmv", "circ12", 512, (80.0, -60.0), ()), ("gss", "circ8", 1024, (80.0, -60.0, 150.0), ((20, "theta", 20.0), (30, "interf", 4, 120.0))),
    ("gss", "aira3", 2048, (), ())])
def test_gated_nodes_other_shapes_match_oracle(algo, mics, hop, interf, events):
    """mvdr / lcmv / gss outside the 1024-point, <= 8 microphone fast path: 512- to 4096-point frames and up to 16
    microphones (north_star: solves for M <= 16) run the general gated kernel; same gates as the fast path."""
    cfg = bf.make_config(algo, mics=mics, hop=hop, initial_angle=10.0, interferers=interf)
    n_hops = 61
    x = np.stack([synth_stream(bf.GEOMETRIES[mics], n_hops * hop, seed=500 + b) for b in range(2)])
    ref, sel, _ = oracle_with_flags(cfg, x, events=events)
    got, flags, b = run_device(cfg, x, events=events, H=hop)
    assert sel.sum() > 500, "the test signal must exercise the gate"
    assert np.array_equal(flags & 1, sel), "selected-bin set must be bit-exact"
    o = Oracle(cfg)
    o.process(x[0], events=events)
    assert b.interferences == o.interferences
    err = finite_rel_l2(got, ref)
    print(algo, mics, "hop", hop, "rel_l2", err, "selected fraction", sel.mean())
    assert err <= REL_L2_TOL


@pytest.mark.parametrize("algo,mics,hop,kw", [("mcra", "aira3", 512, dict(L=50)), ("mcra", "binaural", 2048, dict(L=20)), ("mcra", "circ8", 256, dict(L=30, out_only_noise=True)),
                                              ("mcra", "aira3", 1024, {}), ("ref", "aira3", 512, {}), ("ref", "binaural", 2048, {})])
def test_mcra_and_ref_nodes_match_oracle(algo, mics, hop, kw):
    """the stand-alone MCRA node (mcra.cpp) and rosjack_ref (jack_ref.cpp), state carried across calls."""
    cfg = bf.make_config(algo, mics=mics, hop=hop, **kw)
    n_hops = 131
    x = np.stack([synth_stream(bf.GEOMETRIES[mics], n_hops * hop, seed=600 + b, gate_hz=1.3) for b in range(3)])
    ref = oracle_batch(cfg, x)
    b = bf.Beamformer(cfg, n_streams=3)
    k = 20 * hop
    got = np.concatenate([b.process(x[:, :, :k]), b.process(x[:, :, k:k + hop]), b.process(x[:, :, k + hop:])], axis=1)
    err = rel_l2(got, ref)
    print(algo, mics, "hop", hop, kw, "rel_l2", err)
    assert err <= (0.0 if algo == "ref" else REL_L2_TOL)   # rosjack_ref is elementwise float arithmetic: bit-exact


@pytest.mark.parametrize("kw", [dict(L=50), dict(L=7, out_only_noise=True), {}])
def test_mcra_1024_both_kernels(monkeypatch, kw):
    """1024-point MCRA runs the warp-per-stream kernel (mcra_pairs_kernel); BF_MCRA_OLD keeps the CTA-per-stream kernel.  Both
    must match the oracle, with more streams than one CTA has warps, odd hop counts and state carried across calls."""
    cfg = bf.make_config("mcra", mics="aira3", hop=512, **kw)
    n_hops = 131
    x = np.stack([synth_stream(bf.GEOMETRIES["aira3"], n_hops * 512, seed=640 + b, gate_hz=1.3) for b in range(11)])
    ref = oracle_batch(cfg, x)
    for old in (False, True):
        if old:
            monkeypatch.setenv("BF_MCRA_OLD", "1")
        b = bf.Beamformer(cfg, n_streams=11)
        k = 21 * 512
        got = np.concatenate([b.process(x[:, :, :k]), b.process(x[:, :, k:k + 512]), b.process(x[:, :, k + 512:])], axis=1)
        err = rel_l2(got, ref)
        print("mcra 1024", kw, "old kernel" if old else "warp kernel", "rel_l2", err)
        assert err <= REL_L2_TOL


@pytest.mark.parametrize("mics,hop,kw,events", [("aira3", 512, {}, ((20, "theta", 25.0),)), ("circ8", 256, dict(initial_angle=-35.0, filter_size=64), ()),
                                                ("binaural", 2048, dict(use_vad=True, vad_threshold=0.08), ()), ("circ12", 512, dict(filter_size=256, mu0=0.0005, mu_max=0.01), ())])
def test_gsc_matches_oracle(mics, hop, kw, events):
    """generalized sidelobe canceller (gsc.cpp): per-microphone alignment + NLMS, state carried across calls."""
    cfg = bf.make_config("gsc", mics=mics, hop=hop, **kw)
    n_hops = 70
    x = np.stack([synth_stream(bf.GEOMETRIES[mics], n_hops * hop, seed=700 + b) for b in range(3)])
    ref = oracle_batch(cfg, x, events=events)
    b = bf.Beamformer(cfg, n_streams=3)
    if events:
        got = b.process(x, events=events)
    else:
        k = 20 * hop
        got = np.concatenate([b.process(x[:, :, :k]), b.process(x[:, :, k:k + hop]), b.process(x[:, :, k + hop:])], axis=1)
    err = finite_rel_l2(got, ref)
    print("gsc", mics, "hop", hop, kw, "rel_l2", err)
    assert err <= REL_L2_TOL


# ---------------------------------------------------------------------------------------------
# the drop-in boundary itself: bf_process_hop is the body of jack_callback (das.cpp:72-92) for every node
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo,mics,kw,setters", [
    ("mvdr", "circ8", {}, ()), ("lcmv", "circ8", dict(interferers=(80.0, -60.0, 150.0)), ((9, "interf", 2, -55.0), (15, "interf", 4, 120.0))),
    ("gss", "aira3", {}, ((12, "theta", 25.0),)), ("phase", "aira3", dict(mag_threshold=0.002), ()), ("phasempf", "binaural", {}, ()),
    ("mcra", "aira3", dict(L=10), ()), ("ref", "aira3", {}, ()), ("gsc", "aira3", {}, ((12, "theta", 25.0),))])
def test_hop_at_a_time_callback_matches_oracle_for_every_node(algo, mics, kw, setters):
    cfg = bf.make_config(algo, mics=mics, **kw)
    n_hops = 30
    x = synth_stream(bf.GEOMETRIES[mics], n_hops * H, seed=800)
    ref = Oracle(cfg).process(x, events=setters)
    b = bf.Beamformer(cfg, n_streams=1)
    outs = []
    for t in range(n_hops):
        for e in setters:   # the ROS callbacks (theta_roscallback / interf_theta_roscallback) arrive between two JACK periods
            if e[0] == t:
                b.set_theta(e[2]) if e[1] == "theta" else b.set_interference(e[2], e[3])
        outs.append(b.process_hop(x[:, t * H:(t + 1) * H]))
    got = np.concatenate(outs)
    err = finite_rel_l2(got, ref)
    print(algo, "hop-at-a-time rel_l2", err)
    assert err <= REL_L2_TOL


@pytest.mark.parametrize("algo,mics", [("das", "aira3"), ("mvdr", "circ8"), ("phasempf", "binaural"), ("gsc", "aira3")])
def test_empty_and_ragged_batches(algo, mics):
    """0 hops is a no-op; odd hop counts (a frame pair cut in half) and 1-hop calls give the same stream as one call."""
    cfg = bf.make_config(algo, mics=mics)
    x = np.stack([synth_stream(bf.GEOMETRIES[mics], 23 * H, seed=810 + b) for b in range(2)])
    ref = oracle_batch(cfg, x)
    b = bf.Beamformer(cfg, n_streams=2)
    assert b.process(x[:, :, :0]).shape == (2, 0)
    cuts = [0, 1, 4, 9, 10, 23]   # 1, 3, 5, 1, 13 hops
    got = np.concatenate([b.process(x[:, :, a * H:c * H]) for a, c in zip(cuts[:-1], cuts[1:])], axis=1)
    assert finite_rel_l2(got, ref) <= REL_L2_TOL


def test_two_handles_are_independent():
    """Per-handle state only (the reference keeps its state in process globals): two nodes interleaved on one device."""
    cfg_a, cfg_b = bf.make_config("mvdr", mics="circ8"), bf.make_config("phasempf", mics="binaural")
    xa = synth_stream(bf.GEOMETRIES["circ8"], 20 * H, seed=820)[None]
    xb = synth_stream(bf.GEOMETRIES["binaural"], 20 * H, seed=821)[None]
    a, b = bf.Beamformer(cfg_a, 1), bf.Beamformer(cfg_b, 1)
    ga = np.concatenate([a.process(xa[:, :, :8 * H]), a.process(xa[:, :, 8 * H:])], axis=1) if False else None
    pa, pb = [], []
    for k in range(0, 20, 4):
        pa.append(a.process(xa[:, :, k * H:(k + 4) * H]))
        pb.append(b.process(xb[:, :, k * H:(k + 4) * H]))
    assert finite_rel_l2(np.concatenate(pa, axis=1), oracle_batch(cfg_a, xa)) <= REL_L2_TOL
    assert finite_rel_l2(np.concatenate(pb, axis=1), oracle_batch(cfg_b, xb)) <= REL_L2_TOL


# ---------------------------------------------------------------------------------------------
# BASELINE.json's full sizes: size-independent properties + spot checks against the oracle
# ---------------------------------------------------------------------------------------------
def _full_size_case(algo, mics, n_streams, n_hops, hop, spot, **kw):
    """Runs the bench-shaped batch twice (streams in order, then reversed) on device-resident data:
       * a stream's output depends on that stream's input only: the reversed batch gives the reversed output, bit for bit
         (streams land on other SMs / warps / batch slots, so this also checks that no state leaks between streams);
       * `spot` streams are compared with the CPU oracle at the full length."""
    import torch
    from bench import device_synth
    dev = torch.device("cuda", 0)
    cfg = bf.make_config(algo, mics=mics, hop=hop, **kw)
    xy = bf.GEOMETRIES[mics]
    L = n_hops * hop
    x = device_synth(torch, xy, n_streams, L, seed=0xBEA4F0, device=dev)
    y = torch.empty((n_streams, L), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    bf.Beamformer(cfg, n_streams=n_streams).process_device(x.data_ptr(), y.data_ptr(), n_hops, stream_ptr=st)
    torch.cuda.synchronize()
    xr = torch.flip(x, dims=[0]).contiguous()
    yr = torch.empty_like(y)
    bf.Beamformer(cfg, n_streams=n_streams).process_device(xr.data_ptr(), yr.data_ptr(), n_hops, stream_ptr=st)
    torch.cuda.synchronize()
    same = torch.equal(torch.nan_to_num(torch.flip(yr, dims=[0]), nan=12345.0), torch.nan_to_num(y, nan=12345.0))
    assert same, "a stream's output must not depend on its position in the batch"
    for sidx in spot:
        ref = Oracle(cfg).process(x[sidx].cpu().numpy())
        err = finite_rel_l2(y[sidx].cpu().numpy(), ref)
        print(algo, "full size, stream", sidx, "rel_l2", err)
        assert err <= REL_L2_TOL


def test_full_size_c1_das():
    _full_size_case("das", "aira3", 2048, 188, 512, spot=(0, 777, 2047))


def test_full_size_c2_mvdr_batched_1k_streams():
    # 1184 streams on 148 persistent CTAs: eight streams per CTA; the spot streams cover first / middle / last positions of a CTA's sequence
    _full_size_case("mvdr", "circ8", 1184, 188, 512, spot=(0, 1, 147, 148, 149, 295, 296, 500, 591, 592, 740, 887, 888, 1035, 1036, 1183))


def test_full_size_c3_lcmv_and_gss():
    _full_size_case("lcmv", "circ8", 592, 94, 512, spot=(5,), interferers=(80.0, -60.0, 150.0))
    _full_size_case("gss", "circ8", 592, 94, 512, spot=(5,), interferers=(80.0, -60.0, 150.0))


def test_full_size_c4_phasempf_4096():
    _full_size_case("phasempf", "binaural", 1184, 47, 2048, spot=(3, 1100))


def test_c1_sixty_second_stream_matches_oracle():
    """BASELINE.json configs[0]: DAS, 3 microphones (aira3), 48 kHz, 1024-point frames, a 60 s signal (5625 hops)."""
    cfg = bf.make_config("das", mics="aira3", initial_angle=0.0)
    x = synth_stream(bf.GEOMETRIES["aira3"], 5625 * H, seed=0xC1)
    ref = Oracle(cfg).process(x)
    got = bf.Beamformer(cfg, n_streams=1).process(x[None])[0]
    err = rel_l2(got, ref)
    print("C1 60 s rel_l2", err)
    assert err <= REL_L2_TOL


def _launch_text(algo):
    """A launch file in the reference's layout (launch/mvdr.launch:1-13) carrying the node's <rosparam> block."""
    from beamform_b200 import tables
    body = "".join("      %s: %s\n" % (k, ("true" if v else "false") if isinstance(v, bool) else v) for k, v in tables.LAUNCH_PARAMS[algo].items())
    return ('<launch>\n  <node name="beamform" pkg="beamform" type="%s" output="screen">\n    <rosparam command="load" file="$(find beamform)/rosjack_config.yaml" />\n'
            '    <rosparam command="load" file="$(find beamform)/beamform_config.yaml" />\n    <rosparam>\n%s    </rosparam>\n  </node>\n</launch>\n' % (algo, body))


def test_offline_file_driver(tmp_path):
    """tools/bf_offline (C++): wav + beamform_config.yaml + launch file in, wav out (the stand-in for the JACK/ROS transport);
    tools/beamform_file.py is its wrapper."""
    import subprocess
    import sys as _sys
    from scipy.io import wavfile
    xy = bf.GEOMETRIES["aira3"]
    x = synth_stream(xy, 40 * H + 123, seed=0xF11E)           # a ragged tail: JACK only delivers whole periods
    yaml = tmp_path / "beamform_config.yaml"
    yaml.write_text("initial_angle: 20.0\n" + "".join("mic%d: {id: %d, x: %r, y: %r}\n" % (i, i + 1, px, py) for i, (px, py) in enumerate(xy)))
    wavfile.write(str(tmp_path / "in.wav"), 48000, x.T.astype(np.float32))
    root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    subprocess.run([_sys.executable, _os.path.join(root, "tools", "beamform_file.py"), "--algo", "mvdr", "--config", str(yaml), "--in", str(tmp_path / "in.wav"),
                    "--out", str(tmp_path / "out.wav"), "--theta-at", "15:-30"], check=True)
    sr, got = wavfile.read(str(tmp_path / "out.wav"))
    cfg = bf.load_yaml_config("mvdr", str(yaml))
    ref = Oracle(cfg).process(x[:, :40 * H], events=[(15, "theta", -30.0)])
    assert sr == 48000 and got.shape == ref.shape
    assert finite_rel_l2(got, ref) <= REL_L2_TOL
    # the C++ tool itself: launch file, event file, 16-bit PCM input, raw float32 output
    (tmp_path / "lcmv.launch").write_text(_launch_text("lcmv"))
    (tmp_path / "events.txt").write_text("12 theta 15.0\n20 interf 1 60.0\n")
    xy8 = bf.GEOMETRIES["circ8"]
    x8 = synth_stream(xy8, 36 * H, seed=0xF12E)
    q = np.clip(np.round(x8 * 32768.0), -32768, 32767).astype(np.int16)
    wavfile.write(str(tmp_path / "in16.wav"), 48000, q.T)
    yaml8 = tmp_path / "circ8.yaml"
    yaml8.write_text("initial_angle: 0.0\n" + "".join("mic%d: {id: %d, x: %r, y: %r}\n" % (i, i + 1, px, py) for i, (px, py) in enumerate(xy8)) + "angle_interf1: 80\nangle_interf2: 181\n")
    subprocess.run([_os.path.join(root, "tools", "bf_offline"), "--algo", "lcmv", "--config", str(yaml8), "--launch", str(tmp_path / "lcmv.launch"), "--in", str(tmp_path / "in16.wav"),
                    "--out", str(tmp_path / "out.f32"), "--events", str(tmp_path / "events.txt")], check=True)
    got = np.fromfile(str(tmp_path / "out.f32"), dtype=np.float32)
    cfg8 = bf.make_config("lcmv", mics="circ8", interferers=(80.0,))
    ref = Oracle(cfg8).process(q.astype(np.float32) / 32768.0, events=[(12, "theta", 15.0), (20, "interf", 1, 60.0)])
    assert got.shape == ref.shape and finite_rel_l2(got, ref) <= REL_L2_TOL


def test_unsupported_shapes_fail_loudly():
    with pytest.raises(bf.BeamformError):
        bf.Beamformer(bf.make_config("mvdr", mics="grid64"), 1)               # solves are built for <= 16 microphones
    with pytest.raises(bf.BeamformError):
        bf.Beamformer(bf.make_config("phasempf", mics="circ16", hop=2048), 1)   # the phase mask needs all 16 x 4096-point FP64 spectra of a bin at once
    with pytest.raises(bf.BeamformError):
        bf.Beamformer(bf.make_config("das", mics="aira3", hop=300), 1)        # JACK periods are powers of two (256..2048)
    with pytest.raises(bf.BeamformError):
        bf.Beamformer(bf.make_config("lcmv", mics="circ8", interferers=tuple(range(-160, 160, 20))), 1)   # 16 interferers: the yaml ships 15 slots


@pytest.mark.parametrize("algo,mics,hop,interf,events,kw", [
    ("das", "circ8", 2048, (), (), {}), ("das", "circ16", 2048, (), ((9, "theta", 40.0),), {}), ("das", "circ16", 1024, (), (), {}),
    ("mvdr", "circ8", 2048, (), (), {}), ("lcmv", "circ8", 2048, (80.0, -60.0, 150.0), ((11, "interf", 2, -55.0),), {}),
    # more than 7 interferers.  lcmv with that many constraints is only well-posed where the steering vectors differ (the 0.3 m array
    # cannot tell ten directions apart at 100 Hz: the reference's own output blows up there), hence the 4-9 kHz band
    ("lcmv", "circ16", 512, tuple(-170.0 + 34.0 * k for k in range(9)), ((15, "interf", 10, 5.0), (25, "interf", 3, -100.5)),
     dict(past_windows=40, freq_min=4000, freq_max=9000)),
    ("gss", "circ12", 512, tuple(-150.0 + 30.0 * k for k in range(9)), tuple((10 + 3 * k, "interf", 10 + k, -165.0 + 30.0 * k) for k in range(6)), dict(mu=1e-5))])
def test_lifted_shape_limits_match_oracle(algo, mics, hop, interf, events, kw):
    """Shapes round 1 refused: das with 8-16 microphones at 4096-point frames (microphones pass through shared memory in
    chunks and accumulate, das.cpp:60-63 is linear), mvdr / lcmv with 8 microphones at 4096 points (spectra spill to a global
    workspace), and up to the 15 interferers beamform_config.yaml:43-57 has slots for (general gated kernel)."""
    cfg = bf.make_config(algo, mics=mics, hop=hop, initial_angle=10.0, interferers=interf, **kw)
    n_hops = 31 if hop >= 1024 else 61
    x = np.stack([synth_stream(bf.GEOMETRIES[mics], n_hops * hop, seed=900 + b, sources=((20.0, 0.1, 180.0, 60), (-70.0, 0.05, 233.0, 60))) for b in range(2)])
    o = Oracle(cfg)
    ref0, sel, _ = o.process(x[0], events=events, want_flags=True)
    ref = np.stack([ref0, Oracle(cfg).process(x[1], events=events)])
    b = bf.Beamformer(cfg, n_streams=2)
    got = b.process(x, events=events)
    assert b.interferences == o.interferences
    if algo in ("lcmv", "gss"):
        assert len(o.interferences) >= (10 if mics != "circ8" else 3) and sel.sum() > 300, "the case must exercise the gate with the long list"
    err = finite_rel_l2(got, ref)
    print(algo, mics, "hop", hop, "interferers", len(o.interferences), "rel_l2", err)
    assert err <= REL_L2_TOL


# ---------------------------------------------------------------------------------------------
# steered-response sweep (config C5)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mics,n_dirs,n_hops", [("grid64", 360, 9), ("circ8", 72, 12), ("aira3", 10, 5)])
def test_srp_maps_match_oracle(mics, n_dirs, n_hops):
    import torch
    cfg = bf.make_config("das", mics=mics)
    xy = bf.GEOMETRIES[mics]
    x = np.stack([synth_stream(xy, n_hops * H, seed=400 + b, sources=((25.0 + 40 * b, 0.1, 190.0, 20), (-100.0, 0.05, 233.0, 20)),
                               lead_in=0) for b in range(2)])
    thetas = -180.0 + 360.0 * np.arange(n_dirs) / n_dirs
    b = bf.Beamformer(cfg, n_streams=2)
    xin = torch.from_numpy(x).cuda()
    maps = torch.zeros((2, n_hops, n_dirs), dtype=torch.float32, device="cuda")
    b.srp_device(xin.data_ptr(), thetas, maps.data_ptr(), n_hops, stream_ptr=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = maps.cpu().numpy().astype(np.float64)
    ref = np.stack([Oracle(cfg).srp(x[s], thetas.astype(np.float32).astype(np.float64)) for s in range(2)])
    err = rel_l2(got, ref)
    print("srp", mics, "rel_l2", err)
    assert err <= REL_L2_TOL
    # the map peaks at the dominant source direction (within the sweep resolution and the array's beam width)
    if mics == "grid64":
        peak = thetas[int(np.argmax(ref[0, n_hops - 1]))]
        assert abs(((peak - 25.0 + 180) % 360) - 180) <= 6.0
        assert int(np.argmax(got[0, n_hops - 1])) == int(np.argmax(ref[0, n_hops - 1]))


# ---------------------------------------------------------------------------------------------
# the drop-in from C++: examples/<node>_b200.cpp = the reference's node with rosjack.h / util.h unchanged (handle_params,
# rosjack_create, ros::spin) and the DSP behind the C ABI, driven hop by hop by the offline ROS/JACK stand-in
# ---------------------------------------------------------------------------------------------
import ref_lib as _ref_lib

_NODE_OF = {"ref": "jack_ref"}


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
@pytest.mark.parametrize("seam", [False, True])
def test_cpp_drop_in_nodes_match_reference_golden(name, seam):
    """jack_callback -> bf_process_hop (das.cpp:72-92), theta_roscallback -> bf_set_theta (das.cpp:94-99),
    interf_theta_roscallback -> bf_set_interference (lcmv.cpp:258-309); with seam=True the binding sits at the per-frame
    operator instead: util.h's own do_overlap calls bf_apply_weights as its weight_func (util.h:289)."""
    cfg, x, events = golden_build_case(name)
    algo = GOLDEN_CASES[name]["algo"]
    exe = "%s_b200%s" % (_NODE_OF.get(algo, algo), "_seam" if seam else "")
    if seam and algo not in ("das", "mvdr", "lcmv", "gss", "phasempf"):
        pytest.skip("no per-frame operator seam for this node")
    if not _ref_lib.example_available(exe):
        pytest.skip("examples/_bin/%s was not built (needs the reference's rosjack.h / util.h at build time)" % exe)
    got, interf = _ref_lib.run_ref(algo, cfg, x, events=events, want_interf=True, binary=_os.path.join(_ref_lib.EXAMPLE_DIR, exe))
    ref = _GOLD[name + "/out"]
    assert got.shape == ref.shape
    err = finite_rel_l2(got, ref)
    print(exe, name, "rel_l2 vs reference", err)
    assert err <= REL_L2_TOL
    if algo in ("lcmv", "gss"):
        assert interf == list(_GOLD[name + "/interf"]), "interference list must be bit-exact"


def test_srp_closed_loop_follows_a_moving_source():
    """the steered-response arg-max drives bf_set_theta (what scripts/energy2theta*.py do by gradient search)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("srp_steer", _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tools", "srp_steer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    xy = bf.GEOMETRIES["circ8"]
    a = synth_stream(xy, 64 * H, sources=((30.0, 0.1, 190.0, 30),), lead_in=0, seed=1)
    b = synth_stream(xy, 64 * H, sources=((-70.0, 0.1, 190.0, 30),), lead_in=0, seed=2)
    cfg = bf.make_config("das", mics="circ8")
    y, track = mod.follow(cfg, np.concatenate([a, b], axis=1), block_hops=8)
    assert y.shape == (128 * H,) and np.isfinite(y).all()
    assert abs(track[6] - 30.0) <= 6.0, track          # settled on the first source
    assert abs(track[-1] + 70.0) <= 6.0, track         # ... and followed it to the second position
    # the loop's output equals the oracle run with the same theta schedule (set_theta before block k applies from its first hop)
    events = [(8 * k, "theta", float(th)) for k, th in enumerate(track)]
    ref = Oracle(cfg).process(np.concatenate([a, b], axis=1), events=events)
    assert rel_l2(y, ref) <= REL_L2_TOL


@pytest.mark.parametrize("kw", [dict(past_windows=4), dict(past_windows=1), dict(past_windows=7, freq_min=400, freq_max=4000),
                                dict(freq_max=20000), dict(past_windows=12), dict(freq_min=1000, freq_max=16400)])
def test_mvdr_parameter_corners_match_oracle(kw):
    """The pipelined kernel keeps P + 1 <= 11 ring slots per bin in tensor memory for bins below 352: shorter histories (ring arithmetic
    modulo P + 1), narrower bands, a band edge in the last on-chip block, and the shapes outside it (P > 10, band above bin 351), which
    run sel_pairs_kernel -- all against the oracle, split into three calls so that the ring is saved and reloaded."""
    cfg = bf.make_config("mvdr", mics="circ8", initial_angle=15.0, **kw)
    x = synth_batch(bf.GEOMETRIES["circ8"], 3, 47 * H, seed=77)
    ref, sel, _ = oracle_with_flags(cfg, x)
    b = bf.Beamformer(cfg, n_streams=3)
    got = np.concatenate([b.process(x[:, :, :13 * H]), b.process(x[:, :, 13 * H:14 * H]), b.process(x[:, :, 14 * H:])], axis=1)
    assert sel.sum() > 300
    err = finite_rel_l2(got, ref)
    print("mvdr", kw, "rel_l2", err, "selected", int(sel.sum()))
    assert err <= REL_L2_TOL
