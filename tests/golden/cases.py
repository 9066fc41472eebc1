"""Golden-vector case table shared by make_golden.py (generator) and the tests (consumers)."""
H = 512
EV_C3 = [(20, "theta", 20.0), (40, "interf", 2, -55.0), (60, "interf", 4, 120.0), (80, "interf", 1, 119.5), (90, "interf", 0, 10.0)]

# name -> dict(algo, mics, hops, seed, hop (JACK period), cfg overrides, events, synth overrides)
CASES = {
    "das_aira3_theta_events": dict(algo="das", mics="aira3", hops=40, seed=101, events=[(7, "theta", 20.0), (8, "theta", -45.0), (20, "theta", 110.0)]),
    "das_circ8_hop256": dict(algo="das", mics="circ8", hops=48, seed=102, hop=256, cfg=dict(initial_angle=-35.0)),
    "das_binaural_hop2048": dict(algo="das", mics="binaural", hops=12, seed=103, hop=2048, cfg=dict(initial_angle=90.0)),
    "mvdr_circ8": dict(algo="mvdr", mics="circ8", hops=60, seed=104),
    "mvdr_aira3_coldstart_nan": dict(algo="mvdr", mics="aira3", hops=24, seed=105, synth=dict(lead_in=0), cfg=dict(initial_angle=20.0)),
    "lcmv_circ8_c3_events": dict(algo="lcmv", mics="circ8", hops=100, seed=106, cfg=dict(interferers=(80.0, -60.0, 150.0)), events=EV_C3),
    "gss_circ8_c3_events": dict(algo="gss", mics="circ8", hops=100, seed=107, cfg=dict(interferers=(80.0, -60.0, 150.0)), events=EV_C3),
    "gss_aira3_k0": dict(algo="gss", mics="aira3", hops=50, seed=108, cfg=dict(initial_angle=10.0)),
    "phase_aira3": dict(algo="phase", mics="aira3", hops=60, seed=109, cfg=dict(initial_angle=20.0, mag_threshold=0.002)),
    "phase_circ8": dict(algo="phase", mics="circ8", hops=30, seed=110, cfg=dict(initial_angle=-30.0, mag_threshold=0.002)),
    "phasempf_binaural": dict(algo="phasempf", mics="binaural", hops=170, seed=111, synth=dict(gate_hz=1.3)),
    "phasempf_aira3_only_mcra": dict(algo="phasempf", mics="aira3", hops=120, seed=112, synth=dict(gate_hz=1.3), cfg=dict(initial_angle=15.0, out_only_mcra=True)),
    "phasempf_binaural_hop2048_c4": dict(algo="phasempf", mics="binaural", hops=60, seed=113, hop=2048, synth=dict(gate_hz=1.3)),
    # the stand-alone MCRA node (launch/mcra.launch) and rosjack_ref
    "mcra_aira3": dict(algo="mcra", mics="aira3", hops=170, seed=114, synth=dict(gate_hz=1.3), cfg=dict(L=50)),
    "mcra_binaural_hop2048_only_noise": dict(algo="mcra", mics="binaural", hops=60, seed=115, hop=2048, synth=dict(gate_hz=1.3), cfg=dict(L=20, out_only_noise=True)),
    "ref_aira3": dict(algo="ref", mics="aira3", hops=40, seed=116),
    "ref_circ8_hop256": dict(algo="ref", mics="circ8", hops=48, seed=117, hop=256),
    # generalized sidelobe canceller (launch/gsc.launch)
    "gsc_aira3_theta_event": dict(algo="gsc", mics="aira3", hops=60, seed=118, events=[(25, "theta", 20.0)]),
    "gsc_circ8_hop256_vad": dict(algo="gsc", mics="circ8", hops=80, seed=119, hop=256, cfg=dict(initial_angle=-35.0, use_vad=True, vad_threshold=0.08, filter_size=64)),
}


def build_case(name):
    """-> (cfg (product BfConfig), x [M][L] float32, events)"""
    import beamform_b200 as bf
    from beamform_b200.synth import synth_stream
    c = CASES[name]
    hop = c.get("hop", H)
    cfg = bf.make_config(c["algo"], mics=c["mics"], hop=hop, **c.get("cfg", {}))
    x = synth_stream(bf.GEOMETRIES[c["mics"]], c["hops"] * hop, seed=c["seed"], **c.get("synth", {}))
    return cfg, x, list(c.get("events", []))
