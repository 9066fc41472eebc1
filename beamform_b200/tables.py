"""Pure data shared by the host mirror, the tests and bench.py: the <rosparam> blocks of the reference's launch
files, the microphone geometries, and the getParam fall-backs of every node (the values bf_config_init fills in,
beamform_b200/csrc/capi.cu).  No imports from the package: bench.py's reference arm loads this file by path so that
it never touches the product library."""
import math

# <rosparam> blocks of launch/*.launch: the only place the reference's operating values live.
LAUNCH_PARAMS = {
    "das": {},
    "mvdr": dict(past_windows=10, freq_mag_threshold=0.001, freq_max=16000, freq_min=100, out_amp=1.0),
    "lcmv": dict(past_windows=10, freq_mag_threshold=0.001, freq_max=16000, freq_min=100, out_amp=1.0, interf_angle_threshold=1.0),
    "gss": dict(freq_mag_threshold=0.001, freq_max=16000, freq_min=100, out_amp=0.1, interf_angle_threshold=1.0, mu=0.001, **{"lambda": 0.0}),
    "phase": dict(min_phase=10.0, min_mag=0.05, smooth_size=5),   # min_mag/smooth_size are never read by phase.cpp (B-11)
    "phasempf": dict(min_phase=30.0, min_mag=0.05, smooth_size=3, MCRA_alphaS=0.95, MCRA_alphaD=0.95, MCRA_alphaD2=0.98,
                     MCRA_delta=0.001, MCRA_L=50, MPF_alphaS=0.7, MPF_eta=0.3, MPF_rev_gamma=0.9, MPF_rev_delta=1.0,
                     out_amp=2.5, noise_floor=0.001, out_only_noise=False, out_only_mcra=False),
    "mcra": dict(alphaS=0.95, alphaD=0.95, alphaD2=0.98, delta=0.001, L=300, out_amp=3.5, out_only_noise=False),   # launch/mcra.launch
    "ref": {},
    "gsc": dict(use_vad=False, vad_threshold=0.1, mu0=0.0001, mu_max=0.1, filter_size=128),   # launch/gsc.launch (write_mu is a log file)
}

# beamform_config.yaml geometries (lines 15-17, 38-39) and the synthetic ones SURVEY.md §8d names
GEOMETRIES = {
    "aira3": [(0.000, 0.000), (0.000, -0.180), (-0.156, -0.090)],
    "binaural": [(0.000, 0.000), (0.000, -0.342)],
    "circ8": [(0.10 * math.cos(2 * math.pi * k / 8), 0.10 * math.sin(2 * math.pi * k / 8)) for k in range(8)],
    "circ12": [(0.12 * math.cos(2 * math.pi * k / 12), 0.12 * math.sin(2 * math.pi * k / 12)) for k in range(12)],
    "circ16": [(0.15 * math.cos(2 * math.pi * k / 16), 0.15 * math.sin(2 * math.pi * k / 16)) for k in range(16)],
    "grid64": [(0.04 * (k % 8), 0.04 * (k // 8)) for k in range(64)],
}


ALGOS = {"das": 0, "mvdr": 1, "lcmv": 2, "gss": 3, "phase": 4, "phasempf": 5, "mcra": 6, "ref": 7, "gsc": 8}


def config_defaults(algo):
    """Field values after bf_config_init(cfg, algo): the getParam fall-backs of the reference nodes (mvdr.cpp:155-184,
    lcmv.cpp:179-216, gss.cpp:186-237, phase.cpp:170-189, phasempf.cpp:361-470, mcra.cpp:181-224, gsc.cpp:206-258)."""
    return dict(
        algo=ALGOS[algo], sample_rate=48000.0, hop=512, initial_angle=0.0, past_windows=10, freq_mag_threshold=1.5, freq_max=4000.0,
        freq_min=400.0, out_amp=2.0 if algo in ("phasempf", "mcra") else 4.5, interf_angle_threshold=5.0, mu=0.01, lambda_=0.0,
        min_phase=10.0, mag_mult=0.1, mag_threshold=0.05, min_mag=10.0, smooth_size=20, MCRA_alphaS=0.95, MCRA_alphaD=0.95,
        MCRA_alphaD2=0.97, MCRA_delta=0.001, MCRA_L=0, MPF_alphaS=0.3, MPF_eta=0.3, MPF_rev_gamma=0.3, MPF_rev_delta=1.0,
        noise_floor=0.001, out_only_noise=1 if algo == "mcra" else 0, out_only_mcra=0, use_vad=0, vad_threshold=0.1, mu0=0.0005,
        mu_max=0.01, filter_size=128)


# rosparam key -> config field where the names differ (the mcra node drops the MCRA_ prefix, mcra.cpp:181-224)
KEY_FIELD = {"alphaS": "MCRA_alphaS", "alphaD": "MCRA_alphaD", "alphaD2": "MCRA_alphaD2", "delta": "MCRA_delta", "L": "MCRA_L", "lambda": "lambda_",
             "period": "hop"}
INT_FIELDS = {"past_windows", "smooth_size", "MCRA_L", "filter_size", "hop", "out_only_noise", "out_only_mcra", "use_vad"}


def plain_config_fields(algo, mics="aira3", hop=512, sample_rate=48000, initial_angle=0.0, interferers=(), launch=True, **params):
    """The same configuration beamform_b200.make_config builds, as a plain dict (no library call): fall-backs, then the
    launch-file block, then overrides; plus n_mics / mic_x / mic_y / angle_interf lists."""
    f = config_defaults(algo)
    f.update(hop=int(hop), sample_rate=float(sample_rate), initial_angle=float(initial_angle))
    kv = dict(LAUNCH_PARAMS[algo]) if launch else {}
    kv.update(params)
    for k, v in kv.items():
        name = KEY_FIELD.get(k, k)
        if name not in f:
            continue   # keys a node never reads are ignored (SURVEY B-11)
        if name == "smooth_size":
            v = 20 if int(v) < 1 else int(v)   # phasempf.cpp:377-381
        f[name] = int(v) if name in INT_FIELDS else float(v)
    xy = GEOMETRIES[mics] if isinstance(mics, str) else list(mics)
    f["n_mics"] = len(xy)
    f["mic_x"] = [float(x) for x, _ in xy]
    f["mic_y"] = [float(y) for _, y in xy]
    f["angle_interf"] = [float(a) for a in interferers]
    f["n_angle_interf"] = len(f["angle_interf"])
    return f
