"""Runner for oracle/_ref/<node>_ref: the reference's UNMODIFIED node sources compiled against oracle/shim/.
TEST INFRASTRUCTURE ONLY.  The binaries exist where /root/reference was available at build time
(`make -C oracle -f Makefile.ref`); they travel to the GPU box with the tree but are never required there."""
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

# every rosparam key a node's <algo>_handle_params reads (names as in the reference sources)
NODE_KEYS = {
    "das": [],
    "mvdr": ["past_windows", "freq_mag_threshold", "freq_max", "freq_min", "out_amp"],
    "lcmv": ["past_windows", "freq_mag_threshold", "freq_max", "freq_min", "out_amp", "interf_angle_threshold"],
    "gss": ["freq_mag_threshold", "freq_max", "freq_min", "out_amp", "interf_angle_threshold", "mu", "lambda"],
    "phase": ["min_phase", "mag_mult", "mag_threshold"],
    "phasempf": ["min_phase", "min_mag", "smooth_size", "MCRA_alphaS", "MCRA_alphaD", "MCRA_alphaD2", "MCRA_delta", "MCRA_L",
                 "MPF_alphaS", "MPF_eta", "MPF_rev_gamma", "MPF_rev_delta", "out_amp", "noise_floor", "out_only_noise", "out_only_mcra"],
    "mcra": ["alphaS", "alphaD", "alphaD2", "delta", "L", "out_amp", "out_only_noise"],
    "ref": [],
    "gsc": ["use_vad", "vad_threshold", "mu0", "mu_max", "filter_size"],
}
# rosparam name -> config field where they differ (the mcra node drops the MCRA_ prefix, mcra.cpp:181-224)
KEY_FIELD = {"alphaS": "MCRA_alphaS", "alphaD": "MCRA_alphaD", "alphaD2": "MCRA_alphaD2", "delta": "MCRA_delta", "L": "MCRA_L", "lambda": "lambda_"}
BINARY = {"ref": "jack_ref_ref"}
INT_KEYS = {"past_windows", "smooth_size", "MCRA_L", "L", "filter_size"}
BOOL_KEYS = {"out_only_noise", "out_only_mcra", "use_vad"}


def available(algo="das"):
    return os.path.exists(os.path.join(REF_DIR, BINARY.get(algo, algo + "_ref")))


def build():
    if os.path.isdir("/root/reference/beamform/src"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "-f", "Makefile.ref"], check=True)


def run_ref(algo, cfg, x, events=(), want_interf=False, binary=None):
    """cfg: beamform_b200.BfConfig (or the oracle's BfoConfig) — every field the node reads is passed as a rosparam.
    x: [M][L] float32.  Returns out [L] float32 (and the final interference list).  `binary`: run this executable instead
    of oracle/_ref/<node>_ref (the drop-in nodes of examples/ speak the same offline ROS/JACK stand-in)."""
    x = np.ascontiguousarray(x[:1] if algo == "ref" else x, dtype=np.float32)   # rosjack_ref opens one JACK input (jack_ref.cpp:65)
    M, L = x.shape
    with tempfile.TemporaryDirectory() as td:
        pf, inf, outf, evf, itf = (os.path.join(td, n) for n in ("params.txt", "in.f32", "out.f32", "events.txt", "interf.txt"))
        with open(pf, "w") as f:
            f.write("verbose false\ninitial_angle %r\n" % float(cfg.initial_angle))
            for m in range(cfg.n_mics):
                f.write("mic%d %r %r\n" % (m, float(cfg.mic_x[m]), float(cfg.mic_y[m])))
            for k in range(cfg.n_angle_interf):
                f.write("angle_interf%d %r\n" % (k + 1, float(cfg.angle_interf[k])))
            for key in NODE_KEYS[algo]:
                v = getattr(cfg, KEY_FIELD.get(key, key))
                if key in BOOL_KEYS:
                    f.write("%s %s\n" % (key, "true" if v else "false"))
                elif key in INT_KEYS:
                    f.write("%s %d\n" % (key, int(v)))
                else:
                    f.write("%s %r\n" % (key, float(v)))
        x.tofile(inf)
        with open(evf, "w") as f:
            for e in sorted(events, key=lambda e: e[0]):
                if e[1] == "theta":
                    f.write("%d theta %r\n" % (e[0], float(np.float32(e[2]))))
                else:
                    f.write("%d interf %d %r\n" % (e[0], e[2], float(np.float32(e[3]))))
        env = dict(os.environ, BFREF_PARAMS=pf, BFREF_IN=inf, BFREF_OUT=outf, BFREF_EVENTS=evf, BFREF_INTERF_OUT=itf,
                   BFREF_HOP=str(int(cfg.hop)), BFREF_SR=str(int(cfg.sample_rate)))
        env.pop("BFREF_VERBOSE", None)
        subprocess.run([binary or os.path.join(REF_DIR, BINARY.get(algo, algo + "_ref"))], env=env, check=True, stdout=subprocess.DEVNULL)
        out = np.fromfile(outf, dtype=np.float32)
        interf = [float(s) for s in open(itf).read().split()] if os.path.exists(itf) else []
    return (out, interf) if want_interf else out


EXAMPLE_DIR = os.path.join(ROOT, "examples", "_bin")


def example_available(name):
    return os.path.exists(os.path.join(EXAMPLE_DIR, name))


def build_examples():
    """The drop-in nodes of examples/ need the reference's rosjack.h / util.h: built where /root/reference exists."""
    if os.path.isdir("/root/reference/beamform/src"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "-s"], check=True)
