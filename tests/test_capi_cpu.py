"""CPU: the C-ABI library loads, exports every symbol include/beamform_b200.h declares, and its host-only entry
points (configuration ingest, error paths) behave like the reference's parameter handling.  No compute calls."""
import ctypes as C
import os
import re

import pytest

import beamform_b200 as bf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "beamform_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = bf.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libbeamform_b200.so does not export %s" % n


def test_struct_layout_matches_header():
    # bf_config_init zero-fills sizeof(bf_config) bytes: a ctypes image that is too small would be overrun
    guard = (C.c_uint8 * (C.sizeof(bf.BfConfig) + 64))()
    C.memset(guard, 0xAB, len(guard))
    cfg = bf.BfConfig.from_buffer(guard)
    assert bf.lib().bf_config_init(C.byref(cfg), 1) == 0
    assert all(b == 0xAB for b in guard[C.sizeof(bf.BfConfig):]), "bf_config is larger than its ctypes image"
    assert cfg.device == 0 and cfg.past_windows == 10


def test_getparam_fallbacks_and_launch_values():
    # code fall-backs (mvdr.cpp:155-184) vs launch/mvdr.launch:6-10
    raw = bf.make_config("mvdr", mics="aira3", launch=False)
    assert (raw.past_windows, raw.freq_mag_threshold, raw.freq_min, raw.freq_max, raw.out_amp) == (10, 1.5, 400.0, 4000.0, 4.5)
    lch = bf.make_config("mvdr", mics="aira3")
    assert (lch.freq_mag_threshold, lch.freq_min, lch.freq_max, lch.out_amp) == (0.001, 100.0, 16000.0, 1.0)
    # B-11: phase.launch sets keys phase.cpp never reads -> fall-backs stay
    ph = bf.make_config("phase", mics="aira3")
    assert (ph.mag_mult, ph.mag_threshold, ph.min_phase) == (0.1, 0.05, 10.0)
    # B-11: "MCRA_L = 0.01" fall-back truncates to 0; launch file gives 50
    assert bf.make_config("phasempf", mics="binaural", launch=False).MCRA_L == 0
    assert bf.make_config("phasempf", mics="binaural").MCRA_L == 50


def test_reference_yaml_loads_unchanged(tmp_path):
    # the shipped beamform_config.yaml layout (beamform_config.yaml:1-57): scalars, flow maps, all interferers 181
    y = tmp_path / "beamform_config.yaml"
    y.write_text("verbose: false\ninitial_angle: 15\n# aira3\nmic0: {id: 1, x:  0.000, y:  0.000}\nmic1: {id: 2, x:  0.000, y: -0.180}\n"
                 "mic2: {id: 3, x: -0.156, y: -0.090}\n#mic3: {id: 4, x: 1, y: 1}\nangle_interf1: 40\nangle_interf2: 181\nangle_interf3: -20\n")
    cfg = bf.load_yaml_config("lcmv", str(y))
    assert cfg.n_mics == 3 and cfg.initial_angle == 15.0
    assert [round(cfg.mic_y[i], 3) for i in range(3)] == [0.0, -0.18, -0.09]
    assert cfg.n_angle_interf == 3 and list(cfg.angle_interf[:3]) == [40.0, 181.0, -20.0]   # the |a|>180 cut happens in bf_create (util.h:101-112)


def test_error_paths_never_throw():
    lib = bf.lib()
    cfg = bf.BfConfig()
    assert lib.bf_config_init(C.byref(cfg), 99) == 1           # BF_ERR_INVALID
    assert b"bad arguments" in lib.bf_last_error()
    assert lib.bf_config_load_yaml(C.byref(cfg), b"/nonexistent/beamform_config.yaml") == 5   # BF_ERR_IO
    h = C.c_void_p()
    bad = bf.make_config("das", mics="aira3")
    bad.n_mics = 0
    assert lib.bf_create(C.byref(h), C.byref(bad), 1) == 1 and not h.value
    assert lib.bf_set_theta(None, 1.0) == 1
    assert lib.bf_fft_win(None) == 0
    assert b"sm_100a" in lib.bf_version()


def test_create_without_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    cfg = bf.make_config("das", mics="aira3")
    rc = bf.lib().bf_create(C.byref(h), C.byref(cfg), 1)
    assert rc == 2 and not h.value, "no device must be BF_ERR_NO_DEVICE: there is no CPU fallback"
    with pytest.raises(bf.BeamformError):
        bf.Beamformer(cfg, 1)


def test_plain_config_table_matches_bf_config_init():
    """bench.py's reference arm builds its configuration from beamform_b200/tables.py without loading the product library:
    the table must agree field by field with what bf_config_init + bf_config_set produce."""
    from beamform_b200 import tables
    from oracle_lib import BfoConfig, config_from_fields
    cases = [("das", "aira3", {}), ("mvdr", "circ8", dict(freq_mag_threshold=0.0002)), ("lcmv", "circ8", dict(interferers=(80.0, -60.0, 150.0))),
             ("gss", "circ8", dict(interferers=(80.0, -60.0, 150.0))), ("phase", "aira3", {}), ("phasempf", "binaural", dict(hop=2048)),
             ("mcra", "aira3", {}), ("ref", "aira3", {}), ("gsc", "aira3", {}), ("phasempf", "aira3", dict(launch=False)), ("mcra", "aira3", dict(launch=False))]
    for algo, mics, kw in cases:
        prod = bf.make_config(algo, mics=mics, **kw)
        plain = config_from_fields(tables.plain_config_fields(algo, mics=mics, **kw))
        for name, _ in BfoConfig._fields_:
            a, b = getattr(prod, name), getattr(plain, name)
            if hasattr(a, "__len__"):
                assert list(a) == list(b), (algo, name)
            else:
                assert a == b, (algo, name, a, b)


def test_launch_file_reader_matches_launch_table(tmp_path):
    """bf_config_load_launch reads the inline <rosparam> block of a reference launch file (launch/mvdr.launch:5-11) and <param> tags."""
    from beamform_b200 import tables
    for algo, kv in tables.LAUNCH_PARAMS.items():
        body = "".join("      %s: %s\n" % (k, ("true" if v else "false") if isinstance(v, bool) else v) for k, v in kv.items())
        f = tmp_path / (algo + ".launch")
        f.write_text('<launch>\n  <node name="beamform" pkg="beamform" type="%s" output="screen">\n'
                     '    <rosparam command="load" file="$(find beamform)/beamform_config.yaml" />\n    <rosparam>\n%s    </rosparam>\n  </node>\n</launch>\n' % (algo, body))
        cfg = bf.BfConfig()
        assert bf.lib().bf_config_init(C.byref(cfg), bf.ALGOS[algo]) == 0
        assert bf.lib().bf_config_load_launch(C.byref(cfg), str(f).encode()) == 0
        ref = bf.make_config(algo, mics="aira3")
        for name in ("past_windows", "freq_mag_threshold", "freq_max", "freq_min", "out_amp", "interf_angle_threshold", "mu", "lambda_", "min_phase",
                     "min_mag", "smooth_size", "MCRA_alphaS", "MCRA_alphaD2", "MCRA_L", "MPF_alphaS", "MPF_rev_gamma", "noise_floor", "out_only_noise",
                     "use_vad", "mu0", "mu_max", "filter_size"):
            assert getattr(cfg, name) == getattr(ref, name), (algo, name)
    g = tmp_path / "p.launch"
    g.write_text('<launch><node name="b" pkg="beamform" type="mvdr"><param name="freq_max" value="8000"/><param name="past_windows" value="6" /></node></launch>')
    cfg = bf.BfConfig()
    bf.lib().bf_config_init(C.byref(cfg), 1)
    assert bf.lib().bf_config_load_launch(C.byref(cfg), str(g).encode()) == 0
    assert (cfg.freq_max, cfg.past_windows) == (8000.0, 6)
    assert bf.lib().bf_config_load_launch(C.byref(cfg), b"/nonexistent.launch") != 0


def test_offline_tool_fails_loudly_without_a_gpu(tmp_path):
    """tools/bf_offline parses its arguments, the WAV and the parameter files on any machine; with no B200 bf_create reports
    BF_ERR_NO_DEVICE and the tool exits 1 with the message (no CPU fallback)."""
    import subprocess
    import numpy as np
    from scipy.io import wavfile
    from beamform_b200 import build
    exe = build.build_tools()
    y = tmp_path / "c.yaml"
    y.write_text("initial_angle: 0\nmic0: {id: 1, x: 0.0, y: 0.0}\nmic1: {id: 2, x: 0.0, y: -0.18}\n")
    wavfile.write(str(tmp_path / "in.wav"), 48000, np.zeros((2048, 2), dtype=np.int16))
    r = subprocess.run([exe, "--algo", "das", "--config", str(y), "--in", str(tmp_path / "in.wav"), "--out", str(tmp_path / "o.wav")], capture_output=True, text=True)
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0
    else:
        assert r.returncode == 1 and "no CUDA device" in r.stderr
    r = subprocess.run([exe, "--algo", "das", "--config", str(y), "--in", str(y), "--out", str(tmp_path / "o.wav")], capture_output=True, text=True)
    assert r.returncode == 1 and "RIFF" in r.stderr
