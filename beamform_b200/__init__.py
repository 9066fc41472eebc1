"""beamform_b200 — B200-native frequency-domain beamforming behind balkce/beamform's node interface.

The product is the CUDA library `libbeamform_b200.so` (C ABI in include/beamform_b200.h).  This
package is the thin Python host mirror used by the tests, the bench and the offline driver: the same
names and argument meanings as the reference nodes (jack_callback -> process_hop, /theta ->
set_theta, /theta_interference -> set_interference, rosparam keys -> Config fields).

There is no CPU fallback: loading fails loudly when the extension has not been built, and
bf_create fails when no sm_100 device is present.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

from .tables import ALGOS
MAX_MICS = 64
MAX_INTERF = 16


class BfConfig(C.Structure):
    """ctypes image of `bf_config` (include/beamform_b200.h)."""
    _fields_ = [
        ("algo", C.c_int32), ("sample_rate", C.c_double), ("hop", C.c_uint32), ("n_mics", C.c_int32),
        ("mic_x", C.c_double * MAX_MICS), ("mic_y", C.c_double * MAX_MICS), ("initial_angle", C.c_double),
        ("n_angle_interf", C.c_int32), ("angle_interf", C.c_double * MAX_INTERF),
        ("past_windows", C.c_uint32), ("freq_mag_threshold", C.c_double), ("freq_max", C.c_double),
        ("freq_min", C.c_double), ("out_amp", C.c_double), ("interf_angle_threshold", C.c_double),
        ("mu", C.c_double), ("lambda_", C.c_double),
        ("min_phase", C.c_double), ("mag_mult", C.c_double), ("mag_threshold", C.c_double),
        ("min_mag", C.c_double), ("smooth_size", C.c_int32),
        ("MCRA_alphaS", C.c_double), ("MCRA_alphaD", C.c_double), ("MCRA_alphaD2", C.c_double), ("MCRA_delta", C.c_double),
        ("MCRA_L", C.c_int32),
        ("MPF_alphaS", C.c_double), ("MPF_eta", C.c_double), ("MPF_rev_gamma", C.c_double), ("MPF_rev_delta", C.c_double),
        ("noise_floor", C.c_double), ("out_only_noise", C.c_int32), ("out_only_mcra", C.c_int32),
        ("dropped_hops_on_restructure", C.c_int32), ("device", C.c_int32),
        ("use_vad", C.c_int32), ("vad_threshold", C.c_double), ("mu0", C.c_double), ("mu_max", C.c_double), ("filter_size", C.c_int32),
    ]


class BfEvent(C.Structure):
    _fields_ = [("hop_index", C.c_uint32), ("kind", C.c_int32), ("id", C.c_uint32), ("value", C.c_float)]


from .tables import GEOMETRIES, LAUNCH_PARAMS  # noqa: F401  (pure data: launch-file blocks, geometries)

_lib = None


def lib():
    """Load libbeamform_b200.so (building it in-tree if the sources are newer). Never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        raise RuntimeError("libbeamform_b200.so is not built: run `python -m beamform_b200.build` (there is no CPU fallback)")
    L = C.CDLL(path)
    P = C.POINTER
    L.bf_config_init.argtypes = [P(BfConfig), C.c_int]
    L.bf_config_load_yaml.argtypes = [P(BfConfig), C.c_char_p]
    L.bf_config_set.argtypes = [P(BfConfig), C.c_char_p, C.c_char_p]
    L.bf_config_load_launch.argtypes = [P(BfConfig), C.c_char_p]
    L.bf_create.argtypes = [P(C.c_void_p), P(BfConfig), C.c_uint32]
    L.bf_destroy.argtypes = [C.c_void_p]
    L.bf_destroy.restype = None
    L.bf_set_theta.argtypes = [C.c_void_p, C.c_float]
    L.bf_set_interference.argtypes = [C.c_void_p, C.c_uint16, C.c_float]
    L.bf_get_theta.argtypes = [C.c_void_p, P(C.c_double)]
    L.bf_get_interferences.argtypes = [C.c_void_p, P(C.c_double), C.c_uint32, P(C.c_uint32)]
    L.bf_process_hop.argtypes = [C.c_void_p, P(P(C.c_float)), P(C.c_float), C.c_uint32]
    L.bf_apply_weights.argtypes = [C.c_void_p, P(P(C.c_float)), P(C.c_float), C.c_uint32]
    L.bf_process_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32,
                                   P(BfEvent), C.c_uint32]
    L.bf_process_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32,
                                          P(BfEvent), C.c_uint32, C.c_void_p]
    L.bf_set_capture.argtypes = [C.c_void_p, C.c_void_p]
    L.bf_srp_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, P(C.c_float), C.c_uint32, C.c_void_p,
                                      C.c_uint32, C.c_void_p]
    L.bf_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.bf_get_profile.argtypes = [C.c_void_p, P(C.c_double), P(C.c_uint64)]
    L.bf_fft_win.argtypes = [C.c_void_p]
    L.bf_fft_win.restype = C.c_uint32
    L.bf_kernel_launches.argtypes = [C.c_void_p]
    L.bf_kernel_launches.restype = C.c_uint64
    L.bf_last_error.restype = C.c_char_p
    L.bf_version.restype = C.c_char_p
    _lib = L
    return L


class BeamformError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise BeamformError("%s failed (status %d): %s" % (what, rc, lib().bf_last_error().decode()))


def make_config(algo, mics="aira3", hop=512, sample_rate=48000, initial_angle=0.0, interferers=(), launch=True, **params):
    """bf_config for node `algo`: getParam fall-backs, then the launch-file <rosparam> block, then overrides."""
    cfg = BfConfig()
    _check(lib().bf_config_init(C.byref(cfg), ALGOS[algo]), "bf_config_init")
    cfg.hop = hop
    cfg.sample_rate = sample_rate
    cfg.initial_angle = initial_angle
    xy = GEOMETRIES[mics] if isinstance(mics, str) else list(mics)
    cfg.n_mics = len(xy)
    for i, (x, y) in enumerate(xy):
        cfg.mic_x[i], cfg.mic_y[i] = x, y
    cfg.n_angle_interf = len(interferers)
    for i, a in enumerate(interferers):
        cfg.angle_interf[i] = a
    kv = dict(LAUNCH_PARAMS[algo]) if launch else {}
    kv.update(params)
    for k, v in kv.items():
        sv = ("true" if v else "false") if isinstance(v, bool) else repr(float(v))
        _check(lib().bf_config_set(C.byref(cfg), k.encode(), sv.encode()), "bf_config_set")
    return cfg


def load_yaml_config(algo, path, hop=512, sample_rate=48000, **params):
    """beamform_config.yaml (unchanged) + launch-file parameters for node `algo`."""
    cfg = BfConfig()
    _check(lib().bf_config_init(C.byref(cfg), ALGOS[algo]), "bf_config_init")
    cfg.hop, cfg.sample_rate = hop, sample_rate
    _check(lib().bf_config_load_yaml(C.byref(cfg), os.fsencode(path)), "bf_config_load_yaml")
    kv = dict(LAUNCH_PARAMS[algo])
    kv.update(params)
    for k, v in kv.items():
        sv = ("true" if v else "false") if isinstance(v, bool) else repr(float(v))
        _check(lib().bf_config_set(C.byref(cfg), k.encode(), sv.encode()), "bf_config_set")
    return cfg


def make_events(events):
    """[(hop, 'theta', angle) | (hop, 'interf', id, angle)] -> (BfEvent array, n)."""
    arr = (BfEvent * max(1, len(events)))()
    for i, e in enumerate(sorted(events, key=lambda e: e[0])):
        if e[1] == "theta":
            arr[i] = BfEvent(int(e[0]), 0, 0, float(e[2]))
        else:
            arr[i] = BfEvent(int(e[0]), 1, int(e[2]), float(e[3]))
    return arr, len(events)


class Beamformer:
    """One reference node (das/mvdr/lcmv/gss/phase/phasempf) running on a B200, n_streams independent streams."""

    def __init__(self, cfg, n_streams=1):
        self.cfg = cfg
        self.n_streams = n_streams
        self._h = C.c_void_p()
        _check(lib().bf_create(C.byref(self._h), C.byref(cfg), n_streams), "bf_create")
        self.hop = cfg.hop
        self.n_mics = cfg.n_mics
        self.fft_win = lib().bf_fft_win(self._h)

    def close(self):
        if self._h:
            lib().bf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- control topics -------------------------------------------------------------------------
    def set_theta(self, angle_deg):
        _check(lib().bf_set_theta(self._h, angle_deg), "bf_set_theta")

    def set_interference(self, id, angle_deg):
        _check(lib().bf_set_interference(self._h, id, angle_deg), "bf_set_interference")

    @property
    def theta(self):
        a = C.c_double()
        _check(lib().bf_get_theta(self._h, C.byref(a)), "bf_get_theta")
        return a.value

    @property
    def interferences(self):
        buf = (C.c_double * MAX_INTERF)()
        n = C.c_uint32()
        _check(lib().bf_get_interferences(self._h, buf, MAX_INTERF, C.byref(n)), "bf_get_interferences")
        return [buf[i] for i in range(n.value)]

    # --- the process callback -------------------------------------------------------------------
    def process_hop(self, in_hop):
        """jack_callback: in_hop [M][hop] float32 -> [hop] float32 (one hop of latency)."""
        x = np.ascontiguousarray(in_hop, dtype=np.float32)
        assert x.shape == (self.n_mics, self.hop)
        ptrs = (C.POINTER(C.c_float) * self.n_mics)(*[x[m].ctypes.data_as(C.POINTER(C.c_float)) for m in range(self.n_mics)])
        out = np.empty(self.hop, dtype=np.float32)
        _check(lib().bf_process_hop(self._h, ptrs, out.ctypes.data_as(C.POINTER(C.c_float)), self.hop), "bf_process_hop")
        return out

    def apply_weights(self, frames):
        """The weight_func seam (util.h:289): frames [M][fft_win] float32 -> fft_win windowed output samples, before overlap-add."""
        x = np.ascontiguousarray(frames, dtype=np.float32)
        assert x.shape == (self.n_mics, self.fft_win)
        ptrs = (C.POINTER(C.c_float) * self.n_mics)(*[x[m].ctypes.data_as(C.POINTER(C.c_float)) for m in range(self.n_mics)])
        out = np.empty(self.fft_win, dtype=np.float32)
        _check(lib().bf_apply_weights(self._h, ptrs, out.ctypes.data_as(C.POINTER(C.c_float)), self.fft_win), "bf_apply_weights")
        return out

    def process(self, x, events=()):
        """Offline batch through host buffers: x [B][M][L] float32 -> [B][L] float32."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        B, M, L = x.shape
        assert B == self.n_streams and M == self.n_mics and L % self.hop == 0
        out = np.empty((B, L), dtype=np.float32)
        ev, nev = make_events(list(events))
        _check(lib().bf_process_batch(self._h, x.ctypes.data, M * L, L, out.ctypes.data, L, L // self.hop, ev, nev), "bf_process_batch")
        return out

    def process_device(self, in_ptr, out_ptr, n_hops, stream_ptr=0, in_stream_stride=None, in_mic_stride=None,
                       out_stream_stride=None, events=()):
        """Device-resident batch: raw device pointers (e.g. torch .data_ptr()), dense [B][M][L] / [B][L] by default."""
        L = n_hops * self.hop
        ev, nev = make_events(list(events))
        _check(lib().bf_process_batch_device(self._h, in_ptr, in_stream_stride or self.n_mics * L, in_mic_stride or L, out_ptr,
                                             out_stream_stride or L, n_hops, ev, nev, stream_ptr), "bf_process_batch_device")

    def srp_device(self, in_ptr, thetas_deg, maps_ptr, n_hops, stream_ptr=0, in_stream_stride=None, in_mic_stride=None):
        """Steered-response power maps [B][n_hops][len(thetas)] float32 (device) for look directions `thetas_deg`."""
        L = n_hops * self.hop
        th = np.ascontiguousarray(thetas_deg, dtype=np.float32)
        _check(lib().bf_srp_batch_device(self._h, in_ptr, in_stream_stride or self.n_mics * L, in_mic_stride or L,
                                         th.ctypes.data_as(C.POINTER(C.c_float)), len(th), maps_ptr, n_hops, stream_ptr), "bf_srp_batch_device")

    def set_capture(self, dev_ptr):
        _check(lib().bf_set_capture(self._h, dev_ptr), "bf_set_capture")

    def set_profiling(self, on=True):
        _check(lib().bf_set_profiling(self._h, 1 if on else 0), "bf_set_profiling")

    def get_profile(self):
        """(summed device ms of the fused frames kernel, launches) since the last call."""
        ms, n = C.c_double(), C.c_uint64()
        _check(lib().bf_get_profile(self._h, C.byref(ms), C.byref(n)), "bf_get_profile")
        return ms.value, int(n.value)

    @property
    def kernel_launches(self):
        return int(lib().bf_kernel_launches(self._h))
