"""ctypes binding of the CPU oracle (oracle/libbf_oracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIBPATH = os.path.join(ORACLE_DIR, "libbf_oracle.so")
MAX_MICS, MAX_INTERF = 64, 16


class BfoConfig(C.Structure):
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
        ("use_vad", C.c_int32), ("vad_threshold", C.c_double), ("mu0", C.c_double), ("mu_max", C.c_double), ("filter_size", C.c_int32),
    ]


class BfoEvent(C.Structure):
    _fields_ = [("hop_index", C.c_uint32), ("kind", C.c_int32), ("id", C.c_uint32), ("value", C.c_float)]


_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH) or os.path.getmtime(LIBPATH) < max(
                os.path.getmtime(os.path.join(ORACLE_DIR, f)) for f in ("bf_oracle.hpp", "bf_oracle_capi.cpp")):
            build()
        L = C.CDLL(LIBPATH)
        L.bfo_create.restype = C.c_void_p
        L.bfo_create.argtypes = [C.POINTER(BfoConfig)]
        L.bfo_destroy.argtypes = [C.c_void_p]
        L.bfo_fft_win.argtypes = [C.c_void_p]
        L.bfo_fft_win.restype = C.c_uint32
        L.bfo_set_theta.argtypes = [C.c_void_p, C.c_float]
        L.bfo_set_interference.argtypes = [C.c_void_p, C.c_uint16, C.c_float]
        L.bfo_get_interferences.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        L.bfo_get_freqs.argtypes = [C.c_void_p, C.c_void_p]
        L.bfo_get_delays.argtypes = [C.c_void_p, C.c_void_p]
        L.bfo_get_window.argtypes = [C.c_void_p, C.c_void_p]
        L.bfo_get_weights.argtypes = [C.c_void_p, C.c_void_p]
        L.bfo_process_hop.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_float), C.c_uint32]
        L.bfo_process.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.POINTER(BfoEvent), C.c_int,
                                  C.c_int, C.c_void_p, C.c_void_p]
        L.bfo_srp.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def from_product_config(cfg):
    """Copy every field the oracle knows from a beamform_b200.BfConfig (same names by construction)."""
    o = BfoConfig()
    for name, _ in BfoConfig._fields_:
        setattr(o, name, getattr(cfg, name))
    return o


def config_from_fields(fields):
    """BfoConfig from the plain dict of beamform_b200/tables.py:plain_config_fields (no product library involved)."""
    o = BfoConfig()
    for name, _ in BfoConfig._fields_:
        v = fields[name]
        if isinstance(v, (list, tuple)):
            arr = getattr(o, name)
            for i, a in enumerate(v):
                arr[i] = a
        else:
            setattr(o, name, v)
    return o


class Oracle:
    def __init__(self, cfg):
        self.cfg = cfg if isinstance(cfg, BfoConfig) else from_product_config(cfg)
        self.h = lib().bfo_create(C.byref(self.cfg))
        self.N = lib().bfo_fft_win(self.h)
        self.H = self.N // 2
        self.M = self.cfg.n_mics

    def __del__(self):
        try:
            lib().bfo_destroy(self.h)
        except Exception:
            pass

    def set_theta(self, a):
        lib().bfo_set_theta(self.h, a)

    def set_interference(self, id, a):
        return lib().bfo_set_interference(self.h, id, a)

    @property
    def interferences(self):
        buf = (C.c_double * 64)()
        n = lib().bfo_get_interferences(self.h, buf, 64)
        return [buf[i] for i in range(n)]

    def freqs(self):
        out = np.empty(self.N)
        lib().bfo_get_freqs(self.h, out.ctypes.data)
        return out

    def delays(self):
        out = np.empty(self.M)
        lib().bfo_get_delays(self.h, out.ctypes.data)
        return out

    def window(self):
        out = np.empty(self.N)
        lib().bfo_get_window(self.h, out.ctypes.data)
        return out

    def weights(self):
        Cn = lib().bfo_get_weights(self.h, None)
        out = np.empty((self.N, self.M, Cn, 2))
        lib().bfo_get_weights(self.h, out.ctypes.data)
        return out[..., 0] + 1j * out[..., 1]

    def process_hop(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        ptrs = (C.POINTER(C.c_float) * self.M)(*[x[m].ctypes.data_as(C.POINTER(C.c_float)) for m in range(self.M)])
        out = np.empty(self.H, dtype=np.float32)
        lib().bfo_process_hop(self.h, ptrs, out.ctypes.data_as(C.POINTER(C.c_float)), self.H)
        return out

    def process(self, x, events=(), dropped_hops=0, want_flags=False):
        """x [M][L] float32 -> out [L] (and per-hop selection / mask flags [T][N] when asked)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        M, L = x.shape
        T = L // self.H
        out = np.empty(L, dtype=np.float32)
        evs = sorted(events, key=lambda e: e[0])
        arr = (BfoEvent * max(1, len(evs)))()
        for i, e in enumerate(evs):
            arr[i] = BfoEvent(int(e[0]), 0, 0, float(e[2])) if e[1] == "theta" else BfoEvent(int(e[0]), 1, int(e[2]), float(e[3]))
        sel = np.zeros((T, self.N), dtype=np.uint8) if want_flags else None
        msk = np.zeros((T, self.N), dtype=np.uint8) if want_flags else None
        lib().bfo_process(self.h, x.ctypes.data, L, out.ctypes.data, T, arr, len(evs), dropped_hops,
                          sel.ctypes.data if want_flags else None, msk.ctypes.data if want_flags else None)
        return (out, sel, msk) if want_flags else out

    def srp(self, x, thetas):
        x = np.ascontiguousarray(x, dtype=np.float32)
        M, L = x.shape
        T = L // self.H
        th = np.ascontiguousarray(thetas, dtype=np.float64)
        maps = np.empty((T, len(th)))
        lib().bfo_srp(self.h, x.ctypes.data, L, T, th.ctypes.data, len(th), maps.ctypes.data)
        return maps
