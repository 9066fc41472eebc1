"""Independent numpy restatement of das / mvdr / lcmv (TEST INFRASTRUCTURE).

Purpose: cross-check the THIRD-PARTY arithmetic the oracle and oracle/shim restate — FFTW's c2c DFT and Eigen's
PartialPivLU inverse — against independent implementations of the same published definitions: numpy.fft
(pocketfft, double) and numpy.linalg.inv (LAPACK zgetrf/zgetri = partial-pivot LU).  Written from SURVEY.md
Appendix A, not from the oracle's code.  Citations: /root/reference/beamform/src/."""
import numpy as np

V = 343.0


def freqs_vec(N, sr):   # util.h:190-199 incl. quirks B-1/B-2
    f = np.zeros(N)
    i = np.arange(N // 2 - 1)
    f[i + 1] = (i + 1) / N * sr
    f[N - 1 - i] = -((i + 1) / N) * sr
    f[N // 2 - 1] = sr / 2
    return f


def delays(xy, theta):   # util.h:82-92,136-161
    xy = np.asarray(xy, dtype=np.float64)
    dist = np.hypot(xy[:, 0], xy[:, 1])
    ang = np.degrees(np.arctan2(xy[:, 1], xy[:, 0]))
    d = ang - theta
    d = np.where(d > 180, d - 360, np.where(d < -180, d + 360, d))
    tau = dist * np.cos(np.radians(d)) / (-V)
    tau[0] = 0.0
    return tau


def steering(xy, theta, f):   # das.cpp:40-42; row 0 == 1
    w = np.exp(-1j * 2 * np.pi * f[None, :] * delays(xy, theta)[:, None])
    w[0, :] = 1.0
    return w


def run(algo, xy, x, hop=512, sr=48000, theta=0.0, interferers=(), past_windows=10, thr=0.001, fmin=100.0, fmax=16000.0, out_amp=1.0):
    """x [M][L] float32 -> [L] float32.  algo in {"das", "mvdr", "lcmv"}; no control events."""
    M, L = x.shape
    N = 2 * hop
    T = L // hop
    win = np.sqrt(0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N))   # util.h:201-211
    f = freqs_vec(N, sr)
    C = np.stack([steering(xy, a, f) for a in (theta,) + tuple(interferers)], axis=2)   # [M][N][K+1]
    white = np.ones((M, M)) + 0.001 * np.eye(M)                        # mvdr.cpp:239-243
    inband = (np.abs(f) >= fmin) & (np.abs(f) <= fmax)
    hist = np.zeros((N, M, past_windows), dtype=np.complex128)
    xin = np.concatenate([np.zeros((M, hop), dtype=np.float32), x], axis=1)
    prev = np.zeros(N, dtype=np.float32)
    out = np.empty(L, dtype=np.float32)
    for t in range(T):
        frame = xin[:, t * hop:t * hop + N].astype(np.float64) * win[None, :]
        X = np.fft.fft(frame, axis=1)
        if algo == "das":
            Y = np.einsum("mj,mj->j", np.conj(C[:, :, 0]), X) / M
        else:
            Y = np.zeros(N, dtype=np.complex128)
            stat = np.abs(X).sum(axis=0) / (M * N)
            j0 = 1 if algo == "mvdr" else 0
            if algo == "mvdr":
                Y[0] = X[0, 0]
            for j in range(j0, N):
                if not inband[j]:
                    continue
                if stat[j] > thr:
                    P = hist[j]
                    R = (P @ P.conj().T) * white
                    with np.errstate(all="ignore"):
                        try:
                            Ri = np.linalg.inv(R)
                        except np.linalg.LinAlgError:
                            Ri = np.full((M, M), np.nan + 0j)
                        if algo == "mvdr":
                            d = C[:, j, 0]
                            w = (Ri @ d) / (d.conj() @ Ri @ d)
                        else:
                            Cj = C[:, j, :]
                            w = ((Ri @ Cj) @ np.linalg.inv(Cj.conj().T @ Ri @ Cj))[:, 0]
                        Y[j] = w.conj() @ X[:, j]
                else:
                    Y[j] = 0.01 * X[0, j]
                hist[j, :, :-1] = hist[j, :, 1:]
                hist[j, :, -1] = X[:, j]
        y = np.fft.ifft(Y) * N
        cur = (np.real(y) / N).astype(np.float32)
        cur = (cur * win).astype(np.float32)
        if algo != "das":
            cur = (cur * out_amp).astype(np.float32)
        out[t * hop:(t + 1) * hop] = prev[hop:] + cur[:hop]
        prev = cur
    return out
