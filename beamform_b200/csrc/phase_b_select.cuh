// PHASE B for the magnitude-gated nodes: mvdr.cpp:62-115, lcmv.cpp:88-140, gss.cpp:96-156.
//
// Per frame and logical bin: band test (host table), magnitude gate sum_i |X_i| / (M*N) > thr,
// then for selected bins a small per-bin linear-algebra problem, for the others 0.01*X_0 / 0.
// Work is organised in CTA-wide steps so lanes stay busy although only ~20-30 % of bins pass the gate:
//   B1  every (bin, frame): even/odd separation of the packed spectra, gate in FP32
//       (+ a guard band around the threshold; guarded bins are re-decided in FP64 by B1b)
//   B1b FP64 re-decision (exact double DFT of that one bin, warp-cooperative) -> bit-exact selection
//   B2  compacted list of selected items, one thread per item: covariance + Cholesky solves (mvdr/lcmv)
//       or the W recursion (gss)
//   B3  history-ring update, Hermitian assembly of G = Yh_t + i*Yh_{t+1}, diagnostics
#pragma once
#include "bf_device.h"
#include "warp_fft1024.cuh"

namespace bf {

constexpr int kL1K = 514;   // logical bins for N = 1024: 0..512 and the pseudo-bin 513
constexpr int kMaxMSel = 8; // register-resident solves
constexpr int kMaxC = 8;     // columns of the constraint matrix on the register-resident kernels (look direction + 7 interferers)
constexpr int kMaxCGen = 16; // general gated kernel: look direction + the 15 interferer slots beamform_config.yaml ships

struct SelScratch {
  float2 y[2][kL1K];
  float sqrtE[2][BF_MAX_MICS_DEV];
  unsigned short items[2 * kL1K];
  unsigned short recheck[2 * kL1K];
  int n_items, n_recheck;
  unsigned char flag[2][kL1K];   // bit0: selected (gate passed inside the band)
};

// X_i[j] of frame f (0: t, 1: t+1) from the packed half-scaled spectrum Z = FFT(0.5*w*(x_t + i x_{t+1})).
__device__ __forceinline__ float2 unpack_bin(const float2* z, int l, int f) {
  const int j = (l == kL1K - 1) ? 511 : l;   // pseudo-bin: conj of bin N/2-1
  const float2 a = z[j], b = z[(1024 - j) & 1023];
  float2 x;
  if (f == 0) x = make_float2(a.x + b.x, a.y - b.y);        // Z[j] + conj(Z[N-j])
  else x = make_float2(a.y + b.y, b.x - a.x);               // -i (Z[j] - conj(Z[N-j]))
  if (l == kL1K - 1) x.y = -x.y;
  return x;
}

template <typename T>
struct cplx { T x, y; };
template <typename T>
__device__ __forceinline__ cplx<T> mk(T x, T y) { cplx<T> r; r.x = x; r.y = y; return r; }
template <typename T>
__device__ __forceinline__ T fma_t(T a, T b, T c) { return fma(a, b, c); }
template <>
__device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }

template <int MM, typename T>
struct HermLower {   // lower triangle of an MM x MM Hermitian matrix, registers
  T dg[MM];
  cplx<T> lo[MM * (MM - 1) / 2 > 0 ? MM * (MM - 1) / 2 : 1];
  __device__ __forceinline__ static constexpr int idx(int i, int j) { return i * (i - 1) / 2 + j; }
};

// R = (P P^H) .* whiteR over the last P frames of this bin (mvdr.cpp:87, :239-243), Cholesky R = L L^H
// in place.  xprev substitutes the ring slot of frame t when solving for frame t+1 of the same pair.
// T = float for mvdr (measured rel-L2 3e-5 against the oracle), double for lcmv: the constrained solve
// on a rank-deficient cold-start history, or with the zero row 0 left behind by an interference-list
// restructure (SURVEY B-8), needs more than FP32 to stay inside 1e-4.
template <int MM, typename T>
__device__ __forceinline__ void build_cov_chol(const KernelParams& p, HermLower<MM, T>& A, T (&invd)[MM], const float2* ring,
                                               int subst_pos, const float2 (&xprev)[MM]) {
  typedef HermLower<MM, T> HL;
  const int M = p.M;
#pragma unroll
  for (int i = 0; i < MM; i++) A.dg[i] = T(0);
#pragma unroll
  for (int i = 0; i < MM * (MM - 1) / 2; i++) A.lo[i] = mk<T>(T(0), T(0));
  for (int pos = 0; pos < p.P; pos++) {
    cplx<T> h[MM];
    if (pos == subst_pos) {
#pragma unroll
      for (int i = 0; i < MM; i++) h[i] = mk<T>((T)xprev[i].x, (T)xprev[i].y);
    } else {
      const float2* src = ring + (size_t)pos * M;
#pragma unroll
      for (int i = 0; i < MM; i++) {
        const float2 v = (i < M) ? src[i] : make_float2(0.f, 0.f);
        h[i] = mk<T>((T)v.x, (T)v.y);
      }
    }
#pragma unroll
    for (int i = 0; i < MM; i++) {
      A.dg[i] = fma_t<T>(h[i].x, h[i].x, fma_t<T>(h[i].y, h[i].y, A.dg[i]));
#pragma unroll
      for (int j = 0; j < i; j++) {   // h_i * conj(h_j)
        cplx<T>& r = A.lo[HL::idx(i, j)];
        r.x = fma_t<T>(h[i].x, h[j].x, fma_t<T>(h[i].y, h[j].y, r.x));
        r.y = fma_t<T>(h[i].y, h[j].x, fma_t<T>(-h[i].x, h[j].y, r.y));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < MM; j++) {
    if (j < M) {
      T d = A.dg[j] * T(1.001);   // whiteR diagonal (mvdr.cpp:242)
#pragma unroll
      for (int k = 0; k < j; k++) { const cplx<T> l = A.lo[HL::idx(j, k)]; d = fma_t<T>(-l.x, l.x, fma_t<T>(-l.y, l.y, d)); }
      const T ljj = sqrt(d);
      const T inv = T(1) / ljj;
      A.dg[j] = ljj;
      invd[j] = inv;
#pragma unroll
      for (int i = j + 1; i < MM; i++) {
        if (i < M) {
          cplx<T> acc = A.lo[HL::idx(i, j)];
#pragma unroll
          for (int k = 0; k < j; k++) {   // acc -= L[i][k] * conj(L[j][k])
            const cplx<T> a = A.lo[HL::idx(i, k)], b = A.lo[HL::idx(j, k)];
            acc.x = fma_t<T>(-a.x, b.x, fma_t<T>(-a.y, b.y, acc.x));
            acc.y = fma_t<T>(-a.y, b.x, fma_t<T>(a.x, b.y, acc.y));
          }
          A.lo[HL::idx(i, j)] = mk<T>(acc.x * inv, acc.y * inv);
        }
      }
    } else {
      invd[j] = T(0);
    }
  }
}

// v <- L^{-1} v (forward substitution), v in registers
template <int MM, typename T>
__device__ __forceinline__ void fwd_solve(const KernelParams& p, const HermLower<MM, T>& A, const T (&invd)[MM], cplx<T> (&v)[MM]) {
  typedef HermLower<MM, T> HL;
#pragma unroll
  for (int i = 0; i < MM; i++) {
    if (i < p.M) {
      cplx<T> acc = v[i];
#pragma unroll
      for (int k = 0; k < i; k++) {
        const cplx<T> l = A.lo[HL::idx(i, k)];
        acc.x = fma_t<T>(-l.x, v[k].x, fma_t<T>(l.y, v[k].y, acc.x));
        acc.y = fma_t<T>(-l.x, v[k].y, fma_t<T>(-l.y, v[k].x, acc.y));
      }
      v[i] = mk<T>(acc.x * invd[i], acc.y * invd[i]);
    } else {
      v[i] = mk<T>(T(0), T(0));
    }
  }
}

template <int MM, typename T>
__device__ __forceinline__ cplx<T> cdot_conj(const cplx<T> (&a)[MM], const cplx<T> (&b)[MM]) {   // sum conj(a_i) b_i
  cplx<T> s = mk<T>(T(0), T(0));
#pragma unroll
  for (int i = 0; i < MM; i++) {
    s.x = fma_t<T>(a[i].x, b[i].x, fma_t<T>(a[i].y, b[i].y, s.x));
    s.y = fma_t<T>(a[i].x, b[i].y, fma_t<T>(-a[i].y, b[i].x, s.y));
  }
  return s;
}

// mvdr.cpp:86-94: w = R^{-1} d / (d^H R^{-1} d), y = w^H x.  With R = L L^H, z = L^{-1} d, u = L^{-1} x:
// y = (z^H u) / (z^H z).
template <int MM, typename T>
__device__ __forceinline__ float2 mvdr_item(const KernelParams& p, const float2* ring, int subst_pos, const float2 (&xprev)[MM],
                                            const float2 (&x)[MM], const float2* steer_l) {
  HermLower<MM, T> A;
  T invd[MM];
  build_cov_chol<MM, T>(p, A, invd, ring, subst_pos, xprev);
  cplx<T> z[MM], u[MM];
#pragma unroll
  for (int i = 0; i < MM; i++) {
    const float2 d = (i < p.M) ? steer_l[i] : make_float2(0.f, 0.f);
    z[i] = mk<T>((T)d.x, (T)d.y);
    u[i] = mk<T>((T)x[i].x, (T)x[i].y);
  }
  fwd_solve<MM, T>(p, A, invd, z);
  fwd_solve<MM, T>(p, A, invd, u);
  const cplx<T> num = cdot_conj<MM, T>(z, u);
  const T den = cdot_conj<MM, T>(z, z).x;
  return make_float2((float)(num.x / den), (float)(num.y / den));
}

// lcmv.cpp:111-119: W = R^{-1} C (C^H R^{-1} C)^{-1}, y = W(:,0)^H x.  With V = L^{-1} C, u = L^{-1} x,
// G = V^H V, b = V^H u:  y = g^H b  where  G g = e_0.
template <int MM, typename T>
__device__ __forceinline__ float2 lcmv_item(const KernelParams& p, const float2* ring, int subst_pos, const float2 (&xprev)[MM],
                                            const float2 (&x)[MM], const float2* steer_l) {
  HermLower<MM, T> A;
  T invd[MM];
  build_cov_chol<MM, T>(p, A, invd, ring, subst_pos, xprev);
  const int C = p.C, M = p.M;
  cplx<T> u[MM];
#pragma unroll
  for (int i = 0; i < MM; i++) u[i] = mk<T>((T)x[i].x, (T)x[i].y);
  fwd_solve<MM, T>(p, A, invd, u);
  cplx<T> V[kMaxC][MM];      // dynamically indexed by column: lives in local memory (L1), small
  cplx<T> G[kMaxC][kMaxC];   // lower triangle used
  cplx<T> b[kMaxC];
  for (int c = 0; c < C; c++) {
    cplx<T> v[MM];
#pragma unroll
    for (int i = 0; i < MM; i++) {
      const float2 a = (i < M) ? steer_l[(size_t)c * M + i] : make_float2(0.f, 0.f);
      v[i] = mk<T>((T)a.x, (T)a.y);
    }
    fwd_solve<MM, T>(p, A, invd, v);
#pragma unroll
    for (int i = 0; i < MM; i++) V[c][i] = v[i];
    b[c] = cdot_conj<MM, T>(v, u);
    for (int c2 = 0; c2 <= c; c2++) {   // G[c][c2] = v_c^H v_c2
      cplx<T> w[MM];
#pragma unroll
      for (int i = 0; i < MM; i++) w[i] = V[c2][i];
      G[c][c2] = cdot_conj<MM, T>(v, w);
    }
  }
  // Cholesky of G (C x C), solve G g = e0:  g = Lg^{-H} (Lg^{-1} e0)
  T gd[kMaxC];
  for (int j = 0; j < C; j++) {
    T d = G[j][j].x;
    for (int k = 0; k < j; k++) d -= G[j][k].x * G[j][k].x + G[j][k].y * G[j][k].y;
    const T ljj = sqrt(d);
    gd[j] = T(1) / ljj;
    for (int i = j + 1; i < C; i++) {
      cplx<T> acc = G[i][j];
      for (int k = 0; k < j; k++) {
        const cplx<T> a = G[i][k], bb = G[j][k];
        acc.x -= a.x * bb.x + a.y * bb.y;
        acc.y -= a.y * bb.x - a.x * bb.y;
      }
      G[i][j] = mk<T>(acc.x * gd[j], acc.y * gd[j]);
    }
  }
  cplx<T> q[kMaxC];   // q = Lg^{-1} e0
  for (int i = 0; i < C; i++) {
    cplx<T> acc = mk<T>(i == 0 ? T(1) : T(0), T(0));
    for (int k = 0; k < i; k++) {
      const cplx<T> l = G[i][k];
      acc.x -= l.x * q[k].x - l.y * q[k].y;
      acc.y -= l.x * q[k].y + l.y * q[k].x;
    }
    q[i] = mk<T>(acc.x * gd[i], acc.y * gd[i]);
  }
  cplx<T> g[kMaxC];   // g = Lg^{-H} q (back substitution with the conj-transposed lower factor)
  for (int i = C - 1; i >= 0; i--) {
    cplx<T> acc = q[i];
    for (int k = i + 1; k < C; k++) {   // conj(Lg[k][i]) * g[k]
      const cplx<T> l = G[k][i];
      acc.x -= l.x * g[k].x + l.y * g[k].y;
      acc.y -= l.x * g[k].y - l.y * g[k].x;
    }
    g[i] = mk<T>(acc.x * gd[i], acc.y * gd[i]);
  }
  cplx<T> y = mk<T>(T(0), T(0));   // y = g^H b
  for (int c = 0; c < C; c++) {
    y.x += g[c].x * b[c].x + g[c].y * b[c].y;
    y.y += g[c].x * b[c].y - g[c].y * b[c].x;
  }
  return make_float2((float)y.x, (float)y.y);
}

// gss.cpp:118-137 for one selected frame of one bin; W (C x M) lives in global memory (L2-resident between
// frames), element (c, i) at Wg[(c*M + i) * ws]: the bin index is the fastest axis of the state array so that
// neighbouring threads (= neighbouring bins) touch neighbouring addresses.  Returns y_0.
template <int MAXC = kMaxC>
__device__ __forceinline__ float2 gss_item(const KernelParams& p, float2* Wg, size_t ws, const float2* x, const float2* steer_l) {
  const int C = p.C, M = p.M;
  float2 y[MAXC];
  float alpha = 0.f;
  for (int i = 0; i < M; i++) alpha += x[i].x * x[i].x + x[i].y * x[i].y;
  alpha *= alpha;
  for (int c = 0; c < C; c++) {
    float2 acc = make_float2(0.f, 0.f);
    for (int i = 0; i < M; i++) acc = cadd(acc, cmul(Wg[(size_t)(c * M + i) * ws], x[i]));
    y[c] = acc;
  }
  // (E y)_r = sum_{c != r} y_r conj(y_c) y_c = y_r * (sum_c |y_c|^2 - |y_r|^2)   (gss.cpp:124-125)
  float tot = 0.f;
  for (int c = 0; c < C; c++) tot += y[c].x * y[c].x + y[c].y * y[c].y;
  const float s1 = (float)(4 * C) / alpha;   // gss.cpp:132
  float2 wa[MAXC];
  for (int r = 0; r < C; r++) {
    const float e = s1 * (tot - (y[r].x * y[r].x + y[r].y * y[r].y));
    const float2 ey = make_float2(e * y[r].x, e * y[r].y);
    if (p.gss_dj2_scale != 0.f) {   // (W A - I) row r; only K = 0 keeps the geometric term (gss.cpp:133, integer 1/(K+1))
      for (int c = 0; c < C; c++) {
        float2 acc = make_float2(c == r ? -1.f : 0.f, 0.f);
        for (int i = 0; i < M; i++) acc = cadd(acc, cmul(Wg[(size_t)(r * M + i) * ws], steer_l[(size_t)c * M + i]));
        wa[c] = acc;
      }
    }
    for (int i = 0; i < M; i++) {
      float2 dj = cmulc(ey, x[i]);   // dJ1(r,i) = (s1 (E y)_r) conj(x_i)
      if (p.gss_dj2_scale != 0.f) {
        float2 acc = make_float2(0.f, 0.f);
        for (int c = 0; c < C; c++) acc = cadd(acc, cmulc(wa[c], steer_l[(size_t)c * M + i]));   // ((WA-I) A^H)(r,i)
        dj.x += p.gss_dj2_scale * acc.x;
        dj.y += p.gss_dj2_scale * acc.y;
      }
      float2 w = Wg[(size_t)(r * M + i) * ws];
      w.x = p.lambda_mu * w.x - p.mu * dj.x;   // gss.cpp:136
      w.y = p.lambda_mu * w.y - p.mu * dj.y;
      Wg[(size_t)(r * M + i) * ws] = w;
    }
  }
  return y[0];
}

// The same update spread over a group of kGssGroup lanes (sel_pairs_kernel): the rows of W are independent once
// sum_c |y_c|^2 is known, so lane g of the group takes rows g, g + kGssGroup, ... and the group shares the sum through two
// shuffles.  A (bin, frame) item is one long dependent chain for a single thread and the solve phase of the kernel lasts as
// long as one chain (profiles/r02_ncu_c3g.txt: half of the stall samples at the barrier behind it); the group shortens it
// by the number of rows per lane.  `mask` names the lanes of the group (they all take the same branches).
constexpr int kGssGroup = 4;
constexpr int kGssRows = (kMaxC + kGssGroup - 1) / kGssGroup;   // rows of W per lane

// W rows of this lane for one bin, held in registers across the two frames of a pair (loaded once, stored once).
struct GssRows {
  float2 w[kGssRows][8];
};
__device__ __forceinline__ void gss_rows_load(const KernelParams& p, const float2* Wg, size_t ws, int g, GssRows& R) {
#pragma unroll
  for (int k = 0; k < kGssRows; k++) {
    const int r = g + k * kGssGroup;
#pragma unroll
    for (int i = 0; i < 8; i++) R.w[k][i] = (r < p.C && i < p.M) ? Wg[(size_t)(r * p.M + i) * ws] : make_float2(0.f, 0.f);
  }
}
__device__ __forceinline__ void gss_rows_store(const KernelParams& p, float2* Wg, size_t ws, int g, const GssRows& R) {
#pragma unroll
  for (int k = 0; k < kGssRows; k++) {
    const int r = g + k * kGssGroup;
#pragma unroll
    for (int i = 0; i < 8; i++)
      if (r < p.C && i < p.M) Wg[(size_t)(r * p.M + i) * ws] = R.w[k][i];
  }
}
// gss.cpp:118-137 for one selected frame on the register-resident rows; x[i] = 0 for i >= M.  Returns y_0 on lane 0 of the group.
__device__ __forceinline__ float2 gss_rows_step(const KernelParams& p, GssRows& R, const float2 (&x)[8], const float2* steer_l, int g, unsigned mask) {
  const int C = p.C, M = p.M;
  float2 y[kGssRows];
  float alpha = 0.f;
  for (int i = 0; i < M; i++) alpha += x[i].x * x[i].x + x[i].y * x[i].y;
  alpha *= alpha;
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < kGssRows; k++) {
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; i++) acc = cadd(acc, cmul(R.w[k][i], x[i]));   // rows >= C and columns >= M are zero
    y[k] = acc;
    tot += acc.x * acc.x + acc.y * acc.y;
  }
  // (E y)_r = sum_{c != r} y_r conj(y_c) y_c = y_r * (sum_c |y_c|^2 - |y_r|^2)   (gss.cpp:124-125)
  tot += __shfl_xor_sync(mask, tot, 1);
  tot += __shfl_xor_sync(mask, tot, 2);
  const float s1 = (float)(4 * C) / alpha;   // gss.cpp:132
#pragma unroll
  for (int k = 0; k < kGssRows; k++) {
    const int r = g + k * kGssGroup;
    if (r >= C) break;
    const float e = s1 * (tot - (y[k].x * y[k].x + y[k].y * y[k].y));
    const float2 ey = make_float2(e * y[k].x, e * y[k].y);
    float2 wa[kMaxC];
    if (p.gss_dj2_scale != 0.f) {   // (W A - I) row r; only K = 0 keeps the geometric term (gss.cpp:133, integer 1/(K+1))
      for (int c = 0; c < C; c++) {
        float2 acc = make_float2(c == r ? -1.f : 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; i++)
          if (i < M) acc = cadd(acc, cmul(R.w[k][i], steer_l[(size_t)c * M + i]));
        wa[c] = acc;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (i >= M) break;
      float2 dj = cmulc(ey, x[i]);   // dJ1(r,i) = (s1 (E y)_r) conj(x_i)
      if (p.gss_dj2_scale != 0.f) {
        float2 acc = make_float2(0.f, 0.f);
        for (int c = 0; c < C; c++) acc = cadd(acc, cmulc(wa[c], steer_l[(size_t)c * M + i]));   // ((WA-I) A^H)(r,i)
        dj.x += p.gss_dj2_scale * acc.x;
        dj.y += p.gss_dj2_scale * acc.y;
      }
      R.w[k][i].x = p.lambda_mu * R.w[k][i].x - p.mu * dj.x;   // gss.cpp:136
      R.w[k][i].y = p.lambda_mu * R.w[k][i].y - p.mu * dj.y;
    }
  }
  return y[0];
}

// FP64 re-decision of the magnitude gate for one (bin, frame): exact double DFT of that bin for every
// microphone (util.h:235 windowing in double, mvdr.cpp:79-85 statistic), one warp per item.
__device__ __forceinline__ bool gate_fp64(const KernelParams& p, int s, int t, int l, int f, int lane) {
  const int j = (l == kL1K - 1) ? 513 : l;
  double stat = 0.0;
  for (int ch = 0; ch < p.M; ch++) {
    const int hf = t + f;
    const float* h0 = (hf - 1 < 0) ? p.prev_hop + ((size_t)s * p.M + ch) * p.H
                                   : p.in + (size_t)s * p.in_stream_stride + (size_t)ch * p.in_mic_stride + (size_t)(hf - 1) * p.H;
    const float* h1 = p.in + (size_t)s * p.in_stream_stride + (size_t)ch * p.in_mic_stride + (size_t)hf * p.H;
    double re = 0.0, im = 0.0;
    for (int n = lane; n < 1024; n += 32) {
      const double xv = (double)(n < 512 ? h0[n] : h1[n - 512]) * p.win_d[n];
      const double2 w = p.twid_d[(j * n) & 1023];
      re = fma(xv, w.x, re);
      im = fma(xv, w.y, im);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      re += __shfl_xor_sync(0xffffffffu, re, o);
      im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    stat += hypot(re, im);
  }
  stat /= (double)((unsigned)p.M * (unsigned)p.N);
  return stat > p.thr_mag_d;
}

template <int ALGO>
__device__ __forceinline__ void phase_b_select(const KernelParams& p, int s, int t, bool two, const float2* zall, float2* g,
                                               SelScratch& sc, int tid, int nthreads) {
  const int M = p.M, nf = two ? 2 : 1;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  const long long F0 = (long long)p.frame_index0 + t;   // global index of frame t
  if (tid == 0) { sc.n_items = 0; sc.n_recheck = 0; }
  __syncthreads();

  // ---- B1: gate ----
  for (int l = tid; l < kL1K; l += nthreads) {
    const bool inb = p.inband[l] != 0 && !(ALGO == ALGO_MVDR && l == 0);
    for (int f = 0; f < nf; f++) {
      float stat = 0.f;
      float2 x0 = make_float2(0.f, 0.f);
      for (int ch = 0; ch < M; ch++) {
        const float2 x = unpack_bin(zall + ch * kXTile, l, f);
        if (ch == 0) x0 = x;
        stat += sqrtf(fmaf(x.x, x.x, x.y * x.y));
      }
      unsigned char fl = 0;
      float2 y = make_float2(0.f, 0.f);
      if (ALGO == ALGO_MVDR && l == 0) y = x0;   // mvdr.cpp:76 (SURVEY B-15)
      if (inb) {
        float esum = 0.f;
        for (int ch = 0; ch < M; ch++) esum += sc.sqrtE[f][ch];
        const float guard = 2.0e-5f * esum + 1.0e-6f * p.thr_mag;
        if (fabsf(stat - p.thr_mag) <= guard) {
          const int q = atomicAdd(&sc.n_recheck, 1);
          sc.recheck[q] = (unsigned short)(l * 2 + f);
        } else if (stat > p.thr_mag) {
          fl = 1;
        }
        y = make_float2(0.01f * x0.x, 0.01f * x0.y);   // mvdr.cpp:96 (overwritten by B2 when selected)
      }
      sc.flag[f][l] = fl;
      sc.y[f][l] = y;
    }
  }
  __syncthreads();
  // ---- B1b: FP64 re-decision of guarded bins ----
  for (int q = warp; q < sc.n_recheck; q += nwarps) {
    const int l = sc.recheck[q] >> 1, f = sc.recheck[q] & 1;
    const bool sel = gate_fp64(p, s, t, l, f, lane);
    if (lane == 0) sc.flag[f][l] = sel ? 1 : 0;
  }
  __syncthreads();
  // ---- work list ----
  for (int l = tid; l < kL1K; l += nthreads) {
    if (ALGO == ALGO_GSS) {
      if (sc.flag[0][l] | (two ? sc.flag[1][l] : 0)) sc.items[atomicAdd(&sc.n_items, 1)] = (unsigned short)(l * 2);
    } else {
      for (int f = 0; f < nf; f++)
        if (sc.flag[f][l]) sc.items[atomicAdd(&sc.n_items, 1)] = (unsigned short)(l * 2 + f);
    }
  }
  __syncthreads();
  // ---- B2: per-item solves ----
  for (int q = tid; q < sc.n_items; q += nthreads) {
    const int l = sc.items[q] >> 1, f = sc.items[q] & 1;
    const int slot = p.sel_slot[l];
    const float2* steer_l = p.steer + (size_t)l * p.C * M;
    if (ALGO == ALGO_GSS) {
      float2* Wg = p.gss_w + (size_t)s * 8 * M * p.Lsel + slot;   // [B][8][M][Lsel]
      for (int ff = 0; ff < nf; ff++) {
        if (!sc.flag[ff][l]) continue;
        float2 x[BF_MAX_MICS_DEV];
        for (int ch = 0; ch < M; ch++) x[ch] = unpack_bin(zall + ch * kXTile, l, ff);
        sc.y[ff][l] = gss_item(p, Wg, (size_t)p.Lsel, x, steer_l);
      }
    } else {
      float2 x[kMaxMSel], xprev[kMaxMSel];
#pragma unroll
      for (int ch = 0; ch < kMaxMSel; ch++) {
        x[ch] = (ch < M) ? unpack_bin(zall + ch * kXTile, l, f) : make_float2(0.f, 0.f);
        xprev[ch] = (ch < M && f == 1) ? unpack_bin(zall + ch * kXTile, l, 0) : make_float2(0.f, 0.f);
      }
      const float2* ring = p.hist + ((size_t)s * p.Lsel + slot) * p.P * M;
      const int subst = (f == 1) ? (int)(F0 % p.P) : -1;
      sc.y[f][l] = (ALGO == ALGO_MVDR) ? mvdr_item<kMaxMSel, float>(p, ring, subst, xprev, x, steer_l)
                                       : lcmv_item<kMaxMSel, double>(p, ring, subst, xprev, x, steer_l);
    }
  }
  __syncthreads();
  // ---- B3: history update (mvdr.cpp:99-101), Hermitian assembly, diagnostics ----
  for (int l = tid; l < kL1K; l += nthreads) {
    if (ALGO != ALGO_GSS) {
      const int slot = p.sel_slot[l];
      if (slot >= 0 && !(ALGO == ALGO_MVDR && l == 0)) {
        float2* ring = p.hist + ((size_t)s * p.Lsel + slot) * p.P * M;
        for (int f = 0; f < nf; f++) {
          float2* dst = ring + (size_t)((F0 + f) % p.P) * M;
          for (int ch = 0; ch < M; ch++) dst[ch] = unpack_bin(zall + ch * kXTile, l, f);
        }
      }
    }
    if (l <= 512) {
      float2 y0 = sc.y[0][l], y1 = two ? sc.y[1][l] : make_float2(0.f, 0.f);
      if (l == 511) {   // Hermitian part of the asymmetric pair (N/2-1, N/2+1): Yh = (Y[N/2-1] + conj(Y[N/2+1])) / 2
        const float2 p0 = sc.y[0][kL1K - 1], p1 = two ? sc.y[1][kL1K - 1] : make_float2(0.f, 0.f);
        y0 = make_float2(0.5f * (y0.x + p0.x), 0.5f * (y0.y - p0.y));
        y1 = make_float2(0.5f * (y1.x + p1.x), 0.5f * (y1.y - p1.y));
      }
      if (l == 0 || l == 512) { y0.y = 0.f; y1.y = 0.f; }   // Re(): self-conjugate bins
      g[l] = make_float2(y0.x - y1.y, y0.y + y1.x);                       // Yh_t + i Yh_{t+1}
      if (l > 0 && l < 512) g[1024 - l] = make_float2(y0.x + y1.y, y1.x - y0.y);   // conj(Yh_t) + i conj(Yh_{t+1})
    }
    if (p.capture) {
      for (int f = 0; f < nf; f++) {
        unsigned char* cap = p.capture + (size_t)s * p.capture_stream_stride + (size_t)(t + f) * p.N;
        const unsigned char fl = sc.flag[f][l];
        if (l <= 512) {
          cap[l] = fl;
          if (l > 0 && l < 511) cap[1024 - l] = fl;
        } else {
          cap[513] = fl;
        }
      }
    }
  }
}

}   // namespace bf
