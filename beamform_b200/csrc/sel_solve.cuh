// Per-bin linear algebra shared by the gated-node kernels (sel_kernel.cu, sel_stream_kernel.cu): rank-1 covariance
// updates, Cholesky of R .* whiteR (mvdr.cpp:87, :239-243), and the MVDR / LCMV weight formulas (mvdr.cpp:86-94,
// lcmv.cpp:111-119) on the Cholesky factor.  Everything lives in registers; T = float (mvdr) or double (lcmv).
#pragma once
#include "bf_device.h"
#include "phase_b_select.cuh"

namespace bf {

__device__ __forceinline__ float sqrt_approx(float x) {   // MUFU.SQRT; its ~2 ulp error sits far inside the gate's guard band
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// 1/sqrt(d) for the Cholesky pivots: MUFU.RSQ + one Newton step in FP32 (the IEEE sqrt and divide sequences are
// ~40 dependent instructions per pivot on the solve's critical path); d <= 0 / NaN still end non-finite (B-10).
__device__ __forceinline__ float inv_sqrt(float d) {
  const float r = rsqrtf(d);
  return d > 0.f && d < 3.0e38f ? r * fmaf(-0.5f * d * r, r, 1.5f) : r;
}
__device__ __forceinline__ double inv_sqrt(double d) { return 1.0 / sqrt(d); }

template <int MM, typename T>
__device__ __forceinline__ void cov_rank1(HermLower<MM, T>& A, const float2 (&hf)[MM]) {
  typedef HermLower<MM, T> HL;
  cplx<T> h[MM];
#pragma unroll
  for (int i = 0; i < MM; i++) h[i] = mk<T>((T)hf[i].x, (T)hf[i].y);
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

// R .* whiteR (diagonal * 1.001, mvdr.cpp:242) = L L^H in place; invd = 1 / diag(L)
template <int MM, typename T>
__device__ __forceinline__ void chol_in_place(const KernelParams& p, HermLower<MM, T>& A, T (&invd)[MM]) {
  typedef HermLower<MM, T> HL;
  const int M = p.M;
#pragma unroll
  for (int j = 0; j < MM; j++) {
    if (j < M) {
      T d = A.dg[j] * T(1.001);   // whiteR diagonal (mvdr.cpp:242)
#pragma unroll
      for (int k = 0; k < j; k++) { const cplx<T> l = A.lo[HL::idx(j, k)]; d = fma_t<T>(-l.x, l.x, fma_t<T>(-l.y, l.y, d)); }
      const T inv = inv_sqrt(d);
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

// mvdr.cpp:86-94 with R = L L^H: z = L^{-1} d, u = L^{-1} x, y = (z^H u) / (z^H z)
template <int MM, typename T>
__device__ __forceinline__ float2 mvdr_finish(const KernelParams& p, const HermLower<MM, T>& A, const T (&invd)[MM], const float2 (&x)[MM],
                                              const float2* steer_l) {
  cplx<T> z[MM], u[MM];
#pragma unroll
  for (int i = 0; i < MM; i++) {
    const float2 d = (i < p.M) ? steer_l[i] : make_float2(0.f, 0.f);
    z[i] = mk<T>((T)d.x, (T)d.y);
    u[i] = mk<T>((T)x[i].x, (T)x[i].y);
  }
  {   // z <- L^{-1} z and u <- L^{-1} u in one sweep: the two substitutions are independent chains
    typedef HermLower<MM, T> HL;
#pragma unroll
    for (int i = 0; i < MM; i++) {
      if (i < p.M) {
        cplx<T> az = z[i], au = u[i];
#pragma unroll
        for (int k = 0; k < i; k++) {
          const cplx<T> l = A.lo[HL::idx(i, k)];
          az.x = fma_t<T>(-l.x, z[k].x, fma_t<T>(l.y, z[k].y, az.x));
          az.y = fma_t<T>(-l.x, z[k].y, fma_t<T>(-l.y, z[k].x, az.y));
          au.x = fma_t<T>(-l.x, u[k].x, fma_t<T>(l.y, u[k].y, au.x));
          au.y = fma_t<T>(-l.x, u[k].y, fma_t<T>(-l.y, u[k].x, au.y));
        }
        z[i] = mk<T>(az.x * invd[i], az.y * invd[i]);
        u[i] = mk<T>(au.x * invd[i], au.y * invd[i]);
      } else {
        z[i] = mk<T>(T(0), T(0));
        u[i] = mk<T>(T(0), T(0));
      }
    }
  }
  const cplx<T> num = cdot_conj<MM, T>(z, u);
  const T den = cdot_conj<MM, T>(z, z).x;
  return make_float2((float)(num.x / den), (float)(num.y / den));
}

// lcmv.cpp:111-119: W = R^{-1} C (C^H R^{-1} C)^{-1}, y = W(:,0)^H x.  V = L^{-1} C, u = L^{-1} x, G = V^H V, b = V^H u,
// y = g^H b with G g = e_0.
template <int MM, typename T>
__device__ __forceinline__ float2 lcmv_finish(const KernelParams& p, const HermLower<MM, T>& A, const T (&invd)[MM], const float2 (&x)[MM],
                                              const float2* steer_l) {
  const int C = p.C, M = p.M;
  cplx<T> u[MM];
#pragma unroll
  for (int i = 0; i < MM; i++) u[i] = mk<T>((T)x[i].x, (T)x[i].y);
  fwd_solve<MM, T>(p, A, invd, u);
  cplx<T> V[kMaxC][MM];
  cplx<T> G[kMaxC][kMaxC];
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
    for (int c2 = 0; c2 <= c; c2++) {
      cplx<T> w[MM];
#pragma unroll
      for (int i = 0; i < MM; i++) w[i] = V[c2][i];
      G[c][c2] = cdot_conj<MM, T>(v, w);
    }
  }
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
  cplx<T> q[kMaxC];
  for (int i = 0; i < C; i++) {
    cplx<T> acc = mk<T>(i == 0 ? T(1) : T(0), T(0));
    for (int k = 0; k < i; k++) {
      const cplx<T> l = G[i][k];
      acc.x -= l.x * q[k].x - l.y * q[k].y;
      acc.y -= l.x * q[k].y + l.y * q[k].x;
    }
    q[i] = mk<T>(acc.x * gd[i], acc.y * gd[i]);
  }
  cplx<T> g[kMaxC];
  for (int i = C - 1; i >= 0; i--) {
    cplx<T> acc = q[i];
    for (int k = i + 1; k < C; k++) {
      const cplx<T> l = G[k][i];
      acc.x -= l.x * g[k].x + l.y * g[k].y;
      acc.y -= l.x * g[k].y - l.y * g[k].x;
    }
    g[i] = mk<T>(acc.x * gd[i], acc.y * gd[i]);
  }
  cplx<T> y = mk<T>(T(0), T(0));
  for (int c = 0; c < C; c++) {
    y.x += g[c].x * b[c].x + g[c].y * b[c].y;
    y.y += g[c].x * b[c].y - g[c].y * b[c].x;
  }
  return make_float2((float)y.x, (float)y.y);
}

}   // namespace bf
