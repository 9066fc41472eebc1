// PHASE B for the phase-mask nodes: phase.cpp:70-134 and phasempf.cpp:140-302.
//
//   B1  every (bin, frame): mean magnitude, aligned phases arg(conj(w_i) X_i), mean wrapped pairwise
//       difference -> mask decision in FP32; decisions of significant bins that fall inside a guard band
//       around a threshold are re-taken in FP64 (B1b, exact double DFT of the bin) so FP32 spectra do
//       not flip mask bits the output is sensitive to
//   B2  phase: Y = mag * e^{i arg X_0};  phasempf: MCRA noise tracking + bi-channel post-filter, a
//       sequential recursion over frames with 7 state scalars per bin kept in shared memory for the
//       whole launch, then Y = max-with-floor(|soi| - Lambda) * out_amp * e^{i arg soi}
//   B3  Hermitian assembly of G (same rule as the other nodes)
#pragma once
#include "bf_device.h"
#include "phase_b_select.cuh"

namespace bf {

struct PhaseScratch {
  float2 y[2][kL1K];
  float sqrtE[2][BF_MAX_MICS_DEV];
  unsigned short recheck[2 * kL1K];
  int n_recheck;
  unsigned char flag[2][kL1K];     // bit0: magnitude gate passed (phase.cpp:99), bit1: bin kept as source of interest
  float state[7][kL1K];            // phasempf: S_prev, S_tmp, S_min, lambda_noise, Z, rev0, rev1
  float ola[64 + 1024];            // phasempf: post-OLA moving-average window (smooth_size <= 64)
};

__device__ __forceinline__ float wrap_diff(float a, float b) {   // phase.cpp:58-60
  float d = fabsf(a - b);
  return d > 3.14159265358979f ? 6.28318530717959f - d : d;
}

// FP64 re-decision for one (bin, frame), one warp per item; returns flag bits like PhaseScratch::flag.
__device__ __forceinline__ unsigned phase_decide_fp64(const KernelParams& p, int s, int t, int l, int f, int lane, bool use_gate) {
  const int j = (l == kL1K - 1) ? 513 : l;
  double phi[BF_MAX_MICS_DEV];
  double magsum = 0.0;
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
    magsum += hypot(re, im);
    const double2 w = p.steer_d[(size_t)l * p.M + ch];   // weights(i,j); aligned = conj(w) * X
    phi[ch] = atan2(im * w.x - re * w.y, re * w.x + im * w.y);
  }
  unsigned fl = 0;
  const double mag_mean = magsum / p.M;
  if (!use_gate || mag_mean / (double)p.N > p.mag_threshold_d) fl |= 1;
  double tot = 0.0;   // phase.cpp:53-68: sum over all pairs (association order of the recursion)
  int num = 0;
  for (int a = p.M - 2; a >= 0; a--) {
    double lvl = 0.0;
    for (int b = a + 1; b < p.M; b++) {
      double d = fabs(phi[a] - phi[b]);
      if (d > 3.14159265358979323846) d = 2 * 3.14159265358979323846 - d;
      lvl += d;
      num++;
    }
    tot = lvl + tot;
  }
  if (tot / (double)num < p.min_phase_rad_d) fl |= 2;
  return fl;
}

template <int ALGO>
__device__ __forceinline__ void phase_b_phase(const KernelParams& p, int s, int t, bool two, const float2* zall, float2* g,
                                              PhaseScratch& sc, int& cur_L, int& first_L, int tid, int nthreads) {
  const int M = p.M, nf = two ? 2 : 1;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  constexpr bool kGate = (ALGO == ALGO_PHASE);
  if (tid == 0) sc.n_recheck = 0;
  __syncthreads();
  const int npairs = M * (M - 1) / 2;

  // ---- B1: decisions ----
  for (int l = tid; l < kL1K; l += nthreads) {
    for (int f = 0; f < nf; f++) {
      if (l == 0) { sc.flag[f][l] = 0; continue; }
      float magsum = 0.f, guard = 0.f, esum = 0.f;
      float phi[BF_MAX_MICS_DEV];
      const float2* st = p.steer + (size_t)l * p.C * M;
      for (int ch = 0; ch < M; ch++) {
        const float2 x = unpack_bin(zall + ch * kXTile, l, f);
        const float2 w = st[ch];
        const float a = sqrtf(fmaf(x.x, x.x, x.y * x.y));
        magsum += a;
        phi[ch] = atan2f(x.y * w.x - x.x * w.y, x.x * w.x + x.y * w.y);   // arg(conj(w) x)
        const float e = sc.sqrtE[f][ch];
        esum += e;
        guard += fminf(3.2f, 4.0e-6f * e / fmaxf(a, 1e-30f) + 2.0e-6f);
      }
      float tot = 0.f;
      for (int a = M - 2; a >= 0; a--) {
        float lvl = 0.f;
        for (int b = a + 1; b < M; b++) lvl += wrap_diff(phi[a], phi[b]);
        tot = lvl + tot;
      }
      const float mean_diff = npairs > 0 ? tot / (float)npairs : __int_as_float(0x7fc00000);   // M = 1: 0/0 (phase.cpp:111)
      guard *= 2.0f / (float)M;
      unsigned fl = 0;
      bool doubt = false;
      if (kGate) {
        if (magsum > p.thr_phase_mag) fl |= 1;
        if (fabsf(magsum - p.thr_phase_mag) <= 2.0e-5f * esum + 1.0e-6f * p.thr_phase_mag) doubt = true;
      } else {
        fl |= 1;
      }
      if (mean_diff < p.min_phase_rad) fl |= 2;
      if (fabsf(mean_diff - p.min_phase_rad) <= guard && magsum > 1.0e-4f * esum) doubt = true;   // only bins that matter
      if (doubt && p.win_d != nullptr) sc.recheck[atomicAdd(&sc.n_recheck, 1)] = (unsigned short)(l * 2 + f);
      sc.flag[f][l] = (unsigned char)fl;
    }
  }
  __syncthreads();
  // ---- B1b: FP64 re-decisions ----
  for (int q = warp; q < sc.n_recheck; q += nwarps) {
    const int l = sc.recheck[q] >> 1, f = sc.recheck[q] & 1;
    const unsigned fl = phase_decide_fp64(p, s, t, l, f, lane, kGate);
    if (lane == 0) sc.flag[f][l] = (unsigned char)fl;
  }
  __syncthreads();
  // ---- B2: per-bin output (phasempf: sequential over the two frames, state in shared memory) ----
  int cl = cur_L, fst = first_L;
  for (int f = 0; f < nf; f++) {
    // MCRA window bookkeeping is global per frame (phasempf.cpp:162-176): decide the branch once
    bool reset_branch = false;
    if (ALGO == ALGO_PHASEMPF) {
      reset_branch = cl > p.mcra_L;
      if (reset_branch) { cl = 1; fst = 0; } else { cl++; }
    }
    const float inv_cl = 1.0f / (float)cl;
    for (int l = tid; l < kL1K; l += nthreads) {
      float2 y = make_float2(0.f, 0.f);
      if (l == 0) {
        if (ALGO == ALGO_PHASE) y = unpack_bin(zall, 0, f);   // phase.cpp:87; phasempf never writes bin 0 (SURVEY B-5)
        sc.y[f][l] = y;
        continue;
      }
      float magsum = 0.f;
      float2 x0 = make_float2(0.f, 0.f);
      for (int ch = 0; ch < M; ch++) {
        const float2 x = unpack_bin(zall + ch * kXTile, l, f);
        if (ch == 0) x0 = x;
        magsum += sqrtf(fmaf(x.x, x.x, x.y * x.y));
      }
      const float mag_mean = magsum / (float)M;
      const float a0 = sqrtf(fmaf(x0.x, x0.x, x0.y * x0.y));
      const float2 unit = a0 > 0.f ? make_float2(x0.x / a0, x0.y / a0) : make_float2(1.f, 0.f);   // e^{i arg X_0}
      const unsigned fl = sc.flag[f][l];
      if (ALGO == ALGO_PHASE) {
        const float mag = ((fl & 1) && (fl & 2)) ? mag_mean : mag_mean * p.mag_mult;   // phase.cpp:114-123
        y = make_float2(mag * unit.x, mag * unit.y);
      } else {
        const bool kept = (fl & 2) != 0;
        const float soi = kept ? mag_mean : mag_mean * p.min_mag;   // phasempf.cpp:234-244
        const float itf = kept ? mag_mean * p.min_mag : mag_mean;
        const float s2 = soi * soi, i2 = itf * itf;
        // --- MCRA (phasempf.cpp:140-191); the "frequency smoothing" only scales bins 1 and N-1 by 0.75 (SURVEY B-9)
        const float Sf = (l == 1) ? 0.75f * s2 : s2;
        float S_prev = sc.state[0][l], S_tmp = sc.state[1][l], S_min = sc.state[2][l], lam = sc.state[3][l];
        const float S = p.mcra_alphaS * S_prev + (1.0f - p.mcra_alphaS) * Sf;
        if (reset_branch) { S_min = fminf(S_tmp, S); S_tmp = S; }
        else { S_min = fminf(S_min, S); S_tmp = fminf(S_tmp, S); }
        if (fst || S < S_min * p.mcra_delta || lam > s2) {
          if (fst && inv_cl > p.mcra_alphaD) lam = inv_cl * lam + (1.0f - inv_cl) * s2;
          else lam = p.mcra_alphaD2 * lam + (1.0f - p.mcra_alphaD) * s2;   // SURVEY B-16
        }
        sc.state[0][l] = S; sc.state[1][l] = S_tmp; sc.state[2][l] = S_min; sc.state[3][l] = lam;
        // --- bi-channel post-filter (phasempf.cpp:255-271)
        const float Z = p.mpf_alphaS * sc.state[4][l] + (1.0f - p.mpf_alphaS) * i2;
        const float rev0 = p.mpf_gamma * sc.state[5][l] + p.mpf_rev_gain * s2;
        const float rev1 = p.mpf_gamma * sc.state[6][l] + p.mpf_rev_gain * i2;
        sc.state[4][l] = Z; sc.state[5][l] = rev0; sc.state[6][l] = rev1;
        const float Lam = sqrtf(lam + p.mpf_eta * Z + rev0 + rev1);
        float mag;
        if (p.out_only_noise) {
          mag = Lam * p.out_amp;
        } else {
          mag = p.out_only_mcra ? (soi - sqrtf(lam)) * p.out_amp : (soi - Lam) * p.out_amp;
          if (mag < 0.f) mag = p.noise_floor;
        }
        const float2 u2 = soi > 0.f ? unit : make_float2(1.f, 0.f);   // arg(out_soi[j])
        y = make_float2(mag * u2.x, mag * u2.y);
      }
      sc.y[f][l] = y;
    }
  }
  cur_L = cl;
  first_L = fst;
  __syncthreads();
  // ---- B3: Hermitian assembly + diagnostics ----
  for (int l = tid; l < kL1K; l += nthreads) {
    if (l <= 512) {
      float2 y0 = sc.y[0][l], y1 = two ? sc.y[1][l] : make_float2(0.f, 0.f);
      if (l == 511) {
        const float2 p0 = sc.y[0][kL1K - 1], p1 = two ? sc.y[1][kL1K - 1] : make_float2(0.f, 0.f);
        y0 = make_float2(0.5f * (y0.x + p0.x), 0.5f * (y0.y - p0.y));
        y1 = make_float2(0.5f * (y1.x + p1.x), 0.5f * (y1.y - p1.y));
      }
      if (l == 0 || l == 512) { y0.y = 0.f; y1.y = 0.f; }
      g[l] = make_float2(y0.x - y1.y, y0.y + y1.x);
      if (l > 0 && l < 512) g[1024 - l] = make_float2(y0.x + y1.y, y1.x - y0.y);
    }
    if (p.capture) {
      for (int f = 0; f < nf; f++) {
        unsigned char* cap = p.capture + (size_t)s * p.capture_stream_stride + (size_t)(t + f) * p.N;
        unsigned char fl = sc.flag[f][l];
        fl = (unsigned char)(((fl & 1) && (fl & 2)) ? 2 : 0);   // "kept": gate passed and phases agree
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
