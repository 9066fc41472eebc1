// Frame-size-generic fused kernel (512-, 2048- and 4096-point frames; das / phase / phasempf), sm_100a.
//
//   reference path replaced (citations /root/reference/beamform/src/): util.h:217-253,289-314 (window, framing,
//   OLA), das.cpp:47-70, phase.cpp:70-134, phasempf.cpp:140-302 + the output smoother phasempf.cpp:78-83,331-334.
//
// One CTA owns one stream and walks its frame pairs in order.  The transform is a shared-memory Stockham FFT whose
// radix-8/16 butterflies run in registers (fft_reg.cuh): each pass lets every thread gather R points of one
// transform, multiply by the pass twiddles, run an R-point register FFT and scatter the results in place (all
// reads of a round precede all writes: two block barriers per round).  4096 = 16*16*16, 2048 = 16*16*8,
// 1024 = 16*8*8, 512 = 8*8*8.  As in the 1024-point kernels the two real frames of a pair ride in the real and
// imaginary parts of one complex transform, spectra never leave shared memory and every input sample is
// fetched from HBM once (its second touch hits L1/L2).
//
// Phase-mask nodes: their forward transforms run in FP64 (B200 issues DFMA at half the FFMA rate).  Every mask
// decision (phase.cpp:114, phasempf.cpp:234) is a threshold on phases of individual bins, and with FP32 spectra a
// few percent of the bins of a 4096-point frame fall inside the rounding guard band; re-deciding those from an exact
// double DFT of the bin cost 7x the transforms themselves.  With double spectra the decision is taken in FP32
// (atan2f) and only results within 2e-6 rad of the threshold are recomputed in double from the same spectra, O(M).
//
// The 1024-point frames of the headline configurations use the register-resident warp FFT kernels
// (das_kernel.cu, sel_kernel.cu, frames_kernel.cu); this kernel covers the other frame sizes, notably the
// 4096-point PhaseMPF configuration C4.
#include <cstdio>
#include <type_traits>

#include "bf_device.h"
#include "fft_reg.cuh"
#include "fft_reg_d.cuh"
#include "warp_fft1024.cuh"
#include "phase_b_select.cuh"

// tuning switches (profiling builds override them with -D)
#ifndef BF_GEN_PACK_PAIR
#define BF_GEN_PACK_PAIR 0
#endif
#ifndef BF_GEN_U
#define BF_GEN_U 2
#endif
#ifndef BF_GEN_PREFETCH
#define BF_GEN_PREFETCH 1
#endif

#include "block_fft.cuh"

namespace bf {

template <int NN>
struct GenScratch {
  static constexpr int L = NN / 2 + 2;
  float2 y[2][L];
  unsigned char flag[2][L];   // bit0: magnitude gate passed (phase.cpp:99), bit1: bin kept as source of interest
  float tail[NN / 2];
  float ola[64 + NN];         // phasempf: post-OLA moving-average window (smooth_size <= 64)
};


// X_i[j] of frame f from the packed half-scaled double spectrum Z = FFT(0.5*w*(x_t + i x_{t+1}))
template <int NN>
__device__ __forceinline__ double2 unpack_nd(const double2* z, int l, int f) {
  constexpr int L = NN / 2 + 2;
  const int j = (l == L - 1) ? NN / 2 - 1 : l;   // pseudo-bin: conj of bin N/2-1 (SURVEY B-4)
  const double2 a = z[swz(j)], b = z[swz((NN - j) & (NN - 1))];
  double2 x;
  if (f == 0) x = make_double2(a.x + b.x, a.y - b.y);
  else x = make_double2(a.y + b.y, b.x - a.x);
  if (l == L - 1) x.y = -x.y;
  return x;
}

__device__ __forceinline__ float wrap_diff_n(float a, float b) {   // phase.cpp:58-60
  const float d = fabsf(a - b);
  return d > 3.14159265358979f ? 6.28318530717959f - d : d;
}

// Decision of one (bin, frame) in double from the double spectra (phase.cpp:89-123, phasempf.cpp:212-248)
template <int NN>
__device__ __noinline__ unsigned phase_decide_d(const KernelParams& p, const double2* zall, int l, int f, bool use_gate) {
  double phi[BF_MAX_MICS_DEV];
  double magsum = 0.0;
  for (int ch = 0; ch < p.M; ch++) {
    const double2 x = unpack_nd<NN>(zall + (size_t)ch * NN, l, f);
    magsum += hypot(x.x, x.y);
    const double2 w = p.steer_d[(size_t)l * p.M + ch];
    phi[ch] = atan2(x.y * w.x - x.x * w.y, x.x * w.x + x.y * w.y);
  }
  unsigned fl = 0;
  if (!use_gate || (magsum / p.M) / (double)NN > p.mag_threshold_d) fl |= 1;
  double tot = 0.0;   // phase.cpp:53-68: association order of the recursion
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

// phase.cpp:70-134 / phasempf.cpp:193-302 for the two frames of a pair -> sc.y
template <int ALGO, int NN>
__device__ __forceinline__ void phase_pair_n(const KernelParams& p, int s, int t, bool two, const double2* zall, GenScratch<NN>& sc,
                                             int& cur_L, int& first_L, int tid) {
  constexpr int L = NN / 2 + 2;
  const int M = p.M, nf = two ? 2 : 1;
  constexpr bool kGate = (ALGO == ALGO_PHASE);
  const int npairs = M * (M - 1) / 2;
  // ---- decisions: FP32 trigonometry on the double spectra; results within rounding of a threshold redone in double ----
  for (int l = tid; l < L; l += kGenThreads) {
    for (int f = 0; f < nf; f++) {
      if (l == 0) { sc.flag[f][l] = 0; continue; }
      float magsum = 0.f;
      float phi[BF_MAX_MICS_DEV];
      const float2* st = p.steer + (size_t)l * p.C * M;
      for (int ch = 0; ch < M; ch++) {
        const double2 xd = unpack_nd<NN>(zall + (size_t)ch * NN, l, f);
        const float2 x = make_float2((float)xd.x, (float)xd.y);
        const float2 w = st[ch];
        magsum += sqrtf(fmaf(x.x, x.x, x.y * x.y));
        phi[ch] = atan2f(x.y * w.x - x.x * w.y, x.x * w.x + x.y * w.y);   // arg(conj(w) x)
      }
      float tot = 0.f;
      for (int a = M - 2; a >= 0; a--) {
        float lvl = 0.f;
        for (int b = a + 1; b < M; b++) lvl += wrap_diff_n(phi[a], phi[b]);
        tot = lvl + tot;
      }
      const float mean_diff = npairs > 0 ? tot / (float)npairs : __int_as_float(0x7fc00000);   // M = 1: 0/0 (phase.cpp:111)
      unsigned fl = 0;
      bool doubt = fabsf(mean_diff - p.min_phase_rad) <= 4.0e-6f;   // float steering table + atan2f: ~1e-6 rad
      const float thr = p.thr_phase_mag;
      if (kGate) {
        if (magsum > thr) fl |= 1;
        if (fabsf(magsum - thr) <= 2.0e-6f * thr) doubt = true;
      } else {
        fl |= 1;
      }
      if (mean_diff < p.min_phase_rad) fl |= 2;
      if (doubt && M > 1) fl = phase_decide_d<NN>(p, zall, l, f, kGate);
      sc.flag[f][l] = (unsigned char)fl;
    }
  }
  __syncthreads();
  (void)s; (void)t;
  // ---- per-bin output; phasempf: MCRA + bi-channel post-filter, state in global memory ([7][L], bin fastest) ----
  int cl = cur_L, fst = first_L;
  float* stg = (ALGO == ALGO_PHASEMPF) ? p.mpf_state + (size_t)s * 7 * L : nullptr;
  for (int f = 0; f < nf; f++) {
    bool reset_branch = false;
    if (ALGO == ALGO_PHASEMPF) {   // phasempf.cpp:162-176: window bookkeeping is global per frame
      reset_branch = cl > p.mcra_L;
      if (reset_branch) { cl = 1; fst = 0; } else { cl++; }
    }
    const float inv_cl = 1.0f / (float)cl;
    for (int l = tid; l < L; l += kGenThreads) {
      float2 y = make_float2(0.f, 0.f);
      if (l == 0) {
        if (ALGO == ALGO_PHASE) { const double2 xd = unpack_nd<NN>(zall, 0, f); y = make_float2((float)xd.x, (float)xd.y); }   // phase.cpp:87; phasempf leaves bin 0 at 0 (B-5)
        sc.y[f][l] = y;
        continue;
      }
      float magsum = 0.f;
      float2 x0 = make_float2(0.f, 0.f);
      for (int ch = 0; ch < M; ch++) {
        const double2 xd = unpack_nd<NN>(zall + (size_t)ch * NN, l, f);
        const float2 x = make_float2((float)xd.x, (float)xd.y);
        if (ch == 0) x0 = x;
        magsum += sqrtf(fmaf(x.x, x.x, x.y * x.y));
      }
      const float mag_mean = magsum / (float)M;
      const float n0 = fmaf(x0.x, x0.x, x0.y * x0.y);
      const float r0 = rsqrtf(n0);
      const float2 unit = n0 > 0.f ? make_float2(x0.x * r0, x0.y * r0) : make_float2(1.f, 0.f);   // e^{i arg X_0}
      const unsigned fl = sc.flag[f][l];
      if (ALGO == ALGO_PHASE) {
        const float mag = ((fl & 1) && (fl & 2)) ? mag_mean : mag_mean * p.mag_mult;   // phase.cpp:114-123
        y = make_float2(mag * unit.x, mag * unit.y);
      } else {
        const bool kept = (fl & 2) != 0;
        const float soi = kept ? mag_mean : mag_mean * p.min_mag;   // phasempf.cpp:234-244
        const float itf = kept ? mag_mean * p.min_mag : mag_mean;
        const float s2 = soi * soi, i2 = itf * itf;
        const float Sf = (l == 1) ? 0.75f * s2 : s2;   // SURVEY B-9: only bins 1 and N-1 are scaled
        float S_prev = stg[0 * L + l], S_tmp = stg[1 * L + l], S_min = stg[2 * L + l], lam = stg[3 * L + l];
        const float S = p.mcra_alphaS * S_prev + (1.0f - p.mcra_alphaS) * Sf;
        if (reset_branch) { S_min = fminf(S_tmp, S); S_tmp = S; }
        else { S_min = fminf(S_min, S); S_tmp = fminf(S_tmp, S); }
        if (fst || S < S_min * p.mcra_delta || lam > s2) {
          if (fst && inv_cl > p.mcra_alphaD) lam = inv_cl * lam + (1.0f - inv_cl) * s2;
          else lam = p.mcra_alphaD2 * lam + (1.0f - p.mcra_alphaD) * s2;   // SURVEY B-16
        }
        stg[0 * L + l] = S; stg[1 * L + l] = S_tmp; stg[2 * L + l] = S_min; stg[3 * L + l] = lam;
        const float Z = p.mpf_alphaS * stg[4 * L + l] + (1.0f - p.mpf_alphaS) * i2;   // phasempf.cpp:255-271
        const float rev0 = p.mpf_gamma * stg[5 * L + l] + p.mpf_rev_gain * s2;
        const float rev1 = p.mpf_gamma * stg[6 * L + l] + p.mpf_rev_gain * i2;
        stg[4 * L + l] = Z; stg[5 * L + l] = rev0; stg[6 * L + l] = rev1;
        const float Lam = sqrtf(lam + p.mpf_eta * Z + rev0 + rev1);
        float mag;
        if (p.out_only_noise) {
          mag = Lam * p.out_amp;
        } else {
          mag = p.out_only_mcra ? (soi - sqrtf(lam)) * p.out_amp : (soi - Lam) * p.out_amp;
          if (mag < 0.f) mag = p.noise_floor;
        }
        const float2 u2 = soi > 0.f ? unit : make_float2(1.f, 0.f);
        y = make_float2(mag * u2.x, mag * u2.y);
      }
      sc.y[f][l] = y;
    }
    // the same thread owns bin l in both frames (same loop mapping), so the state needs no barrier between frames
  }
  cur_L = cl;
  first_L = fst;
  __syncthreads();
}

// Fused per-bin stage for small arrays (2 <= M <= 4, spectra of a bin held in registers): one pass over the bins
// loads each microphone's packed double spectrum once, takes the mask decisions of BOTH frames, runs the MCRA /
// post-filter recursion over the two frames with one state load and one store, and writes the Hermitian-assembled
// G = Yh_t + i Yh_{t+1} straight into the inverse transform's buffer (phase.cpp:70-134, phasempf.cpp:193-302).
// Two microphones (C4): |phi_0 - phi_1| wrapped equals |arg(z_0 conj z_1)|, z_i = conj(w_i) X_i, so the test
// "< min_phase" is the sign of sin(thr)*Re(u) - cos(thr)*|Im(u)| (0 < thr < pi) and needs no arctangent; results within
// FP32 rounding of the threshold are re-decided in double from the same double spectra (phase_decide_d).
template <int ALGO, int NN, int MM>
__device__ __forceinline__ void phase_pair_fused(const KernelParams& p, int s, int t, bool two, unsigned g_off, int& cur_L, int& first_L,
                                                 int tid, float sin_thr, float cos_thr) {
  const double2* zall = reinterpret_cast<const double2*>(gen_smem_raw);
  float2* gbuf = reinterpret_cast<float2*>(gen_smem_raw + g_off);
  constexpr int H = NN / 2, L = NN / 2 + 2;
  constexpr bool kGate = (ALGO == ALGO_PHASE);
  constexpr bool kMpf = (ALGO == ALGO_PHASEMPF);
  const int M = p.M, nf = two ? 2 : 1;
  const int npairs = M * (M - 1) / 2;
  const bool fast2 = (MM == 2) && p.min_phase_rad > 0.f && p.min_phase_rad < 3.1415925f;
  // MCRA window bookkeeping (phasempf.cpp:162-176) is global per frame: resolve both frames up front
  bool reset_f[2] = {false, false};
  float inv_cl_f[2] = {1.f, 1.f};
  int fst_f[2] = {first_L, first_L};
  {
    int cl = cur_L, fst = first_L;
    if (kMpf) {
      for (int f = 0; f < nf; f++) {
        const bool r = cl > p.mcra_L;
        if (r) { cl = 1; fst = 0; } else { cl++; }
        reset_f[f] = r; inv_cl_f[f] = 1.0f / (float)cl; fst_f[f] = fst;
      }
    }
    cur_L = cl;
    first_L = fst;
  }
  float* stg = kMpf ? p.mpf_state + (size_t)s * 7 * L : nullptr;

  // A bin is evaluated in three straight-line stages so that the U bins a thread carries per trip interleave in the
  // instruction stream (8 warps per SM: the latency of a bin's load -> decision -> recursion chain has to be hidden by
  // the thread's own independent bins):  A loads + mask decisions,  B rare FP64 re-decision,  C recursion + output.
  struct Bin {
    float2 x[2][MM];
    float st[7];       // phasempf state of the bin: S_prev, S_tmp, S_min, lambda_noise, Z, rev0, rev1
    float magsum[2];
    unsigned fl[2];
    bool doubt[2];
  };
  auto stage_a = [&](int l, Bin& B) {
    const int j = (l == L - 1) ? H - 1 : l;   // pseudo-bin: conj of bin N/2-1 (SURVEY B-4)
    if (kMpf) {
#pragma unroll
      for (int k = 0; k < 7; k++) B.st[k] = stg[k * L + l];
    }
    float2 wst[MM];
    {
      const float2* st = p.steer + (size_t)l * p.C * M;
#pragma unroll
      for (int i = 0; i < MM; i++) wst[i] = (i < M) ? __ldg(st + i) : make_float2(1.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < MM; i++) {
      if (i < M) {
        const double2 a = zall[(size_t)i * NN + swz(j)], b = zall[(size_t)i * NN + swz((NN - j) & (NN - 1))];
        B.x[0][i] = make_float2((float)(a.x + b.x), (float)(a.y - b.y));   // Z[j] + conj(Z[N-j])
        B.x[1][i] = make_float2((float)(a.y + b.y), (float)(b.x - a.x));   // -i (Z[j] - conj(Z[N-j]))
        if (l == L - 1) { B.x[0][i].y = -B.x[0][i].y; B.x[1][i].y = -B.x[1][i].y; }
      } else {
        B.x[0][i] = B.x[1][i] = make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int f = 0; f < 2; f++) {
      float magsum = 0.f;
      float2 zr[MM];
#pragma unroll
      for (int i = 0; i < MM; i++) {
        const float2 xi = B.x[f][i];
        if (i < M) magsum += sqrt_fast(fmaf(xi.x, xi.x, xi.y * xi.y));
        zr[i] = make_float2(xi.x * wst[i].x + xi.y * wst[i].y, xi.y * wst[i].x - xi.x * wst[i].y);   // conj(w) x
      }
      unsigned fl = 0;
      bool doubt = false;
      if (fast2) {
        const float d = zr[0].x * zr[1].x + zr[0].y * zr[1].y;    // Re(z0 conj z1)
        const float c = zr[0].y * zr[1].x - zr[0].x * zr[1].y;    // Im(z0 conj z1)
        const float a = sin_thr * d, b = cos_thr * fabsf(c);
        const float q = a - b;                                      // |z0||z1| sin(thr - |dphi|)
        if (q > 0.f) fl |= 2;
        doubt = fabsf(q) <= 4.0e-6f * (fabsf(a) + fabsf(b));
      } else {
        float phi[MM];
#pragma unroll
        for (int i = 0; i < MM; i++) phi[i] = atan2f(zr[i].y, zr[i].x);
        float tot = 0.f;
#pragma unroll
        for (int a = MM - 2; a >= 0; a--) {
          float lvl = 0.f;
#pragma unroll
          for (int b = a + 1; b < MM; b++)
            if (b < M) lvl += wrap_diff_n(phi[a], phi[b]);
          if (a <= M - 2) tot = lvl + tot;
        }
        const float mean_diff = tot / (float)npairs;
        if (mean_diff < p.min_phase_rad) fl |= 2;
        doubt = fabsf(mean_diff - p.min_phase_rad) <= 4.0e-6f;
      }
      if (kGate) {
        const float thr = p.thr_phase_mag;
        if (magsum > thr) fl |= 1;
        if (fabsf(magsum - thr) <= 2.0e-6f * thr) doubt = true;
      } else {
        fl |= 1;
      }
      B.magsum[f] = magsum; B.fl[f] = fl; B.doubt[f] = doubt && f < nf;
    }
  };
  auto stage_b = [&](int l, Bin& B) {
#pragma unroll
    for (int f = 0; f < 2; f++)
      if (B.doubt[f]) B.fl[f] = phase_decide_d<NN>(p, zall, l, f, kGate);
  };
  const float inv_M = 1.0f / (float)M;
  auto stage_c = [&](int l, Bin& B, float2& y0, float2& y1) {
    float2 yy[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    float S_prev = B.st[0], S_tmp = B.st[1], S_min = B.st[2], lam = B.st[3], Zs = B.st[4], rev0 = B.st[5], rev1 = B.st[6];
#pragma unroll
    for (int f = 0; f < 2; f++) {
      if (f >= nf) break;
      const unsigned fl = B.fl[f];
      if (p.capture) {
        unsigned char* cap = p.capture + (size_t)s * p.capture_stream_stride + (size_t)(t + f) * NN;
        const unsigned char cf = (unsigned char)(((fl & 1) && (fl & 2)) ? 2 : 0);
        if (l <= H) {
          cap[l] = cf;
          if (l > 0 && l < H - 1) cap[NN - l] = cf;
        } else {
          cap[H + 1] = cf;
        }
      }
      const float2 x0 = B.x[f][0];
      const float mag_mean = B.magsum[f] * inv_M;
      const float n0 = fmaf(x0.x, x0.x, x0.y * x0.y);
      const float r0 = rsqrtf(n0);
      const float2 unit = n0 > 0.f ? make_float2(x0.x * r0, x0.y * r0) : make_float2(1.f, 0.f);   // e^{i arg X_0}
      float2 y;
      if (ALGO == ALGO_PHASE) {
        const float mag = ((fl & 1) && (fl & 2)) ? mag_mean : mag_mean * p.mag_mult;   // phase.cpp:114-123
        y = make_float2(mag * unit.x, mag * unit.y);
      } else {
        const bool kept = (fl & 2) != 0;
        const float soi = kept ? mag_mean : mag_mean * p.min_mag;   // phasempf.cpp:234-244
        const float itf = kept ? mag_mean * p.min_mag : mag_mean;
        const float s2 = soi * soi, i2 = itf * itf;
        const float Sf = (l == 1) ? 0.75f * s2 : s2;   // SURVEY B-9: only bins 1 and N-1 are scaled
        const float S = p.mcra_alphaS * S_prev + (1.0f - p.mcra_alphaS) * Sf;
        const float m_min = reset_f[f] ? S_tmp : S_min;
        S_min = fminf(m_min, S);
        S_tmp = reset_f[f] ? S : fminf(S_tmp, S);
        const bool upd = fst_f[f] || S < S_min * p.mcra_delta || lam > s2;
        const bool avg = fst_f[f] && inv_cl_f[f] > p.mcra_alphaD;
        const float ca = avg ? inv_cl_f[f] : p.mcra_alphaD2, cb = avg ? 1.0f - inv_cl_f[f] : 1.0f - p.mcra_alphaD;   // SURVEY B-16
        lam = upd ? ca * lam + cb * s2 : lam;
        S_prev = S;
        Zs = p.mpf_alphaS * Zs + (1.0f - p.mpf_alphaS) * i2;   // phasempf.cpp:255-271
        rev0 = p.mpf_gamma * rev0 + p.mpf_rev_gain * s2;
        rev1 = p.mpf_gamma * rev1 + p.mpf_rev_gain * i2;
        const float Lam = sqrt_fast(lam + p.mpf_eta * Zs + rev0 + rev1);
        float mag;
        if (p.out_only_noise) {
          mag = Lam * p.out_amp;
        } else {
          mag = p.out_only_mcra ? (soi - sqrt_fast(lam)) * p.out_amp : (soi - Lam) * p.out_amp;
          if (mag < 0.f) mag = p.noise_floor;
        }
        const float2 u2 = soi > 0.f ? unit : make_float2(1.f, 0.f);
        y = make_float2(mag * u2.x, mag * u2.y);
      }
      yy[f] = y;
    }
    if (kMpf) {
      stg[0 * L + l] = S_prev; stg[1 * L + l] = S_tmp; stg[2 * L + l] = S_min; stg[3 * L + l] = lam;
      stg[4 * L + l] = Zs; stg[5 * L + l] = rev0; stg[6 * L + l] = rev1;
    }
    y0 = yy[0];
    y1 = yy[1];
  };

  // Thread tid takes bins tid, tid + T, ..., U per trip.  Bin 0 has no decision (phase.cpp:87, SURVEY B-5), so its
  // thread takes the Nyquist bin instead; the pseudo-bin N/2+1 rides in an extra slot of the thread that owns bin
  // N/2-1 (last trip) and is folded into it: Yh = (Y[N/2-1] + conj(Y[N/2+1])) / 2.
  constexpr int U = BF_GEN_U;
  constexpr int kTrips = (H + U * kGenThreads - 1) / (U * kGenThreads);
  if (tid == 0) {
    float2 g0 = make_float2(0.f, 0.f);
    if (ALGO == ALGO_PHASE) {   // Y[0] = X_0[0] (real for real input); phasempf leaves bin 0 at 0
      const double2 a = zall[swz(0)];
      g0 = make_float2((float)(2.0 * a.x), two ? (float)(2.0 * a.y) : 0.f);
    }
    gbuf[swz(0)] = g0;
    if (p.capture)
      for (int f = 0; f < nf; f++) p.capture[(size_t)s * p.capture_stream_stride + (size_t)(t + f) * NN] = 0;
  }
#pragma unroll 1
  for (int trip = 0; trip < kTrips; trip++) {
    int lu[U + 1];
    bool on[U + 1];
    Bin B[U + 1];
    float2 y0[U + 1], y1[U + 1];
    on[U] = false; lu[U] = L - 1;
#pragma unroll
    for (int u = 0; u < U; u++) {
      lu[u] = tid + (trip * U + u) * kGenThreads;
      on[u] = lu[u] < H;
      if (!on[u]) lu[u] = 1;
      if (lu[u] == 0) lu[u] = H;
      if (lu[u] == H - 1) on[U] = true;
    }
#pragma unroll
    for (int u = 0; u < U; u++) stage_a(lu[u], B[u]);
    if (on[U]) stage_a(lu[U], B[U]);
#pragma unroll
    for (int u = 0; u <= U; u++)
      if (on[u]) stage_b(lu[u], B[u]);
#pragma unroll
    for (int u = 0; u <= U; u++)
      if (on[u]) stage_c(lu[u], B[u], y0[u], y1[u]);
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (!on[u]) continue;
      const int l = lu[u];
      float2 a0 = y0[u], a1 = y1[u];
      if (l == H - 1) {
        a0 = make_float2(0.5f * (a0.x + y0[U].x), 0.5f * (a0.y - y0[U].y));
        a1 = make_float2(0.5f * (a1.x + y1[U].x), 0.5f * (a1.y - y1[U].y));
      }
      if (l == H) { a0.y = 0.f; a1.y = 0.f; }   // Re(): self-conjugate bin
      gbuf[swz(l)] = make_float2(a0.x - a1.y, a0.y + a1.x);                          // Yh_t + i Yh_{t+1}
      if (l < H) gbuf[swz(NN - l)] = make_float2(a0.x + a1.y, a1.x - a0.y);          // conj(Yh_t) + i conj(Yh_{t+1})
    }
  }
}


template <int ALGO, int NN>
__global__ void __launch_bounds__(kGenThreads, 1) frames_kernel_n(const __grid_constant__ KernelParams p) {
  constexpr int H = NN / 2, L = NN / 2 + 2;
  constexpr bool kPha = (ALGO == ALGO_PHASE || ALGO == ALGO_PHASEMPF);
  constexpr bool kSmooth = (ALGO == ALGO_PHASEMPF);
  typedef typename std::conditional<kPha, double2, float2>::type ZV;   // forward spectra: double for the phase masks
  unsigned char* smem_raw = gen_smem_raw;
  ZV* zall = reinterpret_cast<ZV*>(smem_raw);                            // [M][NN]
  const int m_tiles = (ALGO == ALGO_DAS) ? p.das_chunk : p.M;             // spectrum tiles resident at once
  float2* gbuf = reinterpret_cast<float2*>(zall + (size_t)m_tiles * NN);   // [NN]
  const unsigned g_off = (unsigned)((size_t)m_tiles * NN * sizeof(ZV));
  GenScratch<NN>& sc = *reinterpret_cast<GenScratch<NN>*>(gbuf + NN);
  const int tid = threadIdx.x;
  const int M = p.M;
  const int s = blockIdx.x + p.stream_begin;
  const float2* tw = p.twid_f;
  const float* win = p.win_f;
  int cur_L = p.mcra_cur_L0, first_L = p.mcra_first0;
  double win_s = 0.0, win_c = 0.0;   // 0.5 * (sin, cos)(pi * tid / N)
  if (kPha) {
    sincospi((double)tid / (double)NN, &win_s, &win_c);
    win_s *= 0.5; win_c *= 0.5;
  }
  float sin_thr = 0.f, cos_thr = 1.f;
  if (kPha) {
    double sd, cd;
    sincos(p.min_phase_rad_d, &sd, &cd);
    sin_thr = (float)sd; cos_thr = (float)cd;
  }

  for (int i = tid; i < H; i += kGenThreads) sc.tail[i] = p.tail[(size_t)s * H + i];
  if (kSmooth)
    for (int i = tid; i < p.smooth_size - 1; i += kGenThreads) sc.ola[i] = p.smooth_hist[(size_t)s * 64 + i];
  __syncthreads();

  const int nh = p.hop_end - p.hop_begin;
  const int npairs = (nh + 1) >> 1;
#ifdef BF_PHASE_TIMERS   // nvcc -DBF_PHASE_TIMERS + env BF_DEBUG=1: per-phase cycle totals of CTA 0 (profiling builds only)
  long long ph_clk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long ph_t = clock64();
#define BF_PHASE(i) do { if (p.debug == 1) { __syncthreads(); const long long n_ = clock64(); ph_clk[i] += n_ - ph_t; ph_t = n_; } } while (0)
#else
#define BF_PHASE(i) do { } while (0)
#endif
  for (int ip = 0; ip < npairs; ip++) {
    const int t = p.hop_begin + 2 * ip;
    const bool two = t + 1 < p.hop_end;
    // ---- window + pack: z = 0.5*w*(frame_t + i*frame_{t+1}), frame_t = [hop t-1 | hop t] (util.h:217-242) ----
    // the next pair's two new hops start their trip from HBM to L2 now, one 128-byte line per request
    if (BF_GEN_PREFETCH && t + 2 < p.hop_end) {
      const int lines_per_mic = ((t + 3 < p.hop_end) ? 2 : 1) * (H / 32);
      for (int i = tid; i < M * lines_per_mic; i += kGenThreads) {
        const int ch = i / lines_per_mic, ln = i - ch * lines_per_mic;
        const float* a = p.in + (size_t)s * p.in_stream_stride + (size_t)ch * p.in_mic_stride + (size_t)(t + 2) * H + ln * 32;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
      }
    }
    constexpr int kIter = (H + kGenThreads - 1) / kGenThreads;
    auto pack_load = [&](int ch, float (&fa)[kIter], float (&fb)[kIter], float (&fc)[kIter]) {
      const float* base = p.in + (size_t)s * p.in_stream_stride + (size_t)ch * p.in_mic_stride;
      const float* ha = (t - 1 < 0) ? p.prev_hop + ((size_t)s * M + ch) * H : base + (size_t)(t - 1) * H;
      const float* hb = base + (size_t)t * H;
      const float* hc = two ? base + (size_t)(t + 1) * H : hb;
#pragma unroll
      for (int k = 0; k < kIter; k++) {
        const int n = tid + k * kGenThreads;
        const bool in = (H % kGenThreads == 0) || n < H;
        fa[k] = in ? __ldg(ha + n) : 0.f;
        fb[k] = in ? __ldg(hb + n) : 0.f;
        fc[k] = (in && two) ? __ldg(hc + n) : 0.f;
      }
    };
    auto pack_store = [&](int ch, const float (&fa)[kIter], const float (&fb)[kIter], const float (&fc)[kIter]) {
      static_for<0, kIter>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        const int n = tid + k * kGenThreads;
        if ((H % kGenThreads != 0) && n >= H) return;
        const float b1 = two ? fb[k] : 0.f;
        if constexpr (kPha) {
          // 0.5 * sqrt-Hann (util.h:201-211) = 0.5 sin(pi n / N) at n = tid + 256 k and n + N/2 (cosine), by angle addition
          // from the thread's (sin, cos)(pi tid / N) and compile-time (cos, sin)(pi 256 k / N): no table traffic
          constexpr double ck = cos64((8192 / NN) * k), sk = sin64((8192 / NN) * k);
          const double w0 = fma(win_s, ck, win_c * sk), w1 = fma(win_c, ck, -win_s * sk);
          zall[(size_t)ch * NN + swz(n)] = make_double2((double)fa[k] * w0, (double)b1 * w0);
          zall[(size_t)ch * NN + swz(n + H)] = make_double2((double)fb[k] * w1, (double)fc[k] * w1);
        } else {
          const float w0 = 0.5f * __ldg(win + n), w1 = 0.5f * __ldg(win + n + H);
          zall[(size_t)ch * NN + swz(n)] = make_float2(fa[k] * w0, b1 * w0);
          zall[(size_t)ch * NN + swz(n + H)] = make_float2(fb[k] * w1, fc[k] * w1);
        }
      });
    };
    // das is linear in the spectra (das.cpp:60-63), so large arrays go through the shared-memory tile in chunks of Mc
    // microphones and accumulate into G; every other node needs all M spectra of a bin at once (Mc == M)
    const int Mc = (ALGO == ALGO_DAS) ? p.das_chunk : M;
    for (int c0 = 0; c0 < M; c0 += Mc) {
    const int Mn = min(Mc, M - c0);
    if (c0 > 0) __syncthreads();   // the previous chunk's accumulation has read its spectra
    for (int cc = 0; cc < Mn; cc += 2) {   // two microphones' loads in flight per round trip
      float fa0[kIter], fb0[kIter], fc0[kIter], fa1[kIter], fb1[kIter], fc1[kIter];
      pack_load(c0 + cc, fa0, fb0, fc0);
      if (BF_GEN_PACK_PAIR && cc + 1 < Mn) pack_load(c0 + cc + 1, fa1, fb1, fc1);
      pack_store(cc, fa0, fb0, fc0);
      if (!BF_GEN_PACK_PAIR && cc + 1 < Mn) pack_load(c0 + cc + 1, fa1, fb1, fc1);
      if (cc + 1 < Mn) pack_store(cc + 1, fa1, fb1, fc1);
    }
    __syncthreads();
    BF_PHASE(0);
    if constexpr (kPha) block_fft_fn<NN, -1, double2>(0u, Mn, p.twid_d, tid);
    else block_fft_fn<NN, -1, float2>(0u, Mn, tw, tid);
    BF_PHASE(1);
    // ---- per-bin beamformer ----
    if constexpr (ALGO == ALGO_DAS) {
      // das.cpp:60-63 commutes with the frame packing: G[j] = sum_i ceff_i[j] * Z_i[j] over all N bins
      for (int j = tid; j < NN; j += kGenThreads) {
        float2 acc = make_float2(0.f, 0.f);
        for (int ch = 0; ch < Mn; ch++) {
          const float2 z = zall[(size_t)ch * NN + swz(j)];
          const float2 w = __ldg(p.das_ceff + (size_t)(c0 + ch) * NN + j);
          acc.x = fmaf(z.x, w.x, acc.x); acc.x = fmaf(-z.y, w.y, acc.x);
          acc.y = fmaf(z.x, w.y, acc.y); acc.y = fmaf(z.y, w.x, acc.y);
        }
        const float2 prev = c0 > 0 ? gbuf[swz(j)] : make_float2(0.f, 0.f);
        gbuf[swz(j)] = make_float2(prev.x + 2.0f * acc.x, prev.y + 2.0f * acc.y);
      }
    } else {
      if (M >= 2 && M <= 4) {
        if (M == 2) phase_pair_fused<ALGO, NN, 2>(p, s, t, two, g_off, cur_L, first_L, tid, sin_thr, cos_thr);
        else if (M == 3) phase_pair_fused<ALGO, NN, 3>(p, s, t, two, g_off, cur_L, first_L, tid, sin_thr, cos_thr);
        else phase_pair_fused<ALGO, NN, 4>(p, s, t, two, g_off, cur_L, first_L, tid, sin_thr, cos_thr);
      } else {
            phase_pair_n<ALGO, NN>(p, s, t, two, zall, sc, cur_L, first_L, tid);
      for (int l = tid; l <= H; l += kGenThreads) {   // Hermitian assembly of G = Yh_t + i Yh_{t+1}
        float2 y0 = sc.y[0][l], y1 = two ? sc.y[1][l] : make_float2(0.f, 0.f);
        if (l == H - 1) {
          const float2 p0 = sc.y[0][L - 1], p1 = two ? sc.y[1][L - 1] : make_float2(0.f, 0.f);
          y0 = make_float2(0.5f * (y0.x + p0.x), 0.5f * (y0.y - p0.y));
          y1 = make_float2(0.5f * (y1.x + p1.x), 0.5f * (y1.y - p1.y));
        }
        if (l == 0 || l == H) { y0.y = 0.f; y1.y = 0.f; }
        gbuf[swz(l)] = make_float2(y0.x - y1.y, y0.y + y1.x);
        if (l > 0 && l < H) gbuf[swz(NN - l)] = make_float2(y0.x + y1.y, y1.x - y0.y);
      }
      if (p.capture) {
        for (int l = tid; l < L; l += kGenThreads)
          for (int f = 0; f < (two ? 2 : 1); f++) {
            unsigned char* cap = p.capture + (size_t)s * p.capture_stream_stride + (size_t)(t + f) * NN;
            unsigned char fl = sc.flag[f][l];
            fl = (unsigned char)(((fl & 1) && (fl & 2)) ? 2 : 0);
            if (l <= H) {
              cap[l] = fl;
              if (l > 0 && l < H - 1) cap[NN - l] = fl;
            } else {
              cap[H + 1] = fl;
            }
          }
      }
      }
    }
    }   // microphone chunks
    __syncthreads();
    BF_PHASE(2);
    // the spectra are consumed: their storage is the second buffer of an out-of-place inverse (one barrier per pass)
    const float2* res = reinterpret_cast<const float2*>(smem_raw + block_fft_oop_fn<NN, 1, float2>(g_off, 0u, tw, tid));
    BF_PHASE(3);
    // ---- synthesis window, overlap-add (util.h:244-253, 301-302), optional smoother ----
    float* o0 = p.out + (size_t)s * p.out_stream_stride + (size_t)t * H;
    const int S1 = kSmooth ? p.smooth_size - 1 : 0;
    for (int n = tid; n < H; n += kGenThreads) {
      const float w0 = __ldg(win + n) * p.out_scale, w1 = __ldg(win + n + H) * p.out_scale;
      const float2 a = res[swz(n)], b = res[swz(n + H)];
      const float r0 = sc.tail[n] + a.x * w0;
      if (kSmooth) sc.ola[S1 + n] = r0; else o0[n] = r0;
      if (two) {
        const float r1 = b.x * w1 + a.y * w0;
        if (kSmooth) sc.ola[S1 + H + n] = r1; else o0[H + n] = r1;
        sc.tail[n] = b.y * w1;
      } else {
        sc.tail[n] = b.x * w1;
      }
    }
    BF_PHASE(4);
    if (kSmooth) {
      // phasempf.cpp:78-83,122-130,331-334: every output sample becomes the mean of the last smooth_size OLA samples
      __syncthreads();
      const int cnt = two ? NN : H, S = p.smooth_size;
      const double inv = 1.0 / (double)S;
      for (int n = tid; n < cnt; n += kGenThreads) {
        double acc = 0.0;
        for (int k = 0; k < S; k++) acc += (double)sc.ola[n + k];
        o0[n] = (float)(acc * inv);
      }
      __syncthreads();
      float keep = 0.f;
      if (tid < S1) keep = sc.ola[cnt + tid];
      __syncthreads();
      if (tid < S1) sc.ola[tid] = keep;
    }
    __syncthreads();
    BF_PHASE(5);
  }
#ifdef BF_PHASE_TIMERS
  if (p.debug == 1 && blockIdx.x == 0 && tid == 0)
    printf("frames_kernel_n phases (clk/pair): pack %lld | fwd fft %lld | bins %lld | inv fft %lld | ola %lld | smoother %lld\n", ph_clk[0] / npairs,
           ph_clk[1] / npairs, ph_clk[2] / npairs, ph_clk[3] / npairs, ph_clk[4] / npairs, ph_clk[5] / npairs);
#endif
#undef BF_PHASE
  for (int i = tid; i < H; i += kGenThreads) p.tail[(size_t)s * H + i] = sc.tail[i];
  if (kSmooth)
    for (int i = tid; i < p.smooth_size - 1; i += kGenThreads) p.smooth_hist[(size_t)s * 64 + i] = sc.ola[i];
}

// =====================================================================================================================
// Magnitude-gated nodes (mvdr / lcmv / gss) at any supported frame size and up to 16 microphones.
//
//   reference path replaced: apply_weights of mvdr.cpp:62-115, lcmv.cpp:88-140, gss.cpp:96-156 (+ framing / OLA as above).
//
// The 1024-point, M <= 8 configurations run on sel_kernel.cu (register-resident 8x8 solves).  This kernel covers the
// rest of the parameter space with the same numerics contract: FP32 spectra, magnitude gate with a guard band whose
// bins are re-decided from an exact double DFT (bit-exact selected-bin set), solves in double with the matrices in
// per-thread local memory (run-time M), history ring in the layout of sel_kernel.cu ([B][P+2][M][Lsel], bin fastest).
// =====================================================================================================================
constexpr int kSelMaxM = 16;

template <int NN>
struct SelGenScratch {
  static constexpr int L = NN / 2 + 2;
  float2 y[2][L];
  unsigned char flag[2][L];
  unsigned short items[2 * L];
  unsigned short recheck[2 * L];
  int n_items, n_recheck;
  float esq[2][kSelMaxM];   // windowed frame energies per microphone (scale of the FP32 transform's absolute error)
  int nonfinite[2];
  float tail[NN / 2];
};

template <int NN>
__device__ __forceinline__ void unpack2_n(const float2* z, int l, float2& x0, float2& x1) {
  constexpr int L = NN / 2 + 2;
  const int j = (l == L - 1) ? NN / 2 - 1 : l;   // pseudo-bin: conj of bin N/2-1 (SURVEY B-4)
  const float2 a = z[swz(j)], b = z[swz((NN - j) & (NN - 1))];
  x0 = make_float2(a.x + b.x, a.y - b.y);    // Z[j] + conj(Z[N-j])
  x1 = make_float2(a.y + b.y, b.x - a.x);    // -i (Z[j] - conj(Z[N-j]))
  if (l == L - 1) { x0.y = -x0.y; x1.y = -x1.y; }
}

// FP64 re-decision of the magnitude gate for one (bin, frame): exact double DFT of that bin (mvdr.cpp:79-85), one warp.
template <int NN>
__device__ __noinline__ bool gate_fp64_n(const KernelParams& p, int s, int t, int l, int f, int lane) {
  constexpr int H = NN / 2, L = NN / 2 + 2;
  const int j = (l == L - 1) ? H + 1 : l;
  double stat = 0.0;
  for (int ch = 0; ch < p.M; ch++) {
    const int hf = t + f;
    const float* base = p.in + (size_t)s * p.in_stream_stride + (size_t)ch * p.in_mic_stride;
    const float* h0 = (hf - 1 < 0) ? p.prev_hop + ((size_t)s * p.M + ch) * H : base + (size_t)(hf - 1) * H;
    const float* h1 = base + (size_t)hf * H;
    double re = 0.0, im = 0.0;
    for (int n = lane; n < NN; n += 32) {
      const double xv = (double)(n < H ? h0[n] : h1[n - H]) * p.win_d[n];
      const double2 w = p.twid_d[(j * n) & (NN - 1)];
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
  stat /= (double)((unsigned)p.M * (unsigned)NN);
  return stat > p.thr_mag_d;
}

// One selected (bin, frame) of mvdr / lcmv in double, run-time M <= 16 and C <= 8 (mvdr: C = 1).
//   R = (P P^H) .* whiteR (mvdr.cpp:87, :239-243) = L L^H;  V = L^{-1} C, u = L^{-1} x, G = V^H V, b = V^H u;
//   y = g^H b with G g = e_0  (lcmv.cpp:111-119; for C = 1 this is mvdr.cpp:86-94: y = z^H u / z^H z).
__device__ __noinline__ float2 sel_item_d(const KernelParams& p, const float2* ring_l, int slot0, const float2* x, const float2* steer_l) {
  const int M = p.M, C = p.C, D = p.ring_depth;
  const size_t mic_stride = (size_t)p.Lsel, slot_stride = (size_t)M * p.Lsel;
  double2 A[kSelMaxM][kSelMaxM];   // lower triangle: covariance, then its Cholesky factor
  for (int i = 0; i < M; i++)
    for (int j = 0; j <= i; j++) A[i][j] = make_double2(0.0, 0.0);
  int slot = slot0;
  for (int k = 0; k < p.P; k++) {
    double2 h[kSelMaxM];
    const float2* src = ring_l + (size_t)slot * slot_stride;
    if (++slot == D) slot = 0;
    for (int i = 0; i < M; i++) { const float2 v = src[(size_t)i * mic_stride]; h[i] = make_double2((double)v.x, (double)v.y); }
    for (int i = 0; i < M; i++)
      for (int j = 0; j <= i; j++) {   // h_i conj(h_j)
        A[i][j].x = fma(h[i].x, h[j].x, fma(h[i].y, h[j].y, A[i][j].x));
        A[i][j].y = fma(h[i].y, h[j].x, fma(-h[i].x, h[j].y, A[i][j].y));
      }
  }
  double invd[kSelMaxM];
  for (int j = 0; j < M; j++) {
    double d = A[j][j].x * 1.001;   // whiteR diagonal (mvdr.cpp:242)
    for (int k = 0; k < j; k++) d = fma(-A[j][k].x, A[j][k].x, fma(-A[j][k].y, A[j][k].y, d));
    const double inv = 1.0 / sqrt(d);
    invd[j] = inv;
    for (int i = j + 1; i < M; i++) {
      double2 acc = A[i][j];
      for (int k = 0; k < j; k++) {   // acc -= L[i][k] conj(L[j][k])
        const double2 a = A[i][k], b = A[j][k];
        acc.x = fma(-a.x, b.x, fma(-a.y, b.y, acc.x));
        acc.y = fma(-a.y, b.x, fma(a.x, b.y, acc.y));
      }
      A[i][j] = make_double2(acc.x * inv, acc.y * inv);
    }
  }
  auto fwd = [&](double2* v) {   // v <- L^{-1} v
    for (int i = 0; i < M; i++) {
      double2 acc = v[i];
      for (int k = 0; k < i; k++) {
        const double2 l = A[i][k];
        acc.x = fma(-l.x, v[k].x, fma(l.y, v[k].y, acc.x));
        acc.y = fma(-l.x, v[k].y, fma(-l.y, v[k].x, acc.y));
      }
      v[i] = make_double2(acc.x * invd[i], acc.y * invd[i]);
    }
  };
  auto dotc = [&](const double2* a, const double2* b) {   // sum conj(a_i) b_i
    double2 r = make_double2(0.0, 0.0);
    for (int i = 0; i < M; i++) {
      r.x = fma(a[i].x, b[i].x, fma(a[i].y, b[i].y, r.x));
      r.y = fma(a[i].x, b[i].y, fma(-a[i].y, b[i].x, r.y));
    }
    return r;
  };
  double2 u[kSelMaxM];
  for (int i = 0; i < M; i++) u[i] = make_double2((double)x[i].x, (double)x[i].y);
  fwd(u);
  double2 V[kMaxCGen][kSelMaxM], G[kMaxCGen][kMaxCGen], b[kMaxCGen];
  for (int c = 0; c < C; c++) {
    for (int i = 0; i < M; i++) { const float2 a = steer_l[(size_t)c * M + i]; V[c][i] = make_double2((double)a.x, (double)a.y); }
    fwd(V[c]);
    b[c] = dotc(V[c], u);
    for (int c2 = 0; c2 <= c; c2++) G[c][c2] = dotc(V[c], V[c2]);
  }
  if (C == 1) {   // mvdr.cpp:90-94
    const double den = G[0][0].x;
    return make_float2((float)(b[0].x / den), (float)(b[0].y / den));
  }
  double gd[kMaxCGen];
  for (int j = 0; j < C; j++) {   // Cholesky of G
    double d = G[j][j].x;
    for (int k = 0; k < j; k++) d -= G[j][k].x * G[j][k].x + G[j][k].y * G[j][k].y;
    gd[j] = 1.0 / sqrt(d);
    for (int i = j + 1; i < C; i++) {
      double2 acc = G[i][j];
      for (int k = 0; k < j; k++) {
        const double2 a = G[i][k], bb = G[j][k];
        acc.x -= a.x * bb.x + a.y * bb.y;
        acc.y -= a.y * bb.x - a.x * bb.y;
      }
      G[i][j] = make_double2(acc.x * gd[j], acc.y * gd[j]);
    }
  }
  double2 q[kMaxCGen], g[kMaxCGen];
  for (int i = 0; i < C; i++) {   // q = Lg^{-1} e0
    double2 acc = make_double2(i == 0 ? 1.0 : 0.0, 0.0);
    for (int k = 0; k < i; k++) {
      const double2 l = G[i][k];
      acc.x -= l.x * q[k].x - l.y * q[k].y;
      acc.y -= l.x * q[k].y + l.y * q[k].x;
    }
    q[i] = make_double2(acc.x * gd[i], acc.y * gd[i]);
  }
  for (int i = C - 1; i >= 0; i--) {   // g = Lg^{-H} q
    double2 acc = q[i];
    for (int k = i + 1; k < C; k++) {
      const double2 l = G[k][i];
      acc.x -= l.x * g[k].x + l.y * g[k].y;
      acc.y -= l.x * g[k].y - l.y * g[k].x;
    }
    g[i] = make_double2(acc.x * gd[i], acc.y * gd[i]);
  }
  double2 yv = make_double2(0.0, 0.0);   // y = g^H b
  for (int c = 0; c < C; c++) {
    yv.x += g[c].x * b[c].x + g[c].y * b[c].y;
    yv.y += g[c].x * b[c].y - g[c].y * b[c].x;
  }
  return make_float2((float)yv.x, (float)yv.y);
}

template <int ALGO, int NN>
__global__ void __launch_bounds__(kGenThreads, 1) frames_kernel_sel(const __grid_constant__ KernelParams p) {
  constexpr int H = NN / 2, L = NN / 2 + 2;
  unsigned char* smem_raw = gen_smem_raw;
  // Spectra of all M microphones stay in shared memory when they fit (Mc == M).  Otherwise (e.g. 8 microphones at 4096
  // points: 256 KB) the transforms run Mc microphones at a time through the shared-memory tiles and the finished spectra go to
  // a per-stream workspace in global memory (L2-resident: it is re-read by the per-bin stage right away).
  const int Mc = p.sel_chunk;
  float2* ztile = reinterpret_cast<float2*>(smem_raw);                    // [Mc][NN] transform tiles
  float2* gbuf = ztile + (size_t)Mc * NN;                                  // [NN]
  SelGenScratch<NN>& sc = *reinterpret_cast<SelGenScratch<NN>*>(gbuf + NN);
  const unsigned g_off = (unsigned)((size_t)Mc * NN * sizeof(float2));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = p.M, D = p.ring_depth;
  const int s = blockIdx.x + p.stream_begin;
  float2* zall = (Mc < M) ? p.sel_ws + (size_t)s * M * NN : ztile;        // [M][NN] packed spectra
  const float2* tw = p.twid_f;
  const float* win = p.win_f;

  for (int i = tid; i < H; i += kGenThreads) sc.tail[i] = p.tail[(size_t)s * H + i];
  __syncthreads();
  const int nh = p.hop_end - p.hop_begin;
  const int npairs = (nh + 1) >> 1;
  for (int ip = 0; ip < npairs; ip++) {
    const int t = p.hop_begin + 2 * ip;
    const bool two = t + 1 < p.hop_end;
    const int nf = two ? 2 : 1;
    if (tid < 2 * kSelMaxM) (&sc.esq[0][0])[tid] = 0.f;
    if (tid == 0) { sc.n_items = 0; sc.n_recheck = 0; sc.nonfinite[0] = 0; sc.nonfinite[1] = 0; }
    __syncthreads();
    // ---- window + pack (util.h:217-242), frame energies ----
    for (int c0 = 0; c0 < M; c0 += Mc) {
      const int Mn = min(Mc, M - c0);
      for (int cc = 0; cc < Mn; cc++) {
        const int ch = c0 + cc;
        const float* base = p.in + (size_t)s * p.in_stream_stride + (size_t)ch * p.in_mic_stride;
        const float* ha = (t - 1 < 0) ? p.prev_hop + ((size_t)s * M + ch) * H : base + (size_t)(t - 1) * H;
        const float* hb = base + (size_t)t * H;
        const float* hc = two ? base + (size_t)(t + 1) * H : hb;
        float e0 = 0.f, e1 = 0.f;
        for (int n = tid; n < H; n += kGenThreads) {
          const float a = __ldg(ha + n), b = __ldg(hb + n), c = two ? __ldg(hc + n) : 0.f;
          const float w0 = 0.5f * __ldg(win + n), w1 = 0.5f * __ldg(win + n + H);
          const float2 z0 = make_float2(a * w0, (two ? b : 0.f) * w0), z1 = make_float2(b * w1, c * w1);
          ztile[(size_t)cc * NN + swz(n)] = z0;
          ztile[(size_t)cc * NN + swz(n + H)] = z1;
          e0 += z0.x * z0.x + z1.x * z1.x;
          e1 += z0.y * z0.y + z1.y * z1.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { e0 += __shfl_xor_sync(0xffffffffu, e0, o); e1 += __shfl_xor_sync(0xffffffffu, e1, o); }
        if (lane == 0) { atomicAdd(&sc.esq[0][ch], e0); atomicAdd(&sc.esq[1][ch], e1); }
      }
      __syncthreads();
      block_fft_fn<NN, -1, float2>(0u, Mn, tw, tid);
      if (Mc < M) {   // spill the chunk's spectra (same swizzled order) to the global workspace
        __syncthreads();
        for (int i = tid; i < Mn * NN; i += kGenThreads) zall[(size_t)c0 * NN + i] = ztile[i];
        __syncthreads();
      }
    }
    if (Mc < M) __threadfence_block();
    const int fr0 = (p.ring_slot0 + (t - p.hop_begin)) % D;   // ring slot of frame t
    // ---- B1: gate, history append, defaults ----
    float es0 = 0.f, es1 = 0.f;
    for (int ch = 0; ch < M; ch++) { es0 += 2.0f * sqrtf(sc.esq[0][ch]); es1 += 2.0f * sqrtf(sc.esq[1][ch]); }
    const float g0 = 2.0e-5f * es0 + 1.0e-6f * p.thr_mag, g1 = 2.0e-5f * es1 + 1.0e-6f * p.thr_mag;
    for (int l = tid; l < L; l += kGenThreads) {
      const bool inb = p.inband[l] != 0 && !(ALGO == ALGO_MVDR && l == 0);
      if (!inb) {   // out of band: 0 (mvdr.cpp:103), except mvdr's bin 0 which passes microphone 0 through (mvdr.cpp:76)
        float2 a = make_float2(0.f, 0.f), b = a;
        if (ALGO == ALGO_MVDR && l == 0) unpack2_n<NN>(zall, 0, a, b);
        sc.flag[0][l] = 0; sc.flag[1][l] = 0;
        sc.y[0][l] = a; sc.y[1][l] = b;
        continue;
      }
      float st0 = 0.f, st1 = 0.f;
      float2 x00 = make_float2(0.f, 0.f), x10 = x00;
      float2* ring_l = (ALGO != ALGO_GSS) ? p.hist + (size_t)s * D * M * p.Lsel + p.sel_slot[l] : nullptr;
      const int fs0 = fr0, fs1 = (fr0 + 1 == D) ? 0 : fr0 + 1;
      for (int ch = 0; ch < M; ch++) {
        float2 x0, x1;
        unpack2_n<NN>(zall + (size_t)ch * NN, l, x0, x1);
        if (ch == 0) { x00 = x0; x10 = x1; }
        st0 += sqrtf(fmaf(x0.x, x0.x, x0.y * x0.y));
        st1 += sqrtf(fmaf(x1.x, x1.x, x1.y * x1.y));
        if (ALGO != ALGO_GSS) {   // mvdr.cpp:99-101: every in-band bin appends every frame
          ring_l[((size_t)fs0 * M + ch) * p.Lsel] = x0;
          if (two) ring_l[((size_t)fs1 * M + ch) * p.Lsel] = x1;
        }
      }
      unsigned char f0 = 0, f1 = 0;
      if (fabsf(st0 - p.thr_mag) <= g0) sc.recheck[atomicAdd(&sc.n_recheck, 1)] = (unsigned short)(l * 2);
      else if (st0 > p.thr_mag) f0 = 1;
      if (two) {
        if (fabsf(st1 - p.thr_mag) <= g1) sc.recheck[atomicAdd(&sc.n_recheck, 1)] = (unsigned short)(l * 2 + 1);
        else if (st1 > p.thr_mag) f1 = 1;
      }
      sc.flag[0][l] = f0; sc.flag[1][l] = f1;
      sc.y[0][l] = make_float2(0.01f * x00.x, 0.01f * x00.y);   // mvdr.cpp:96 (overwritten when selected)
      sc.y[1][l] = make_float2(0.01f * x10.x, 0.01f * x10.y);
    }
    __syncthreads();
    // ---- B1b: FP64 re-decision of guarded bins ----
    for (int q = warp; q < sc.n_recheck; q += kGenThreads / 32) {
      const int l = sc.recheck[q] >> 1, f = sc.recheck[q] & 1;
      const bool sel = gate_fp64_n<NN>(p, s, t, l, f, lane);
      if (lane == 0 && sel) sc.flag[f][l] = 1;
    }
    __syncthreads();
    // ---- work list (ordered: neighbouring threads take neighbouring bins) ----
    for (int f = 0; f < (ALGO == ALGO_GSS ? 1 : nf); f++)
      for (int base = 0; base < L; base += kGenThreads) {
        const int l = base + tid;
        bool on = false;
        if (l < L) on = (ALGO == ALGO_GSS) ? ((sc.flag[0][l] | sc.flag[1][l]) != 0) : (sc.flag[f][l] != 0);
        const unsigned m = __ballot_sync(0xffffffffu, on);
        int pos = 0;
        if (lane == 0 && m) pos = atomicAdd(&sc.n_items, __popc(m));
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (on) sc.items[pos + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(l * 2 + f);
      }
    __syncthreads();
    // ---- B2: per-item solves ----
    for (int q = tid; q < sc.n_items; q += kGenThreads) {
      const int l = sc.items[q] >> 1, f = sc.items[q] & 1;
      const float2* steer_l = p.steer + (size_t)l * p.C * M;
      float2 x[kSelMaxM];
      if (ALGO == ALGO_GSS) {
        float2* Wg = p.gss_w + (size_t)s * BF_GSS_ROWS * M * p.Lsel + p.sel_slot[l];   // [B][BF_GSS_ROWS][M][Lsel]
        for (int ff = 0; ff < nf; ff++) {
          if (!sc.flag[ff][l]) continue;
          for (int ch = 0; ch < M; ch++) {
            float2 a, b;
            unpack2_n<NN>(zall + (size_t)ch * NN, l, a, b);
            x[ch] = ff ? b : a;
          }
          sc.y[ff][l] = gss_item<kMaxCGen>(p, Wg, (size_t)p.Lsel, x, steer_l);
        }
      } else {
        for (int ch = 0; ch < M; ch++) {
          float2 a, b;
          unpack2_n<NN>(zall + (size_t)ch * NN, l, a, b);
          x[ch] = f ? b : a;
        }
        int slot = (fr0 + f - p.P) % D;   // ring slot of frame (t+f) - P
        if (slot < 0) slot += D;
        sc.y[f][l] = sel_item_d(p, p.hist + (size_t)s * D * M * p.Lsel + p.sel_slot[l], slot, x, steer_l);
      }
    }
    __syncthreads();
    // ---- B3: non-finite frames (cold start, SURVEY B-10), Hermitian assembly, diagnostics ----
    {
      bool b0 = false, b1 = false;
      for (int l = tid; l < L; l += kGenThreads) {
        const float2 y0 = sc.y[0][l], y1 = sc.y[1][l];
        b0 |= !(isfinite(y0.x) && isfinite(y0.y));
        b1 |= two && !(isfinite(y1.x) && isfinite(y1.y));
      }
      if (b0) sc.nonfinite[0] = 1;
      if (b1) sc.nonfinite[1] = 1;
    }
    __syncthreads();
    const bool z0 = sc.nonfinite[0] != 0, z1 = sc.nonfinite[1] != 0;
    for (int l = tid; l < L; l += kGenThreads) {
      if (l <= H) {
        float2 y0 = z0 ? make_float2(0.f, 0.f) : sc.y[0][l], y1 = (two && !z1) ? sc.y[1][l] : make_float2(0.f, 0.f);
        if (l == H - 1) {   // Hermitian part of the asymmetric pair (N/2-1, N/2+1)
          const float2 p0 = z0 ? make_float2(0.f, 0.f) : sc.y[0][L - 1], p1 = (two && !z1) ? sc.y[1][L - 1] : make_float2(0.f, 0.f);
          y0 = make_float2(0.5f * (y0.x + p0.x), 0.5f * (y0.y - p0.y));
          y1 = make_float2(0.5f * (y1.x + p1.x), 0.5f * (y1.y - p1.y));
        }
        if (l == 0 || l == H) { y0.y = 0.f; y1.y = 0.f; }
        gbuf[swz(l)] = make_float2(y0.x - y1.y, y0.y + y1.x);
        if (l > 0 && l < H) gbuf[swz(NN - l)] = make_float2(y0.x + y1.y, y1.x - y0.y);
      }
      if (p.capture) {
        for (int f = 0; f < nf; f++) {
          unsigned char* cap = p.capture + (size_t)s * p.capture_stream_stride + (size_t)(t + f) * NN;
          const unsigned char fl = sc.flag[f][l];
          if (l <= H) {
            cap[l] = fl;
            if (l > 0 && l < H - 1) cap[NN - l] = fl;
          } else {
            cap[H + 1] = fl;
          }
        }
      }
    }
    __syncthreads();
    const float2* res = reinterpret_cast<const float2*>(smem_raw + block_fft_oop_fn<NN, 1, float2>(g_off, 0u, tw, tid));
    // ---- synthesis window, overlap-add (util.h:244-253, 301-302); a poisoned frame turns its samples into NaN ----
    const float bad0 = z0 ? __int_as_float(0x7fc00000) : 0.f, bad1 = z1 ? __int_as_float(0x7fc00000) : 0.f;
    float* o0 = p.out + (size_t)s * p.out_stream_stride + (size_t)t * H;
    for (int n = tid; n < H; n += kGenThreads) {
      const float w0 = __ldg(win + n) * p.out_scale, w1 = __ldg(win + n + H) * p.out_scale;
      const float2 a = res[swz(n)], b = res[swz(n + H)];
      o0[n] = sc.tail[n] + (a.x * w0 + bad0);
      if (two) {
        o0[H + n] = (b.x * w1 + bad0) + (a.y * w0 + bad1);
        sc.tail[n] = b.y * w1 + bad1;
      } else {
        sc.tail[n] = b.x * w1 + bad0;
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < H; i += kGenThreads) p.tail[(size_t)s * H + i] = sc.tail[i];
}

template <int NN>
static size_t sel_gen_smem(int M) { return sizeof(float2) * ((size_t)M + 1) * NN + sizeof(SelGenScratch<NN>) + 16; }

size_t frames_kernel_sel_smem(int N, int M);
// microphones transformed per pass: all of them when their spectra fit the shared memory, else as many as fit
int frames_kernel_sel_chunk(int N, int M) {
  int mc = M;
  while (mc > 1 && frames_kernel_sel_smem(N, mc) > 232448) mc--;
  return mc;
}

size_t frames_kernel_sel_smem(int N, int M) {
  switch (N) {
    case 512: return sel_gen_smem<512>(M);
    case 1024: return sel_gen_smem<1024>(M);
    case 2048: return sel_gen_smem<2048>(M);
    case 4096: return sel_gen_smem<4096>(M);
  }
  return ~(size_t)0;
}

template <int ALGO, int NN>
static cudaError_t launch_sel_n(const KernelParams& p, cudaStream_t st) {
  const size_t smem = sel_gen_smem<NN>(p.sel_chunk);
  cudaError_t e = cudaFuncSetAttribute(frames_kernel_sel<ALGO, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  frames_kernel_sel<ALGO, NN><<<p.n_streams, kGenThreads, smem, st>>>(p);
  return cudaGetLastError();
}
template <int ALGO>
static cudaError_t launch_sel_algo(const KernelParams& p, cudaStream_t st) {
  switch (p.N) {
    case 512: return launch_sel_n<ALGO, 512>(p, st);
    case 1024: return launch_sel_n<ALGO, 1024>(p, st);
    case 2048: return launch_sel_n<ALGO, 2048>(p, st);
    case 4096: return launch_sel_n<ALGO, 4096>(p, st);
  }
  return cudaErrorNotSupported;
}
cudaError_t launch_frames_kernel_sel(int algo, const KernelParams& p, cudaStream_t st) {
  if (p.M > kSelMaxM || p.C > kMaxCGen) return cudaErrorNotSupported;
  switch (algo) {
    case ALGO_MVDR: return launch_sel_algo<ALGO_MVDR>(p, st);
    case ALGO_LCMV: return launch_sel_algo<ALGO_LCMV>(p, st);
    case ALGO_GSS: return launch_sel_algo<ALGO_GSS>(p, st);
  }
  return cudaErrorNotSupported;
}

// =====================================================================================================================
// Stand-alone MCRA noise-reduction node: mcra.cpp:62-155 on the FIRST microphone only.
//   |X|^2 -> 3-tap smoothing over neighbouring bins (this node does read its neighbours, unlike phasempf's copy,
//   SURVEY B-9) -> recursive averaging -> minima tracking over windows of L frames -> noise estimate lambda ->
//   Y = max(0, |X| - sqrt(lambda)) * out_amp * e^{i arg X}  (or the noise estimate itself); Y[0] is never written
//   (mcra.cpp:127 has the one-past-the-end write of SURVEY B-5).
// The recursion is conjugate-symmetric (|X[N-j]| = |X[j]|, symmetric window, symmetric index clipping at 1 and N-1),
// so the half spectrum carries it; no frequency table is involved, hence no pseudo-bin.
// State [B][7][L] (layout shared with phasempf; slots 0-3: S_prev, S_tmp, S_min, lambda).
// =====================================================================================================================
template <int NN>
struct McraScratch {
  float psq[2][NN / 2 + 2];   // |X_f[l]|^2, l = 0..N/2
  float tail[NN / 2];
};

template <int NN>
__global__ void __launch_bounds__(kGenThreads, 1) frames_kernel_mcra(const __grid_constant__ KernelParams p) {
  constexpr int H = NN / 2, L = NN / 2 + 2;
  unsigned char* smem_raw = gen_smem_raw;
  float2* zbuf = reinterpret_cast<float2*>(smem_raw);   // [NN] packed spectrum of microphone 0
  float2* gbuf = zbuf + NN;                              // [NN]
  McraScratch<NN>& sc = *reinterpret_cast<McraScratch<NN>*>(gbuf + NN);
  const unsigned g_off = (unsigned)(NN * sizeof(float2));
  const int tid = threadIdx.x;
  const int s = blockIdx.x + p.stream_begin;
  const float2* tw = p.twid_f;
  const float* win = p.win_f;
  int cur_L = p.mcra_cur_L0, first_L = p.mcra_first0;
  float* stg = p.mpf_state + (size_t)s * 7 * L;

  for (int i = tid; i < H; i += kGenThreads) sc.tail[i] = p.tail[(size_t)s * H + i];
  __syncthreads();
  const int nh = p.hop_end - p.hop_begin;
  const int npairs = (nh + 1) >> 1;
  for (int ip = 0; ip < npairs; ip++) {
    const int t = p.hop_begin + 2 * ip;
    const bool two = t + 1 < p.hop_end;
    const int nf = two ? 2 : 1;
    {   // window + pack (util.h:217-242), first channel only
      const float* base = p.in + (size_t)s * p.in_stream_stride;
      const float* ha = (t - 1 < 0) ? p.prev_hop + (size_t)s * p.M * H : base + (size_t)(t - 1) * H;
      const float* hb = base + (size_t)t * H;
      const float* hc = two ? base + (size_t)(t + 1) * H : hb;
      for (int n = tid; n < H; n += kGenThreads) {
        const float a = __ldg(ha + n), b = __ldg(hb + n), c = two ? __ldg(hc + n) : 0.f;
        const float w0 = 0.5f * __ldg(win + n), w1 = 0.5f * __ldg(win + n + H);
        zbuf[swz(n)] = make_float2(a * w0, (two ? b : 0.f) * w0);
        zbuf[swz(n + H)] = make_float2(b * w1, c * w1);
      }
    }
    __syncthreads();
    block_fft_fn<NN, -1, float2>(0u, 1, tw, tid);
    for (int l = tid; l <= H; l += kGenThreads) {   // in_fft_square (mcra.cpp:73-76)
      const float2 a = zbuf[swz(l)], b = zbuf[swz((NN - l) & (NN - 1))];
      const float2 x0 = make_float2(a.x + b.x, a.y - b.y), x1 = make_float2(a.y + b.y, b.x - a.x);
      sc.psq[0][l] = fmaf(x0.x, x0.x, x0.y * x0.y);
      sc.psq[1][l] = fmaf(x1.x, x1.x, x1.y * x1.y);
    }
    __syncthreads();
    // window bookkeeping of the two frames (mcra.cpp:100-113), global per frame
    bool reset_f[2] = {false, false};
    float inv_cl_f[2] = {1.f, 1.f};
    int fst_f[2] = {first_L, first_L};
    for (int f = 0; f < nf; f++) {
      const bool r = cur_L > p.mcra_L;
      if (r) { cur_L = 1; first_L = 0; } else { cur_L++; }
      reset_f[f] = r; inv_cl_f[f] = 1.0f / (float)cur_L; fst_f[f] = first_L;
    }
    for (int l = tid; l <= H; l += kGenThreads) {
      const float2 a = zbuf[swz(l)], b = zbuf[swz((NN - l) & (NN - 1))];
      const float2 xf[2] = {make_float2(a.x + b.x, a.y - b.y), make_float2(a.y + b.y, b.x - a.x)};
      float S_prev = stg[0 * L + l], S_tmp = stg[1 * L + l], S_min = stg[2 * L + l], lam = stg[3 * L + l];
      float2 yy[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      for (int f = 0; f < nf; f++) {
        const float* ps = sc.psq[f];
        const float pq = ps[l];
        float Sf;
        if (l == 0) {
          Sf = sqrtf(pq);   // "passing on the DC component": abs, not squared (mcra.cpp:81)
        } else {
          // neighbours j-1, j, j+1 clipped to [1, N-1]; bin N/2+1 is the mirror of N/2-1
          const float lo = (l - 1 >= 1) ? ps[l - 1] : 0.f;
          const float hi = (l + 1 <= H) ? ps[l + 1] : ps[H - 1];
          Sf = 0.25f * lo + 0.5f * pq + 0.25f * hi;
        }
        const float S = p.mcra_alphaS * S_prev + (1.0f - p.mcra_alphaS) * Sf;
        if (reset_f[f]) { S_min = fminf(S_tmp, S); S_tmp = S; }
        else { S_min = fminf(S_min, S); S_tmp = fminf(S_tmp, S); }
        if (fst_f[f] || S < S_min * p.mcra_delta || lam > pq) {
          if (fst_f[f] && inv_cl_f[f] > p.mcra_alphaD) lam = inv_cl_f[f] * lam + (1.0f - inv_cl_f[f]) * pq;
          else lam = p.mcra_alphaD2 * lam + (1.0f - p.mcra_alphaD) * pq;
        }
        S_prev = S;
        if (l > 0) {
          const float mag_x = sqrt_fast(pq);
          float mag;
          if (p.out_only_noise) {
            mag = sqrt_fast(lam) * p.out_amp;
          } else {
            mag = (mag_x - sqrt_fast(lam)) * p.out_amp;
            if (mag < 0.f) mag = 0.f;
          }
          const float r0 = rsqrtf(pq);
          const float2 unit = pq > 0.f ? make_float2(xf[f].x * r0, xf[f].y * r0) : make_float2(1.f, 0.f);   // e^{i arg X}
          yy[f] = make_float2(mag * unit.x, mag * unit.y);
        }
      }
      stg[0 * L + l] = S_prev; stg[1 * L + l] = S_tmp; stg[2 * L + l] = S_min; stg[3 * L + l] = lam;
      float2 y0 = yy[0], y1 = yy[1];
      if (l == 0 || l == H) { y0.y = 0.f; y1.y = 0.f; }
      gbuf[swz(l)] = make_float2(y0.x - y1.y, y0.y + y1.x);
      if (l > 0 && l < H) gbuf[swz(NN - l)] = make_float2(y0.x + y1.y, y1.x - y0.y);
    }
    __syncthreads();
    const float2* res = reinterpret_cast<const float2*>(smem_raw + block_fft_oop_fn<NN, 1, float2>(g_off, 0u, tw, tid));
    float* o0 = p.out + (size_t)s * p.out_stream_stride + (size_t)t * H;
    for (int n = tid; n < H; n += kGenThreads) {   // util.h:244-253, 301-302
      const float w0 = __ldg(win + n) * p.out_scale, w1 = __ldg(win + n + H) * p.out_scale;
      const float2 a = res[swz(n)], b = res[swz(n + H)];
      o0[n] = sc.tail[n] + a.x * w0;
      if (two) {
        o0[H + n] = b.x * w1 + a.y * w0;
        sc.tail[n] = b.y * w1;
      } else {
        sc.tail[n] = b.x * w1;
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < H; i += kGenThreads) p.tail[(size_t)s * H + i] = sc.tail[i];
}

template <int NN>
static cudaError_t launch_mcra_n(const KernelParams& p, cudaStream_t st) {
  const size_t smem = sizeof(float2) * 2 * NN + sizeof(McraScratch<NN>) + 16;
  cudaError_t e = cudaFuncSetAttribute(frames_kernel_mcra<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  frames_kernel_mcra<NN><<<p.n_streams, kGenThreads, smem, st>>>(p);
  return cudaGetLastError();
}
cudaError_t launch_frames_kernel_mcra(const KernelParams& p, cudaStream_t st) {
  switch (p.N) {
    case 512: return launch_mcra_n<512>(p, st);
    case 1024: return launch_mcra_n<1024>(p, st);
    case 2048: return launch_mcra_n<2048>(p, st);
    case 4096: return launch_mcra_n<4096>(p, st);
  }
  return cudaErrorNotSupported;
}

// =====================================================================================================================
// rosjack_ref: jack_ref.cpp:19-60 + util.h:347-372.  The first microphone, windowed twice
// (analysis in overlap_and_add_prepare_input, then jack_ref.cpp:28) and overlap-added: both terms of an output sample
// come from the same input sample of the PREVIOUS hop,
//     out[t][j] = fl32(fl32(x w[j+H]) w[j+H]) + fl32(fl32(x w[j]) w[j]),   x = hop_{t-1}[j], products in double,
// i.e. the input delayed by one hop up to float rounding (w^2[j] + w^2[j+H] = 1).  Pure streaming: 8 bytes per sample.
// =====================================================================================================================
__global__ void __launch_bounds__(256) ref_kernel(const __grid_constant__ KernelParams p) {
  // A thread owns four consecutive sample positions j..j+3 of the hop and walks (stream, hop) pairs with them, so its
  // eight window values stay in registers; a warp covers 128 consecutive samples (one 512-byte run per load / store).
  const int H = p.H;
  const int nh = p.hop_end - p.hop_begin;
  const int quads = H / 4;                                   // threads that tile one hop
  const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
  const int j = (int)(gtid % quads) * 4;
  const long long row0 = gtid / quads, row_step = gsz / quads;   // gsz is a multiple of quads (launch)
  const double4 wa = make_double4(p.win_d[j], p.win_d[j + 1], p.win_d[j + 2], p.win_d[j + 3]);
  const double4 wb = make_double4(p.win_d[j + H], p.win_d[j + H + 1], p.win_d[j + H + 2], p.win_d[j + H + 3]);
  auto one = [](float xv, double w0, double w1) {
    const double xd = (double)xv;
    const float c = (float)((double)(float)(xd * w0) * w0);   // this frame, first half
    const float q = (float)((double)(float)(xd * w1) * w1);   // previous frame, second half
    return q + c;                                               // util.h:362
  };
  const long long rows = (long long)p.n_streams * nh;          // (stream, hop) pairs of this launch
  auto src_of = [&](long long r) {
    const int sl = (int)(r / nh), t = p.hop_begin + (int)(r - (long long)sl * nh);
    const int s = sl + p.stream_begin;
    return (t - 1 < 0) ? p.prev_hop + (size_t)s * p.M * H + j : p.in + (size_t)s * p.in_stream_stride + (size_t)(t - 1) * H + j;
  };
  auto dst_of = [&](long long r) {
    const int sl = (int)(r / nh), t = p.hop_begin + (int)(r - (long long)sl * nh);
    return p.out + (size_t)(sl + p.stream_begin) * p.out_stream_stride + (size_t)t * H + j;
  };
  constexpr int kU = 1;   // rows in flight per thread: more did not help (4: 52.8 % vs 55.3 % of HBM peak), the FP32<->FP64 conversions bound the kernel
  for (long long r = row0; r < rows; r += kU * row_step) {
    float4 x[kU];
#pragma unroll
    for (int u = 0; u < kU; u++)
      if (r + u * row_step < rows) x[u] = __ldcs(reinterpret_cast<const float4*>(src_of(r + u * row_step)));   // streamed once
#pragma unroll
    for (int u = 0; u < kU; u++) {
      if (r + u * row_step >= rows) break;
      float4 y;
      y.x = one(x[u].x, wa.x, wb.x); y.y = one(x[u].y, wa.y, wb.y); y.z = one(x[u].z, wa.z, wb.z); y.w = one(x[u].w, wa.w, wb.w);
      __stcs(reinterpret_cast<float4*>(dst_of(r + u * row_step)), y);
    }
  }
}
cudaError_t launch_ref_kernel(const KernelParams& p, cudaStream_t st) {
  // float4 path: 16-byte aligned bases and strides (the C ABI's staging buffers are; a caller's odd layout is rejected)
  if ((reinterpret_cast<uintptr_t>(p.in) & 15) || (reinterpret_cast<uintptr_t>(p.out) & 15) || (p.in_stream_stride & 3) || (p.out_stream_stride & 3))
    return cudaErrorMisalignedAddress;
  // grid threads = a multiple of H/4 (the threads that tile one hop): H/4 is 64..512, 148*8*256 threads are a multiple of 512
  ref_kernel<<<148 * 8, 256, 0, st>>>(p);
  return cudaGetLastError();
}

// =====================================================================================================================
// Generalized sidelobe canceller, gsc.cpp:54-197, in two kernels:
//   gsc_align_kernel   do_overlap_bymic + apply_weights (gsc.cpp:54-75, util.h:347-379): every microphone is windowed,
//                      transformed, phase-aligned towards the look direction (x_fft *= conj(w), no 1/M), transformed
//                      back and overlap-added on its own -> aligned[s][m][sample]
//   gsc_nlms_kernel    the per-sample loop (gsc.cpp:113-183): fixed beamformer (mean), blocking matrix (differences of
//                      neighbouring microphones), one adaptive FIR per blocking channel, NLMS update with the step
//                      normalised by the output or the blocking-channel power.  Sequential in time; one WARP per
//                      stream, taps spread over the lanes, delay lines as rings in shared memory.
// Float arithmetic of the reference is kept per element (separate multiply and add, same expression order); only
// the three F-term sums per sample (filter output and the two powers) are formed as per-lane partial sums + a
// butterfly instead of one sequential chain.
// =====================================================================================================================
template <int NN>
__global__ void __launch_bounds__(kGenThreads, 1) gsc_align_kernel(const __grid_constant__ KernelParams p) {
  constexpr int H = NN / 2;
  unsigned char* smem_raw = gen_smem_raw;
  float2* zall = reinterpret_cast<float2*>(smem_raw);                       // [M][NN]
  float* tails = reinterpret_cast<float*>(zall + (size_t)p.M * NN);          // [M][H]
  const int tid = threadIdx.x, M = p.M;
  const int s = blockIdx.x + p.stream_begin;
  const float2* tw = p.twid_f;
  const float* win = p.win_f;
  for (int i = tid; i < M * H; i += kGenThreads) tails[i] = p.gsc_tail[(size_t)s * M * H + i];
  __syncthreads();
  const int nh = p.hop_end - p.hop_begin;
  const int npairs = (nh + 1) >> 1;
  const long long mic_stride = (long long)nh * H;
  for (int ip = 0; ip < npairs; ip++) {
    const int t = p.hop_begin + 2 * ip;
    const bool two = t + 1 < p.hop_end;
    for (int ch = 0; ch < M; ch++) {   // window + pack (util.h:217-242)
      const float* base = p.in + (size_t)s * p.in_stream_stride + (size_t)ch * p.in_mic_stride;
      const float* ha = (t - 1 < 0) ? p.prev_hop + ((size_t)s * M + ch) * H : base + (size_t)(t - 1) * H;
      const float* hb = base + (size_t)t * H;
      const float* hc = two ? base + (size_t)(t + 1) * H : hb;
      for (int n = tid; n < H; n += kGenThreads) {
        const float a = __ldg(ha + n), b = __ldg(hb + n), c = two ? __ldg(hc + n) : 0.f;
        const float w0 = 0.5f * __ldg(win + n), w1 = 0.5f * __ldg(win + n + H);
        zall[(size_t)ch * NN + swz(n)] = make_float2(a * w0, (two ? b : 0.f) * w0);
        zall[(size_t)ch * NN + swz(n + H)] = make_float2(b * w1, c * w1);
      }
    }
    __syncthreads();
    block_fft_fn<NN, -1, float2>(0u, M, tw, tid);
    // x_fft[j] *= conj(weights[mic][j]) (gsc.cpp:62-65); the real part taken afterwards keeps the Hermitian part, which
    // lets both frames of the pair share the transform: G = ceff .* Z, ceff = (conj(w_j) + w_{N-j}) / 2 (host, double)
    for (int e = tid; e < M * NN; e += kGenThreads) {
      const int ch = e / NN, j = e - ch * NN;
      const float2 z = zall[(size_t)ch * NN + swz(j)];
      const float2 w = __ldg(p.das_ceff + (size_t)ch * NN + j);
      zall[(size_t)ch * NN + swz(j)] = make_float2(2.0f * (z.x * w.x - z.y * w.y), 2.0f * (z.x * w.y + z.y * w.x));   // Z carries the 0.5 of the packing
    }
    __syncthreads();
    block_fft_fn<NN, 1, float2>(0u, M, tw, tid);
    for (int ch = 0; ch < M; ch++) {   // util.h:244-253 + the per-microphone overlap-add of util.h:361-363
      float* o0 = p.gsc_aligned + (size_t)(s - p.stream_begin) * p.gsc_aligned_stream_stride + (size_t)ch * mic_stride + (size_t)(t - p.hop_begin) * H;
      const float2* res = zall + (size_t)ch * NN;
      float* tl = tails + ch * H;
      for (int n = tid; n < H; n += kGenThreads) {
        const float w0 = __ldg(win + n) * p.out_scale, w1 = __ldg(win + n + H) * p.out_scale;
        const float2 a = res[swz(n)], b = res[swz(n + H)];
        o0[n] = tl[n] + a.x * w0;
        if (two) {
          o0[H + n] = b.x * w1 + a.y * w0;
          tl[n] = b.y * w1;
        } else {
          tl[n] = b.x * w1;
        }
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < M * H; i += kGenThreads) p.gsc_tail[(size_t)s * M * H + i] = tails[i];
}

constexpr int kNlmsWarps = 4;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kNlmsWarps * 32) gsc_nlms_kernel(const __grid_constant__ KernelParams p) {
  extern __shared__ __align__(16) float nlms_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sl = blockIdx.x * kNlmsWarps + warp;   // stream inside this launch
  if (sl >= p.n_streams) return;
  const int s = sl + p.stream_begin;
  const int M = p.M, M1 = M - 1, F = p.gsc_F, Q = F / 32;
  const int per_warp = (2 * M1 + 1) * F + M * 32 + 32;
  float* blk = nlms_smem + (size_t)warp * per_warp;   // [M1][F] blocking-matrix delay lines (rings)
  float* flt = blk + M1 * F;                           // [M1][F] filters, by logical tap
  float* lst = flt + M1 * F;                           // [F] last outputs (ring)
  float* tile = lst + F;                               // [M][32] aligned samples of the current block of 32
  float* otile = tile + M * 32;                        // [32] outputs of the block
  float* st = p.gsc_state + (size_t)s * (2 * M1 + 1) * F;
  for (int i = lane; i < (2 * M1 + 1) * F; i += 32) blk[i] = st[i];
  int head = p.gsc_head[s];
  __syncwarp();
  const int nh = p.hop_end - p.hop_begin;
  const long long n_samples = (long long)nh * p.H, mic_stride = n_samples;
  const float* al = p.gsc_aligned + (size_t)sl * p.gsc_aligned_stream_stride;
  float* out = p.out + (size_t)s * p.out_stream_stride + (size_t)p.hop_begin * p.H;
  const float fM = (float)M, fF = (float)F, invF = 1.0f / (float)F;
  const bool pow2F = (F & (F - 1)) == 0;   // then the division by F is an exact scaling
  for (long long base = 0; base < n_samples; base += 32) {
    for (int i = 0; i < M; i++) tile[i * 32 + lane] = al[(size_t)i * mic_stride + base + lane];
    __syncwarp();
    for (int jj = 0; jj < 32; jj++) {
      // fixed beamformer (gsc.cpp:116-121) and the new blocking-matrix samples (gsc.cpp:124)
      float das = 0.0f;
      for (int i = 0; i < M; i++) das = __fadd_rn(das, tile[i * 32 + jj]);
      float o = __fdiv_rn(das, fM);
      if (lane < M1) blk[lane * F + head] = __fsub_rn(tile[(lane + 1) * 32 + jj], tile[lane * 32 + jj]);   // shift_data: the oldest slot takes the new sample
      const int hd = (head + 1 == F) ? 0 : head + 1;   // ring position of logical tap 0 after the shift
      float keep_pw = 0.f;
      __syncwarp();
      // filter outputs and blocking-channel powers (gsc.cpp:127-131, 84-91)
      for (int i = 0; i < M1; i++) {
        float dot = 0.f, pw = 0.f;
        for (int q = 0; q < Q; q++) {
          const int k = lane + 32 * q;
          int ph = hd + k; if (ph >= F) ph -= F;
          const float b = blk[i * F + ph];
          dot = fmaf(flt[i * F + k], b, dot);
          pw = fmaf(b, b, pw);
        }
        dot = warp_sum(dot);
        pw = warp_sum(pw);
        o = __fsub_rn(o, dot);   // out[j] -= block_out
        if (lane == i) keep_pw = pw;   // the power of channel i is needed once the output is known: parked in lane i
      }
      // last outputs and their power (gsc.cpp:138-140)
      if (lane == 0) lst[head] = o;
      __syncwarp();
      float lp = 0.f;
      for (int q = 0; q < Q; q++) { const float v = lst[lane + 32 * q]; lp = fmaf(v, v, lp); }
      lp = warp_sum(lp);
      const float last_out_power = __fsqrt_rn(pow2F ? lp * invF : __fdiv_rn(lp, fF));
      if ((double)last_out_power < p.gsc_vad_threshold || !p.gsc_use_vad) {
        // step sizes (gsc.cpp:147-156): lane i owns channel i, so the double-precision quotients are formed once per
        // sample instead of once per channel by every lane.  mu0 / last_out_power is shared by all channels.
        const float bpw = keep_pw;
        const float block_power = __fsqrt_rn(pow2F ? bpw * invF : __fdiv_rn(bpw, fF));
        float my_mu;
        if (p.gsc_mu0 * (double)block_power / (double)last_out_power < p.gsc_mu_max) my_mu = (float)(p.gsc_mu0 / (double)last_out_power);
        else my_mu = (float)(p.gsc_mu0 / (double)block_power);
        if (isnan(my_mu) || isinf(my_mu)) my_mu = 0.0f;
        for (int i = 0; i < M1; i++) {
          const float this_mu = __shfl_sync(0xffffffffu, my_mu, i);
          const float g = __fmul_rn(this_mu, o);
          for (int q = 0; q < Q; q++) {
            const int k = lane + 32 * q;
            int ph = hd + k; if (ph >= F) ph -= F;
            float f = __fadd_rn(flt[i * F + k], __fmul_rn(g, blk[i * F + ph]));   // filter[i][k] += this_mu*out[j]*block_matrix[i][k]
            if (isnan(f)) f = 0.0f;
            flt[i * F + k] = f;
          }
        }
      }
      if (lane == 0) otile[jj] = o;
      head = hd;
      __syncwarp();
    }
    out[base + lane] = otile[lane];
    __syncwarp();
  }
  // persist: rings are stored as they are, with their head
  for (int i = lane; i < (2 * M1 + 1) * F; i += 32) st[i] = blk[i];
  if (lane == 0) p.gsc_head[s] = head;
}

// Fast path of the NLMS for the launch-file shape (filter_size 128, 2-4 microphones): compile-time channel count, the
// lane's four taps of every filter and delay line in registers between the filter output and the update, ring positions
// computed once per sample, and the 2(M-1) sums of a sample reduced by one interleaved butterfly.  Same arithmetic per
// element and the same state layout as gsc_nlms_kernel.
template <int MC>
__global__ void __launch_bounds__(kNlmsWarps * 32, 8) gsc_nlms_fast_kernel(const __grid_constant__ KernelParams p) {
  constexpr int F = 128, Q = 4, M = MC + 1;
  extern __shared__ __align__(16) float nlms_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sl = blockIdx.x * kNlmsWarps + warp;
  if (sl >= p.n_streams) return;
  const int s = sl + p.stream_begin;
  constexpr int per_warp = (MC + 1) * F + M * 32 + 32;
  float* blk = nlms_smem + (size_t)warp * per_warp;   // [MC][F] blocking-matrix delay lines (rings)
  float* lst = blk + MC * F;                           // [F] last outputs (ring)
  float* tile = lst + F;                               // [M][32]
  float* otile = tile + M * 32;                        // [32]
  float* st = p.gsc_state + (size_t)s * (2 * MC + 1) * F;
  for (int i = lane; i < MC * F; i += 32) blk[i] = st[i];
  for (int i = lane; i < F; i += 32) lst[i] = st[2 * MC * F + i];
  float flt[MC][Q];
#pragma unroll
  for (int i = 0; i < MC; i++)
#pragma unroll
    for (int q = 0; q < Q; q++) flt[i][q] = st[(MC + i) * F + lane + 32 * q];
  int head = p.gsc_head[s];
  __syncwarp();
  const int nh = p.hop_end - p.hop_begin;
  const long long n_samples = (long long)nh * p.H, mic_stride = n_samples;
  const float* al = p.gsc_aligned + (size_t)sl * p.gsc_aligned_stream_stride;
  float* out = p.out + (size_t)s * p.out_stream_stride + (size_t)p.hop_begin * p.H;
  const float fM = (float)M, invF = 1.0f / (float)F;
  const int my_ch = lane < MC ? lane : 0;
  for (long long base = 0; base < n_samples; base += 32) {
#pragma unroll
    for (int i = 0; i < M; i++) tile[i * 32 + lane] = al[(size_t)i * mic_stride + base + lane];
    __syncwarp();
#pragma unroll 1
    for (int jj = 0; jj < 32; jj++) {
      float a[M];
#pragma unroll
      for (int i = 0; i < M; i++) a[i] = tile[i * 32 + jj];
      float das = 0.0f;
#pragma unroll
      for (int i = 0; i < M; i++) das = __fadd_rn(das, a[i]);   // gsc.cpp:116-121
      float o = __fdiv_rn(das, fM);
      if (lane < MC) blk[lane * F + head] = __fsub_rn(tile[(lane + 1) * 32 + jj], tile[lane * 32 + jj]);   // gsc.cpp:124 (shift_data)
      const int hd = (head + 1) & (F - 1);
      __syncwarp();
      float b[MC][Q], dot[MC], pw[MC];
#pragma unroll
      for (int i = 0; i < MC; i++) {
        dot[i] = 0.f; pw[i] = 0.f;
#pragma unroll
        for (int q = 0; q < Q; q++) {
          b[i][q] = blk[i * F + ((hd + lane + 32 * q) & (F - 1))];
          dot[i] = fmaf(flt[i][q], b[i][q], dot[i]);
          pw[i] = fmaf(b[i][q], b[i][q], pw[i]);
        }
      }
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1)
#pragma unroll
        for (int i = 0; i < MC; i++) {
          dot[i] += __shfl_xor_sync(0xffffffffu, dot[i], sh);
          pw[i] += __shfl_xor_sync(0xffffffffu, pw[i], sh);
        }
#pragma unroll
      for (int i = 0; i < MC; i++) o = __fsub_rn(o, dot[i]);   // out[j] -= block_out (gsc.cpp:131-134)
      if (lane == 0) lst[head] = o;                             // gsc.cpp:138
      __syncwarp();
      float lp = 0.f;
#pragma unroll
      for (int q = 0; q < Q; q++) { const float v = lst[lane + 32 * q]; lp = fmaf(v, v, lp); }
      lp = warp_sum(lp);
      const float last_out_power = __fsqrt_rn(lp * invF);
      if ((double)last_out_power < p.gsc_vad_threshold || !p.gsc_use_vad) {
        // step sizes (gsc.cpp:147-156): lane i forms the quotients of channel i
        float my_pw = pw[0];
#pragma unroll
        for (int i = 1; i < MC; i++) my_pw = (my_ch == i) ? pw[i] : my_pw;
        const float block_power = __fsqrt_rn(my_pw * invF);
        float my_mu;
        if (p.gsc_mu0 * (double)block_power / (double)last_out_power < p.gsc_mu_max) my_mu = (float)(p.gsc_mu0 / (double)last_out_power);
        else my_mu = (float)(p.gsc_mu0 / (double)block_power);
        if (isnan(my_mu) || isinf(my_mu)) my_mu = 0.0f;
#pragma unroll
        for (int i = 0; i < MC; i++) {
          const float g = __fmul_rn(__shfl_sync(0xffffffffu, my_mu, i), o);
#pragma unroll
          for (int q = 0; q < Q; q++) {
            float f = __fadd_rn(flt[i][q], __fmul_rn(g, b[i][q]));   // filter[i][k] += this_mu*out[j]*block_matrix[i][k]
            if (isnan(f)) f = 0.0f;
            flt[i][q] = f;
          }
        }
      }
      if (lane == 0) otile[jj] = o;
      head = hd;
      __syncwarp();
    }
    out[base + lane] = otile[lane];
    __syncwarp();
  }
  for (int i = lane; i < MC * F; i += 32) st[i] = blk[i];
  for (int i = lane; i < F; i += 32) st[2 * MC * F + i] = lst[i];
#pragma unroll
  for (int i = 0; i < MC; i++)
#pragma unroll
    for (int q = 0; q < Q; q++) st[(MC + i) * F + lane + 32 * q] = flt[i][q];
  if (lane == 0) p.gsc_head[s] = head;
}

template <int MC>
static cudaError_t launch_nlms_fast(const KernelParams& p, cudaStream_t st) {
  const size_t smem = sizeof(float) * kNlmsWarps * ((size_t)(MC + 1) * 128 + (size_t)(MC + 1) * 32 + 32);
  gsc_nlms_fast_kernel<MC><<<(p.n_streams + kNlmsWarps - 1) / kNlmsWarps, kNlmsWarps * 32, smem, st>>>(p);
  return cudaGetLastError();
}

template <int NN>
static cudaError_t launch_gsc_n(const KernelParams& p, cudaStream_t st) {
  const size_t smem = sizeof(float2) * (size_t)p.M * NN + sizeof(float) * (size_t)p.M * (NN / 2) + 16;
  cudaError_t e = cudaFuncSetAttribute(gsc_align_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  gsc_align_kernel<NN><<<p.n_streams, kGenThreads, smem, st>>>(p);
  return cudaGetLastError();
}
size_t gsc_align_smem(int N, int M) { return sizeof(float2) * (size_t)M * N + sizeof(float) * (size_t)M * (N / 2) + 16; }
cudaError_t launch_gsc(const KernelParams& p, cudaStream_t st) {
  cudaError_t e = cudaErrorNotSupported;
  switch (p.N) {
    case 512: e = launch_gsc_n<512>(p, st); break;
    case 1024: e = launch_gsc_n<1024>(p, st); break;
    case 2048: e = launch_gsc_n<2048>(p, st); break;
    case 4096: e = launch_gsc_n<4096>(p, st); break;
  }
  if (e != cudaSuccess) return e;
  const int M1 = p.M - 1;
  if (p.gsc_F == 128 && M1 >= 1 && M1 <= 3) {   // launch-file shape: filter_size 128, 2-4 microphones
    switch (M1) {
      case 1: return launch_nlms_fast<1>(p, st);
      case 2: return launch_nlms_fast<2>(p, st);
      default: return launch_nlms_fast<3>(p, st);
    }
  }
  const size_t smem = sizeof(float) * kNlmsWarps * ((size_t)(2 * M1 + 1) * p.gsc_F + (size_t)p.M * 32 + 32);
  e = cudaFuncSetAttribute(gsc_nlms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  gsc_nlms_kernel<<<(p.n_streams + kNlmsWarps - 1) / kNlmsWarps, kNlmsWarps * 32, smem, st>>>(p);
  return cudaGetLastError();
}

size_t frames_kernel_n_smem(int N, int M, int algo);
size_t frames_kernel_sel_smem(int N, int M);
template <int NN>
static size_t gen_smem(int M, bool pha) { return (pha ? sizeof(double2) : sizeof(float2)) * (size_t)M * NN + sizeof(float2) * NN + sizeof(GenScratch<NN>) + 16; }

// das: the largest number of microphones whose spectra fit the shared memory together (the others follow in further chunks)
int frames_kernel_n_das_chunk(int N, int M) {
  int mc = M;
  while (mc > 1 && frames_kernel_n_smem(N, mc, ALGO_DAS) > 232448) mc--;
  return mc;
}

size_t frames_kernel_n_smem(int N, int M, int algo) {
  const bool pha = algo == ALGO_PHASE || algo == ALGO_PHASEMPF;
  switch (N) {
    case 512: return gen_smem<512>(M, pha);
    case 1024: return gen_smem<1024>(M, pha);
    case 2048: return gen_smem<2048>(M, pha);
    case 4096: return gen_smem<4096>(M, pha);
  }
  return ~(size_t)0;
}

template <int ALGO, int NN>
static cudaError_t launch_n(const KernelParams& p, cudaStream_t st) {
  const size_t smem = gen_smem<NN>(ALGO == ALGO_DAS ? p.das_chunk : p.M, ALGO == ALGO_PHASE || ALGO == ALGO_PHASEMPF);
  cudaError_t e = cudaFuncSetAttribute(frames_kernel_n<ALGO, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  frames_kernel_n<ALGO, NN><<<p.n_streams, kGenThreads, smem, st>>>(p);
  return cudaGetLastError();
}

template <int ALGO>
static cudaError_t launch_algo(const KernelParams& p, cudaStream_t st) {
  switch (p.N) {
    case 512: return launch_n<ALGO, 512>(p, st);
    case 1024: return launch_n<ALGO, 1024>(p, st);
    case 2048: return launch_n<ALGO, 2048>(p, st);
    case 4096: return launch_n<ALGO, 4096>(p, st);
  }
  return cudaErrorNotSupported;
}

cudaError_t launch_frames_kernel_n(int algo, const KernelParams& p, cudaStream_t st) {
  switch (algo) {
    case ALGO_DAS: return launch_algo<ALGO_DAS>(p, st);
    case ALGO_PHASE: return launch_algo<ALGO_PHASE>(p, st);
    case ALGO_PHASEMPF: return launch_algo<ALGO_PHASEMPF>(p, st);
  }
  return cudaErrorNotSupported;
}

}   // namespace bf
