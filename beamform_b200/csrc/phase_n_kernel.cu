// Phase-mask nodes (phase / phasempf) on FP32 spectra for small arrays (2 <= M <= 4), every frame size; sm_100a.
//
//   reference path replaced (citations /root/reference/beamform/src/): util.h:217-253,289-314 (window, framing, OLA),
//   phase.cpp:70-134, phasempf.cpp:140-302 + the output smoother phasempf.cpp:78-83,331-334.
//
// One CTA (256 threads) owns one stream and walks its frame pairs in order, like frames_kernel_n<phase*> in
// generic_kernel.cu, but the forward transforms are FP32 and the kernel is sized for THREE CTAs per SM (24 warps instead
// of 8): M spectrum tiles of N float2 (G is assembled in place in tile 0 and the out-of-place inverse ping-pongs between
// tiles 0 and 1; the smoother's window lives in whichever of the two is free afterwards), plus the OLA tail: 73 KB for
// the 4096-point two-microphone configuration C4 against 170 KB with double spectra.
//
// Mask decisions stay exact.  Every decision (phase.cpp:114, phasempf.cpp:234) is a threshold on bin phases; an output
// Z[k] of the FP32 transform carries an absolute error of at most
//     kFftErr ||z||_2 + kFftErrMax max_k |Z[k]| + kFftErrRel |Z[k]|
// (z = the packed windowed frame pair of the microphone; the middle term is the leakage of a strong spectral line into
// the outputs that share its last butterflies; constants measured with margin by tools/fft_err.cu:
// profiles/r02_fp32_fft_err.txt), which bounds the phase error of an unpacked bin X = Z[j] +- conj(Z[N-j]).  A decision
// whose margin is inside that bound is listed, the CTA re-takes the listed decisions from exact double DFTs of those
// bins (256 threads, batches that share the sample loads), and the owner threads finish those bins afterwards - the
// same construction as the magnitude gate of the mvdr/lcmv/gss kernels.  Bins below -80 dB of the frame (|X| mean
// <= 1e-4 E) are not re-decided: their contribution to the output is below FP32 resolution of the frame (the 1024-point
// kernel has the same rule; tests bound the mask mismatches it admits).
#include <cstdio>
#include <cstdlib>

#include "block_fft.cuh"

namespace bf {

constexpr float kFftErr = 1.0e-6f;    // * ||z||_2
constexpr float kFftErrMax = 1.5e-7f; // * max_k |Z[k]| (a strong line leaks into the outputs that share its last butterflies)
constexpr float kFftErrRel = 4.0e-7f; // * |Z[k]| of the output itself

constexpr int kPhnList = 160;   // exact re-decisions per round (a pair with more takes further rounds)

template <int NN, int T>
struct PhnScratch {
  static constexpr int kGenThreads = T;
  float tail[NN / 2];
  float hist[64];                       // phasempf smoother: the last smooth_size-1 OLA samples
  float epart[kGenThreads / 32][4];     // per-warp partial energies of the packed frame pair, per microphone
  float kE[4];                          // bound on the absolute error of a microphone's unpacked bins in this pair
  float e_all;                          // sum over microphones of E
  float sin_thr, cos_thr;
  double red[kGenThreads / 32][8][2];   // exact DFTs of a batch of items: per-warp partial sums, [item * MM + microphone]
  float zpart[kGenThreads / 32][4];     // per-warp max |Z[k]|_1 per microphone
  int n_list;
  unsigned short list[kPhnList];        // (bin << 1 | frame) awaiting an exact decision
  unsigned char res[kPhnList];          // its flag bits
};

// Exact decisions of the listed (bin, frame) items: double DFT of the bin from the input samples by the whole CTA, then
// phase.cpp:89-123 / phasempf.cpp:212-248 in double by one thread per item.  Items go in batches of B that share the
// sample and window loads (the three hops of the pair serve both frames).  With n = n' + H h (n' < H):
//   X[j] = sum_{n'} W^{j n'} (x[n'] w[n'] + (-1)^j x[n'+H] w[n'+H]),   W^{j n'} = W^{j tid} W^{256 j i},  n' = tid + 256 i:
// one gathered twiddle per thread and item and H/256 that are the same for every thread.
// Flag bits: bit0 = magnitude gate passed, bit1 = phases agree.
template <int NN, int MM, int T>
__device__ __noinline__ void phn_decide_exact(const KernelParams& p, PhnScratch<NN, T>& sc, int s, int t, bool two, int n_items, bool use_gate) {
  constexpr int kGenThreads = T;   // threads of the CTA (shadows the default)
  constexpr int H = NN / 2, L = NN / 2 + 2, W = kGenThreads / 32;
  constexpr int B = MM == 2 ? 4 : 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* ha[MM];
  const float* hb[MM];
#pragma unroll
  for (int ch = 0; ch < MM; ch++) {
    const float* base = p.in + (size_t)s * p.in_stream_stride + (size_t)ch * p.in_mic_stride;
    ha[ch] = (t - 1 < 0) ? p.prev_hop + ((size_t)s * MM + ch) * H : base + (size_t)(t - 1) * H;
    hb[ch] = base + (size_t)t * H;   // hop t; hop t+1 follows it
  }
#pragma unroll 1
  for (int q0 = 0; q0 < n_items; q0 += B) {
    int jj[B], ff[B];
    double2 wb[B];
    double re[B][MM], im[B][MM];
#pragma unroll
    for (int e = 0; e < B; e++) {
      const int q = min(q0 + e, n_items - 1);
      const int l = sc.list[q] >> 1;
      ff[e] = sc.list[q] & 1;
      jj[e] = (l == L - 1) ? H + 1 : l;
      wb[e] = p.twid_d[(jj[e] * tid) & (NN - 1)];
#pragma unroll
      for (int ch = 0; ch < MM; ch++) re[e][ch] = im[e][ch] = 0.0;
    }
#pragma unroll 1
    for (int i = 0; i < H / kGenThreads; i++) {
      const int n = tid + i * kGenThreads;
      const double w0 = p.win_d[n], w1 = p.win_d[n + H];
      double xa[MM], xb0[MM], xb1[MM], xc[MM];
#pragma unroll
      for (int ch = 0; ch < MM; ch++) {
        const double b = (double)hb[ch][n];
        xa[ch] = (double)ha[ch][n] * w0;
        xb1[ch] = b * w1;   // second half of frame t
        xb0[ch] = b * w0;   // first half of frame t+1
        xc[ch] = two ? (double)hb[ch][n + H] * w1 : 0.0;
      }
#pragma unroll
      for (int e = 0; e < B; e++) {
        const double2 u = p.twid_d[(jj[e] * kGenThreads * i) & (NN - 1)];
        const double2 w = make_double2(fma(-wb[e].y, u.y, wb[e].x * u.x), fma(wb[e].y, u.x, wb[e].x * u.y));
        const double sg = (jj[e] & 1) ? -1.0 : 1.0;
#pragma unroll
        for (int ch = 0; ch < MM; ch++) {
          const double v = ff[e] ? fma(sg, xc[ch], xb0[ch]) : fma(sg, xb1[ch], xa[ch]);
          re[e][ch] = fma(v, w.x, re[e][ch]);
          im[e][ch] = fma(v, w.y, im[e][ch]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < B; e++)
#pragma unroll
      for (int ch = 0; ch < MM; ch++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          re[e][ch] += __shfl_xor_sync(0xffffffffu, re[e][ch], o);
          im[e][ch] += __shfl_xor_sync(0xffffffffu, im[e][ch], o);
        }
        if (lane == 0) { sc.red[warp][e * MM + ch][0] = re[e][ch]; sc.red[warp][e * MM + ch][1] = im[e][ch]; }
      }
    __syncthreads();
    if (lane == 0 && warp < B && q0 + warp < n_items) {   // warp e decides item e of the batch
      const int l = sc.list[q0 + warp] >> 1;
      double phi[MM];
      double magsum = 0.0;
#pragma unroll
      for (int ch = 0; ch < MM; ch++) {
        double xr = 0.0, xi = 0.0;
#pragma unroll
        for (int w = 0; w < W; w++) { xr += sc.red[w][warp * MM + ch][0]; xi += sc.red[w][warp * MM + ch][1]; }
        magsum += hypot(xr, xi);
        const double2 w = p.steer_d[(size_t)l * MM + ch];   // weights(i,j); aligned = conj(w) * X
        phi[ch] = atan2(xi * w.x - xr * w.y, xr * w.x + xi * w.y);
      }
      unsigned fl = 0;
      if (!use_gate || (magsum / MM) / (double)NN > p.mag_threshold_d) fl |= 1;
      double tot = 0.0;   // phase.cpp:53-68: sum over all pairs in the association order of the recursion
#pragma unroll
      for (int a = MM - 2; a >= 0; a--) {
        double lvl = 0.0;
#pragma unroll
        for (int b = a + 1; b < MM; b++) {
          double d = fabs(phi[a] - phi[b]);
          if (d > 3.14159265358979323846) d = 2 * 3.14159265358979323846 - d;
          lvl += d;
        }
        tot = lvl + tot;
      }
      if (tot / (double)(MM * (MM - 1) / 2) < p.min_phase_rad_d) fl |= 2;
      sc.res[q0 + warp] = (unsigned char)fl;
    }
    __syncthreads();
  }
}

// atan2 in ~22 instructions (atan2f: 53 per call, 19 % of the kernel's samples with three microphones): odd minimax polynomial of
// degree 15 on [0, 1] (max error 1.3e-7 evaluated in FP32) + octant folding; with the approximate division the absolute error stays
// below 4e-7 rad, which the decision's guard band carries (3e-6 below).  NaN in, NaN out, like atan2f.
__device__ __forceinline__ float phn_atan2(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float a = mx > 0.f ? __fdividef(mn, mx) : 0.f;   // atan2(0, 0) = 0
  const float s = a * a;
  float q = -0.004054515156894922f;
  q = fmaf(q, s, 0.021862763911485672f);
  q = fmaf(q, s, -0.055912040174007416f);
  q = fmaf(q, s, 0.0964217558503151f);
  q = fmaf(q, s, -0.13908620178699493f);
  q = fmaf(q, s, 0.19946563243865967f);
  q = fmaf(q, s, -0.33329859375953674f);
  q = fmaf(q, s, 0.9999993443489075f);
  float r = q * a;
  r = ay > ax ? 1.57079632679489662f - r : r;
  r = x < 0.f ? 3.14159265358979324f - r : r;
  const float t = x + y;
  r = (t != t) ? t : r;
  return copysignf(r, y);
}

__device__ __forceinline__ float phn_wrap_diff(float a, float b) {   // phase.cpp:58-60
  const float d = fabsf(a - b);
  return d > 3.14159265358979f ? 6.28318530717959f - d : d;
}

template <int ALGO, int NN, int MM, int T, int CTAS>
__global__ void __launch_bounds__(T, CTAS) phase_n_kernel(const __grid_constant__ KernelParams p) {
  constexpr int kGenThreads = T;   // threads of the CTA (shadows the default)
  constexpr int H = NN / 2, L = NN / 2 + 2;
  constexpr bool kGate = (ALGO == ALGO_PHASE);
  constexpr bool kMpf = (ALGO == ALGO_PHASEMPF);
  constexpr int kIter = (H + kGenThreads - 1) / kGenThreads;
  static_assert(H % kGenThreads == 0, "every thread owns the same number of bins");
  float2* zall = reinterpret_cast<float2*>(gen_smem_raw);   // [MM][NN]
  PhnScratch<NN, T>& sc = *reinterpret_cast<PhnScratch<NN, T>*>(zall + (size_t)MM * NN);
  const unsigned t1_off = (unsigned)(NN * sizeof(float2));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.x + p.stream_begin;
  const float2* tw = p.twid_f;
  const float* win = p.win_f;
  int cur_L = p.mcra_cur_L0, first_L = p.mcra_first0;
  if (tid == 0) {
    double sd, cd;
    sincos(p.min_phase_rad_d, &sd, &cd);
    sc.sin_thr = (float)sd; sc.cos_thr = (float)cd;
  }
  const bool fast2 = (MM == 2) && p.min_phase_rad > 0.f && p.min_phase_rad < 3.1415925f;
  const bool recheck_on = p.win_d != nullptr && p.debug != 2;
  int n_recheck = 0;
  const float kappa = p.debug >= 10 ? (float)p.debug * 1.0e-7f : kFftErr;   // BF_DEBUG >= 10: error-bound experiments
  float* stg = kMpf ? p.mpf_state + (size_t)s * 7 * L : nullptr;

  for (int i = tid; i < H; i += kGenThreads) sc.tail[i] = p.tail[(size_t)s * H + i];
  if (kMpf)
    for (int i = tid; i < p.smooth_size - 1; i += kGenThreads) sc.hist[i] = p.smooth_hist[(size_t)s * 64 + i];
  __syncthreads();

  const int nh = p.hop_end - p.hop_begin;
  const int npairs = (nh + 1) >> 1;
  const float* in_s = p.in + (size_t)s * p.in_stream_stride;
#pragma unroll 1
  for (int ip = 0; ip < npairs; ip++) {
    const int t = p.hop_begin + 2 * ip;
    const bool two = t + 1 < p.hop_end;
    const int nf = two ? 2 : 1;
    // the next pair's two new hops start their trip from HBM to L2 now, one 128-byte line per request
    if (t + 2 < p.hop_end) {
      const int lines_per_mic = ((t + 3 < p.hop_end) ? 2 : 1) * (H / 32);
      for (int i = tid; i < MM * lines_per_mic; i += kGenThreads) {
        const int ch = i / lines_per_mic, ln = i - ch * lines_per_mic;
        const float* a = in_s + (size_t)ch * p.in_mic_stride + (size_t)(t + 2) * H + ln * 32;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
      }
    }
    // ---- window + pack: z = 0.5*w*(frame_t + i*frame_{t+1}), frame_t = [hop t-1 | hop t] (util.h:217-242) ----
    {
      float esum[MM];
#pragma unroll
      for (int i = 0; i < MM; i++) esum[i] = 0.f;
#pragma unroll 2
      for (int k = 0; k < kIter; k++) {
        const int n = tid + k * kGenThreads;
        const float w0 = 0.5f * __ldg(win + n), w1 = 0.5f * __ldg(win + n + H);
        float fa[MM], fb[MM], fc[MM];
#pragma unroll
        for (int i = 0; i < MM; i++) {
          const float* base = in_s + (size_t)i * p.in_mic_stride;
          const float* ha = (t - 1 < 0) ? p.prev_hop + ((size_t)s * MM + i) * H : base + (size_t)(t - 1) * H;
          fa[i] = __ldg(ha + n);
          fb[i] = __ldg(base + (size_t)t * H + n);
          fc[i] = two ? __ldg(base + (size_t)(t + 1) * H + n) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < MM; i++) {
          const float2 v0 = make_float2(fa[i] * w0, (two ? fb[i] : 0.f) * w0);
          const float2 v1 = make_float2(fb[i] * w1, fc[i] * w1);
          zall[(size_t)i * NN + swz(n)] = v0;
          zall[(size_t)i * NN + swz(n + H)] = v1;
          esum[i] = fmaf(v0.x, v0.x, fmaf(v0.y, v0.y, fmaf(v1.x, v1.x, fmaf(v1.y, v1.y, esum[i]))));
        }
      }
#pragma unroll
      for (int i = 0; i < MM; i++) {
        float e = esum[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        if (lane == 0) sc.epart[warp][i] = e;
      }
    }
    __syncthreads();
    if (tid == 0) sc.n_list = 0;   // published by the transform's own barriers
    block_fft_fn<NN, -1, float2, T>(0u, MM, tw, tid);
    {   // error bound of this pair's transforms: needs ||z||_2 (pack) and max |Z| (here)
      float zm[MM];
#pragma unroll
      for (int i = 0; i < MM; i++) zm[i] = 0.f;
#pragma unroll 2
      for (int k = 0; k < NN / kGenThreads; k++) {
#pragma unroll
        for (int i = 0; i < MM; i++) {
          const float2 a = zall[(size_t)i * NN + swz(tid + k * kGenThreads)];
          zm[i] = fmaxf(zm[i], fabsf(a.x) + fabsf(a.y));
        }
      }
#pragma unroll
      for (int i = 0; i < MM; i++) {
        float m = zm[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) sc.zpart[warp][i] = m;
      }
      __syncthreads();
      if (tid == 0) {
        float e_all = 0.f;
#pragma unroll
        for (int i = 0; i < MM; i++) {
          float e = 0.f, m = 0.f;
#pragma unroll
          for (int w = 0; w < kGenThreads / 32; w++) { e += sc.epart[w][i]; m = fmaxf(m, sc.zpart[w][i]); }
          const float E = 2.0f * sqrtf(e);
          sc.kE[i] = kappa * E + 2.0f * kFftErrMax * m;   // both transform outputs a bin is unpacked from
          e_all += E;
        }
        sc.e_all = e_all;
      }
      __syncthreads();
    }

    // ---- per-bin stage: decisions of both frames, MCRA / post-filter recursion, Hermitian assembly of G into tile 0 ----
    // MCRA window bookkeeping (phasempf.cpp:162-176) is global per frame: resolve both frames up front
    bool reset_0 = false, reset_1 = false;
    float inv_cl_0 = 1.f, inv_cl_1 = 1.f;
    int fst_0 = first_L, fst_1 = first_L;
    if (kMpf) {
      reset_0 = cur_L > p.mcra_L;
      if (reset_0) { cur_L = 1; first_L = 0; } else { cur_L++; }
      inv_cl_0 = 1.0f / (float)cur_L; fst_0 = first_L;
      if (two) {
        reset_1 = cur_L > p.mcra_L;
        if (reset_1) { cur_L = 1; first_L = 0; } else { cur_L++; }
        inv_cl_1 = 1.0f / (float)cur_L; fst_1 = first_L;
      }
    }
    if (tid == 0) {
      float2 g0 = make_float2(0.f, 0.f);
      if (ALGO == ALGO_PHASE) {   // Y[0] = X_0[0] (phase.cpp:87; real for real input); phasempf leaves bin 0 at 0 (SURVEY B-5)
        const float2 a = zall[swz(0)];
        g0 = make_float2(2.0f * a.x, two ? 2.0f * a.y : 0.f);
      }
      zall[swz(0)] = g0;
      if (p.capture)
        for (int f = 0; f < nf; f++) p.capture[(size_t)s * p.capture_stream_stride + (size_t)(t + f) * NN] = 0;
    }
    // Thread tid owns bins tid + 256 k.  Bin 0 has no decision, so its thread takes the Nyquist bin in that slot; the
    // pseudo-bin N/2+1 (SURVEY B-4) is an extra leading slot of the thread that owns bin N/2-1 and is folded into that bin:
    // Yh = (Y[N/2-1] + conj(Y[N/2+1])) / 2.  Only the owner of (j, N-j) touches those two cells of every tile, so G can
    // replace Z_0 in place.  A bin with a decision inside the error bound is left pending (its spectra stay in place) and
    // listed; the CTA then takes the listed decisions exactly (phn_decide_exact) and the owners finish their pending bins.
    float2 py0 = make_float2(0.f, 0.f), py1 = make_float2(0.f, 0.f);
    unsigned pending = (1u << (kIter + 1)) - 1u;   // bit k+1: slot k
    if (tid != kGenThreads - 1) pending &= ~1u;
    bool apply = false;
#pragma unroll 1
    while (true) {
#pragma unroll 1
      for (int k = -1; k < kIter; k++) {
        if (!((pending >> (k + 1)) & 1u)) continue;
        int l = (k < 0) ? L - 1 : tid + k * kGenThreads;
        if (l == 0) l = H;
        if (l == H - 1 && (pending & 1u)) continue;   // waits for its pseudo-bin
        const int j = (l == L - 1) ? H - 1 : l;
        float2 x[2][MM];
        float2 wst[MM];
        float dx[MM];
#pragma unroll
        for (int i = 0; i < MM; i++) {
          wst[i] = __ldg(p.steer + (size_t)l * p.C * MM + i);
          const float2 a = zall[(size_t)i * NN + swz(j)], b = zall[(size_t)i * NN + swz((NN - j) & (NN - 1))];
          x[0][i] = make_float2(a.x + b.x, a.y - b.y);   // Z[j] + conj(Z[N-j])
          x[1][i] = make_float2(a.y + b.y, b.x - a.x);   // -i (Z[j] - conj(Z[N-j]))
          if (l == L - 1) { x[0][i].y = -x[0][i].y; x[1][i].y = -x[1][i].y; }
          dx[i] = fmaf(kFftErrRel, fabsf(a.x) + fabsf(a.y) + fabsf(b.x) + fabsf(b.y), sc.kE[i]);   // |error| of both unpacked bins
        }
        float magsum[2];
        unsigned fl[2];
        bool doubt[2];
#pragma unroll
        for (int f = 0; f < 2; f++) {
          float ms = 0.f, guard = 0.f, esum_dx = 0.f;
          float2 zr[MM];
#pragma unroll
          for (int i = 0; i < MM; i++) {
            const float2 xi = x[f][i];
            const float n2 = fmaf(xi.x, xi.x, xi.y * xi.y);
            ms += sqrt_fast(n2);
            esum_dx += dx[i];
            guard += fminf(3.2f, fmaf(dx[i], rsqrtf(n2), 2.0e-6f));   // phase error bound of this microphone (n2 = 0 -> 3.2)
            zr[i] = make_float2(xi.x * wst[i].x + xi.y * wst[i].y, xi.y * wst[i].x - xi.x * wst[i].y);   // conj(w) x
          }
          unsigned b = 0;
          bool d = false;
          if (fast2) {
            // two microphones: |phi_0 - phi_1| wrapped = |arg(z_0 conj z_1)|, so "< thr" is the sign of
            // q = sin(thr) Re(u) - cos(thr) |Im(u)| = |u| sin(thr - |dphi|): no arctangent
            const float dd = zr[0].x * zr[1].x + zr[0].y * zr[1].y;
            const float cc = zr[0].y * zr[1].x - zr[0].x * zr[1].y;
            const float q = sc.sin_thr * dd - sc.cos_thr * fabsf(cc);
            if (q > 0.f) b |= 2;
            const float mod = sqrt_fast(fmaf(dd, dd, cc * cc));
            d = !(fabsf(q) > (guard + 4.0e-6f) * mod);
          } else {
            float phi[MM];
#pragma unroll
            for (int i = 0; i < MM; i++) phi[i] = phn_atan2(zr[i].y, zr[i].x);
            float tot = 0.f;
#pragma unroll
            for (int a = MM - 2; a >= 0; a--) {
              float lvl = 0.f;
#pragma unroll
              for (int c = a + 1; c < MM; c++) lvl += phn_wrap_diff(phi[a], phi[c]);
              tot = lvl + tot;
            }
            const float mean_diff = tot / (float)(MM * (MM - 1) / 2);
            if (mean_diff < p.min_phase_rad) b |= 2;
            d = !(fabsf(mean_diff - p.min_phase_rad) > guard * (2.0f / (float)MM) + 3.0e-6f);
          }
          const float e_all = sc.e_all;
          d = d && ms > 1.0e-4f * e_all;   // only bins that matter
          if (kGate) {
            const float thr = p.thr_phase_mag;
            if (ms > thr) b |= 1;
            if (fabsf(ms - thr) <= 2.0f * esum_dx + 1.0e-6f * thr) d = true;
          } else {
            b |= 1;
          }
          magsum[f] = ms; fl[f] = b; doubt[f] = d && f < nf && recheck_on;
        }
        if (doubt[0] || doubt[1]) {
          if (apply) {   // exact decisions of this round: take the ones that are there
            const int n_items = min(sc.n_list, kPhnList);
#pragma unroll
            for (int f = 0; f < 2; f++) {
              if (!doubt[f]) continue;
              const unsigned short key = (unsigned short)(l * 2 + f);
              for (int q = 0; q < n_items; q++)
                if (sc.list[q] == key) { fl[f] = sc.res[q]; doubt[f] = false; break; }
            }
          } else {
            // all-or-nothing: a bin that needs both frames takes two adjacent slots or none.  (Taken one by one, an overflowing
            // round hands every slot to the first frames of the first warps, no bin ever gets both, and the pair never ends.)
            const int cnt = (doubt[0] ? 1 : 0) + (doubt[1] ? 1 : 0);
            int q = atomicAdd(&sc.n_list, cnt);
            if (q + cnt <= kPhnList) {   // else: a later round
              if (doubt[0]) sc.list[q++] = (unsigned short)(l * 2);
              if (doubt[1]) sc.list[q] = (unsigned short)(l * 2 + 1);
            } else if (q < kPhnList) {
              sc.list[q] = (unsigned short)(l * 2);   // the one slot a refused pair leaves behind: a valid key, so the round decides it (unused)
            }
          }
          if (doubt[0] || doubt[1]) continue;   // stays pending
        }
        pending &= ~(1u << (k + 1));
        // ---- output of the bin (phase.cpp:114-123 / phasempf.cpp:140-191,234-302) ----
        float st7[7];
        if (kMpf) {
#pragma unroll
          for (int q = 0; q < 7; q++) st7[q] = stg[q * L + l];
        }
        float2 yy[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int f = 0; f < 2; f++) {
          if (f >= nf) break;
          const unsigned b = fl[f];
          const bool reset_f = f ? reset_1 : reset_0;
          const float inv_cl_f = f ? inv_cl_1 : inv_cl_0;
          const int fst_f = f ? fst_1 : fst_0;
          if (p.capture) {
            unsigned char* cap = p.capture + (size_t)s * p.capture_stream_stride + (size_t)(t + f) * NN;
            const unsigned char cf = (unsigned char)(((b & 1) && (b & 2)) ? 2 : 0);
            if (l <= H) {
              cap[l] = cf;
              if (l > 0 && l < H - 1) cap[NN - l] = cf;
            } else {
              cap[H + 1] = cf;
            }
          }
          const float2 x0 = x[f][0];
          const float mag_mean = magsum[f] * (1.0f / (float)MM);
          const float n0 = fmaf(x0.x, x0.x, x0.y * x0.y);
          const float r0 = rsqrtf(n0);
          const float2 unit = n0 > 0.f ? make_float2(x0.x * r0, x0.y * r0) : make_float2(1.f, 0.f);   // e^{i arg X_0}
          float2 y;
          if (ALGO == ALGO_PHASE) {
            const float mag = ((b & 1) && (b & 2)) ? mag_mean : mag_mean * p.mag_mult;
            y = make_float2(mag * unit.x, mag * unit.y);
          } else {
            const bool kept = (b & 2) != 0;
            const float soi = kept ? mag_mean : mag_mean * p.min_mag;   // phasempf.cpp:234-244
            const float itf = kept ? mag_mean * p.min_mag : mag_mean;
            const float s2 = soi * soi, i2 = itf * itf;
            const float Sf = (l == 1) ? 0.75f * s2 : s2;   // SURVEY B-9: only bins 1 and N-1 are scaled
            const float S = p.mcra_alphaS * st7[0] + (1.0f - p.mcra_alphaS) * Sf;
            const float m_min = reset_f ? st7[1] : st7[2];
            st7[2] = fminf(m_min, S);
            st7[1] = reset_f ? S : fminf(st7[1], S);
            const bool upd = fst_f || S < st7[2] * p.mcra_delta || st7[3] > s2;
            const bool avg = fst_f && inv_cl_f > p.mcra_alphaD;
            const float ca = avg ? inv_cl_f : p.mcra_alphaD2, cb = avg ? 1.0f - inv_cl_f : 1.0f - p.mcra_alphaD;   // SURVEY B-16
            st7[3] = upd ? ca * st7[3] + cb * s2 : st7[3];
            st7[0] = S;
            st7[4] = p.mpf_alphaS * st7[4] + (1.0f - p.mpf_alphaS) * i2;   // phasempf.cpp:255-271
            st7[5] = p.mpf_gamma * st7[5] + p.mpf_rev_gain * s2;
            st7[6] = p.mpf_gamma * st7[6] + p.mpf_rev_gain * i2;
            const float Lam = sqrt_fast(st7[3] + p.mpf_eta * st7[4] + st7[5] + st7[6]);
            float mag;
            if (p.out_only_noise) {
              mag = Lam * p.out_amp;
            } else {
              mag = p.out_only_mcra ? (soi - sqrt_fast(st7[3])) * p.out_amp : (soi - Lam) * p.out_amp;
              if (mag < 0.f) mag = p.noise_floor;
            }
            const float2 u2 = soi > 0.f ? unit : make_float2(1.f, 0.f);
            y = make_float2(mag * u2.x, mag * u2.y);
          }
          yy[f] = y;
        }
        if (kMpf) {
#pragma unroll
          for (int q = 0; q < 7; q++) stg[q * L + l] = st7[q];
        }
        if (l == L - 1) {
          py0 = yy[0]; py1 = yy[1];
        } else {
          float2 a0 = yy[0], a1 = yy[1];
          if (l == H - 1) {
            a0 = make_float2(0.5f * (a0.x + py0.x), 0.5f * (a0.y - py0.y));
            a1 = make_float2(0.5f * (a1.x + py1.x), 0.5f * (a1.y - py1.y));
          }
          if (l == H) { a0.y = 0.f; a1.y = 0.f; }   // Re(): self-conjugate bin
          zall[swz(l)] = make_float2(a0.x - a1.y, a0.y + a1.x);                          // Yh_t + i Yh_{t+1}
          if (l < H) zall[swz(NN - l)] = make_float2(a0.x + a1.y, a1.x - a0.y);          // conj(Yh_t) + i conj(Yh_{t+1})
        }
      }
      if (!__syncthreads_or(pending != 0u)) break;
      if (apply) {   // everything listed has been consumed: the rest (a list overflow) is a new round
        apply = false;
        __syncthreads();
        if (tid == 0) sc.n_list = 0;
        __syncthreads();
      } else {
        const int n_items = min(sc.n_list, kPhnList);
        phn_decide_exact<NN, MM, T>(p, sc, s, t, two, n_items, kGate);
        n_recheck += n_items;
        apply = true;
      }
    }
    // ---- inverse (out of place, tiles 0 <-> 1), synthesis window, overlap-add (util.h:244-253, 301-302), smoother ----
    const unsigned r_off = block_fft_oop_fn<NN, 1, float2>(0u, t1_off, tw, tid);
    const float2* res = reinterpret_cast<const float2*>(gen_smem_raw + r_off);
    float* ola = reinterpret_cast<float*>(gen_smem_raw + (r_off == 0u ? t1_off : 0u));   // the free tile
    float* o0 = p.out + (size_t)s * p.out_stream_stride + (size_t)t * H;
    const int S1 = kMpf ? p.smooth_size - 1 : 0;
#pragma unroll 2
    for (int k = 0; k < kIter; k++) {
      const int n = tid + k * kGenThreads;
      const float w0 = __ldg(win + n) * p.out_scale, w1 = __ldg(win + n + H) * p.out_scale;
      const float2 a = res[swz(n)], b = res[swz(n + H)];
      const float r0 = sc.tail[n] + a.x * w0;
      if (kMpf) ola[S1 + n] = r0; else o0[n] = r0;
      if (two) {
        const float r1 = b.x * w1 + a.y * w0;
        if (kMpf) ola[S1 + H + n] = r1; else o0[H + n] = r1;
        sc.tail[n] = b.y * w1;
      } else {
        sc.tail[n] = b.x * w1;
      }
    }
    if (kMpf) {
      // phasempf.cpp:78-83,122-130,331-334: every output sample becomes the mean of the last smooth_size OLA samples
      if (tid < S1) ola[tid] = sc.hist[tid];
      __syncthreads();
      const int cnt = two ? NN : H, S = p.smooth_size;
      const double inv = 1.0 / (double)S;
      for (int n = tid; n < cnt; n += kGenThreads) {
        double acc = 0.0;
        for (int k = 0; k < S; k++) acc += (double)ola[n + k];
        o0[n] = (float)(acc * inv);
      }
      if (tid < S1) sc.hist[tid] = ola[cnt + tid];
    }
    __syncthreads();
  }
  if ((p.debug == 3 || (p.debug >= 1000 && p.debug % 1000 == 3)) && blockIdx.x < 4 && tid == 0) printf("phase_n_kernel: stream %d: %d exact re-decisions over %d pairs\n", s, n_recheck, npairs);
  for (int i = tid; i < H; i += kGenThreads) p.tail[(size_t)s * H + i] = sc.tail[i];
  if (kMpf)
    for (int i = tid; i < p.smooth_size - 1; i += kGenThreads) p.smooth_hist[(size_t)s * 64 + i] = sc.hist[i];
}

template <int NN, int T>
static size_t phn_smem(int M) { return sizeof(float2) * (size_t)M * NN + sizeof(PhnScratch<NN, T>) + 16; }

template <int ALGO, int NN, int MM>
static cudaError_t launch_phn(const KernelParams& p, cudaStream_t st) {
  // CTA size and CTAs per SM.  Up to 1024 points: 128 threads x 4 CTAs at 128 registers (a 1024-point transform is 64 radix-16 / 128
  // radix-8 tasks, the warp that carries the pseudo-bin delays 4 warps instead of 8, and the per-bin stage does not spill; measured on
  // the 3-microphone phase node: 256 x 3 at 80 registers 5.54 ms, 128 x 6 at 80 registers 5.33 ms, 128 x 4 at 128 registers 4.22 ms).
  // Longer frames: 256 threads x 3 CTAs (the shared memory allows no more; 80 registers).  BF_PHN_T / BF_PHN_CTAS override (tuning).
  const char* e_t = getenv("BF_PHN_T");
  const char* e_c = getenv("BF_PHN_CTAS");
  const int threads = e_t ? atoi(e_t) : (NN <= 1024 ? 128 : 256);
  const int ctas = e_c ? atoi(e_c) : 0;
  void (*k)(KernelParams) = nullptr;
  size_t smem = 0;
  int nthreads = 0;
  if constexpr (NN <= 1024) {
    if (threads == 128) {
      k = ctas == 6 ? phase_n_kernel<ALGO, NN, MM, 128, 6> : ctas == 3 ? phase_n_kernel<ALGO, NN, MM, 128, 3> : phase_n_kernel<ALGO, NN, MM, 128, 4>;
      smem = phn_smem<NN, 128>(MM);
      nthreads = 128;
    }
  }
  if (!k) {
    k = ctas == 2 ? phase_n_kernel<ALGO, NN, MM, 256, 2> : phase_n_kernel<ALGO, NN, MM, 256, 3>;
    smem = phn_smem<NN, 256>(MM);
    nthreads = 256;
  }
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<p.n_streams, nthreads, smem, st>>>(p);
  return cudaGetLastError();
}
template <int ALGO, int NN>
static cudaError_t launch_phn_m(const KernelParams& p, cudaStream_t st) {
  switch (p.M) {
    case 2: return launch_phn<ALGO, NN, 2>(p, st);
    case 3: return launch_phn<ALGO, NN, 3>(p, st);
    case 4: return launch_phn<ALGO, NN, 4>(p, st);
  }
  return cudaErrorNotSupported;
}
template <int ALGO>
static cudaError_t launch_phn_n(const KernelParams& p, cudaStream_t st) {
  switch (p.N) {
    case 512: return launch_phn_m<ALGO, 512>(p, st);
    case 1024: return launch_phn_m<ALGO, 1024>(p, st);
    case 2048: return launch_phn_m<ALGO, 2048>(p, st);
    case 4096: return launch_phn_m<ALGO, 4096>(p, st);
  }
  return cudaErrorNotSupported;
}

// Default dispatch: frames of up to 1024 points.  Longer frames work (tests run them with BF_PHASE_F32=1) but do not pay
// on tonal input: the leakage term of the error bound grows with the spectral peaks, the 4096-point configuration C4 on
// the bench signal then re-decides ~8 bins per frame pair (4 K-sample double DFTs each), and the double-spectra kernel
// of generic_kernel.cu is faster (profiles/r02_experiments.md).
bool phase_n_supported(const KernelParams& p, int algo) {
  const bool all_sizes = getenv("BF_PHASE_F32") != nullptr;   // read per launch: tests switch it
  if (algo != ALGO_PHASE && algo != ALGO_PHASEMPF) return false;
  if (p.M < 2 || p.M > 4) return false;
  return p.N == 512 || p.N == 1024 || (all_sizes && (p.N == 2048 || p.N == 4096));
}
cudaError_t launch_phase_n(int algo, const KernelParams& p, cudaStream_t st) {
  if (algo == ALGO_PHASE) return launch_phn_n<ALGO_PHASE>(p, st);
  if (algo == ALGO_PHASEMPF) return launch_phn_n<ALGO_PHASEMPF>(p, st);
  return cudaErrorNotSupported;
}

}   // namespace bf
