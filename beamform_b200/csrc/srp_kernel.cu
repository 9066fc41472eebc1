// Steered-response power sweep (BASELINE config C5): the reference DAS response (das.cpp:41,61-62)
//   y_d[j] = (1/M) sum_i conj(w_{d,i}[j]) X_i[j],   w_{d,i}[j] = exp(-i 2 pi freqs[j] tau_{d,i})   (util.h:136-161)
// evaluated for D look directions per frame; map[s][t][d] = sum_{j=0}^{N-1} |y_d[j]|^2.
//
// Two kernels, 1024-point frames, M <= 64:
//   srp_spectra_kernel  window -> packed FFT of every microphone of a frame pair (same warp-private transform as the
//                       beamforming kernels) -> half spectra + pseudo-bin, written bin-major XS[l][frame][mic]
//   srp_power_kernel    per bin l a complex GEMM  Y_l[d][f] = A_l[d][i] X_l[i][f]  (A_l generated on the fly from the
//                       delay table, exact phase reduction in double), |Y|^2 weighted by the bin's multiplicity
//                       (mirror bins share |y|, SURVEY B-3/B-4 pair excepted) and accumulated over l in registers.
// The product path runs the contraction on tcgen05 (srp_tc_kernel.cu, BF16x3 split, TMEM accumulators); the FP32-pipe
// version below (64x64x64 shared-memory tiles, 4x4 register tiles) is kept as an A/B cross-check (env BF_SRP_FP32).
#include <cstdlib>

#include "bf_device.h"
#include "fft_reg.cuh"
#include "warp_fft1024.cuh"

namespace bf {

constexpr int kSrpL = 514;   // logical bins of a 1024-point frame
constexpr int kSrpTilePitch = 1026;   // float2 per warp tile in srp_spectra_kernel

__device__ __forceinline__ void srp_fft1024_fwd(float2 (&v)[32], float2* tile, const float2* __restrict__ tw, int lane) {
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    fft_dit<32, -1>(v);
    if (pass == 0) {
#pragma unroll
      for (int k1 = 1; k1 < 32; k1++) v[k1] = cmul(v[k1], tw[k1 * 32 + lane]);
#pragma unroll
      for (int k1 = 0; k1 < 32; k1++) tile[k1 * 32 + (lane ^ ((k1 & 15) << 1))] = v[k1];
      __syncwarp();
      const float4* row = reinterpret_cast<const float4*>(tile + lane * 32);
      const int sw = lane & 15;
      static_for<0, 16>([&](auto q) {
        const float4 r = row[q ^ sw];
        v[brev5(2 * q)] = make_float2(r.x, r.y);
        v[brev5(2 * q + 1)] = make_float2(r.z, r.w);
      });
      __syncwarp();
    }
  }
}

// grid = (pairs, streams); XS[l][f][i], f = s*n_hops + t
__global__ void __launch_bounds__(256, 1) srp_spectra_kernel(const KernelParams p, float2* __restrict__ xs, int n_hops) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* tw = reinterpret_cast<float2*>(smem_raw);
  float2* tiles = tw + 1024;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = p.M;
  constexpr int H = 512;
  const int s = blockIdx.y, t = 2 * blockIdx.x;
  const bool two = t + 1 < n_hops;
  for (int i = tid; i < 1024; i += blockDim.x) {
    const int k1 = i >> 5, l = i & 31;
    float sn, cs;
    sincospif(-2.0f * (float)((k1 * l) & 1023) / 1024.0f, &sn, &cs);
    tw[i] = make_float2(cs, sn);
  }
  __syncthreads();
  double sd, cd;
  sincospi((double)lane / 1024.0, &sd, &cd);
  const float s_l = (float)(0.5 * sd), c_l = (float)(0.5 * cd);
  // tile pitch 1026 float2: 16-byte aligned, and the eight tiles of a round sit two 8-byte banks apart, so the
  // transposing read below (eight lanes = eight microphones at the same bin) is conflict-free
  float2* tile = tiles + (size_t)warp * kSrpTilePitch;
  const size_t F = (size_t)p.n_streams * n_hops;
  const size_t f0 = (size_t)s * n_hops + t;
  for (int r0 = 0; r0 < M; r0 += 8) {
    const int ch = r0 + warp;
    if (ch < M) {
      const float* base = p.in + (size_t)s * p.in_stream_stride + (size_t)ch * p.in_mic_stride;
      const float* ha = (t - 1 < 0) ? p.prev_hop + ((size_t)s * M + ch) * H : base + (size_t)(t - 1) * H;
      const float* hb = base + (size_t)t * H;
      const float* hc = two ? base + (size_t)(t + 1) * H : hb;
      float2 v[32];
      static_for<0, 16>([&](auto r) {
        const float a = __ldg(ha + 32 * r + lane), bb = __ldg(hb + 32 * r + lane);
        const float c = two ? __ldg(hc + 32 * r + lane) : 0.0f;
        const float w0 = win1024<r>(s_l, c_l);
        const float w1 = win1024<r + 16>(s_l, c_l);
        v[brev5(r)] = make_float2(a * w0, bb * w0);
        v[brev5(r + 16)] = make_float2(bb * w1, c * w1);
      });
      srp_fft1024_fwd(v, tile, tw, lane);
#pragma unroll
      for (int k2 = 0; k2 < 32; k2++) tile[k2 * 32 + lane] = v[k2];
    }
    __syncthreads();
    // The eight spectra of the round leave together: lanes 8q..8q+7 carry the eight microphones of one bin, so every
    // (bin, frame) gets one 64-byte run of XS[l][f][r0..r0+7] (full sectors) instead of eight scattered 8-byte stores.
    {
      const int m = tid & 7;
      if (r0 + m < M) {
        const float2* zt = tiles + (size_t)m * kSrpTilePitch;
        for (int l = tid >> 3; l < kSrpL; l += 32) {
          const int j = (l == kSrpL - 1) ? 511 : l;
          const float2 a = zt[j], b = zt[(1024 - j) & 1023];
          float2 x0 = make_float2(a.x + b.x, a.y - b.y);
          float2 x1 = make_float2(a.y + b.y, b.x - a.x);
          if (l == kSrpL - 1) { x0.y = -x0.y; x1.y = -x1.y; }   // pseudo-bin: conj of bin N/2-1 (SURVEY B-4)
          xs[((size_t)l * F + f0) * M + r0 + m] = x0;
          if (two) xs[((size_t)l * F + f0 + 1) * M + r0 + m] = x1;
        }
      }
    }
    __syncthreads();
  }
}

constexpr int kTD = 64, kTF = 64;   // CTA tile: directions x frames
constexpr int kPad = 65;            // float2 row pitch of the shared tiles (odd: conflict-free column reads)

// grid = (ceil(F/64), ceil(D/64)); maps[f][d]
__global__ void __launch_bounds__(256, 2) srp_power_kernel(const float2* __restrict__ xs, const double* __restrict__ tau /*[D][M]*/,
                                                            const double* __restrict__ freqs_l /*[514]*/, float* __restrict__ maps, int D, int M,
                                                            long long F) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* As = reinterpret_cast<float2*>(smem_raw);   // [64 i][65]: As[i][d]
  float2* Xs = As + 64 * kPad;                         // [64 i][65]: Xs[i][f]
  const int tid = threadIdx.x;
  const int td = tid & 15, tf = tid >> 4;              // 16 x 16 threads, each a 4 (d) x 4 (f) register tile
  const long long fbase = (long long)blockIdx.x * kTF;
  const int dbase = blockIdx.y * kTD;
  float pw[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) pw[a][b] = 0.f;
  const float invM = 1.0f / (float)M;

  for (int l = 0; l < kSrpL; l++) {
    const double fl = freqs_l[l];
    float2 acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[a][b] = make_float2(0.f, 0.f);
    for (int i0 = 0; i0 < M; i0 += 64) {
      __syncthreads();
      // A_l[d][i] = exp(+i 2 pi f_l tau_{d,i}) / M  (conj of das.cpp:41), phase reduced to one turn in double
      for (int e = tid; e < 64 * 64; e += 256) {
        const int d = e >> 6, i = e & 63;
        float2 val = make_float2(0.f, 0.f);
        if (dbase + d < D && i0 + i < M) {
          const double turns = fl * tau[(size_t)(dbase + d) * M + i0 + i];
          const float fr = (float)(turns - rint(turns));
          float sn, cs;
          sincospif(2.0f * fr, &sn, &cs);
          val = make_float2(cs * invM, sn * invM);
        }
        As[i * kPad + d] = val;
      }
      for (int e = tid; e < 64 * 64; e += 256) {
        const int f = e >> 6, i = e & 63;
        float2 val = make_float2(0.f, 0.f);
        if (fbase + f < F && i0 + i < M) val = xs[((size_t)l * F + fbase + f) * M + i0 + i];
        Xs[i * kPad + f] = val;
      }
      __syncthreads();
#pragma unroll 4
      for (int i = 0; i < 64; i++) {
        float2 av[4], xv[4];
#pragma unroll
        for (int a = 0; a < 4; a++) av[a] = As[i * kPad + td + 16 * a];
#pragma unroll
        for (int b = 0; b < 4; b++) xv[b] = Xs[i * kPad + tf + 16 * b];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 4; b++) {
            acc[a][b].x = fmaf(av[a].x, xv[b].x, acc[a][b].x); acc[a][b].x = fmaf(-av[a].y, xv[b].y, acc[a][b].x);
            acc[a][b].y = fmaf(av[a].x, xv[b].y, acc[a][b].y); acc[a][b].y = fmaf(av[a].y, xv[b].x, acc[a][b].y);
          }
      }
    }
    // bins 1..N/2-2 stand for themselves and their mirrors (|y[N-j]| = |y[j]|); 0, N/2-1, N/2 and the pseudo-bin count once
    const float wgt = (l == 0 || l >= 511) ? 1.0f : 2.0f;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) pw[a][b] = fmaf(wgt, fmaf(acc[a][b].x, acc[a][b].x, acc[a][b].y * acc[a][b].y), pw[a][b]);
  }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int d = dbase + td + 16 * a;
      const long long f = fbase + tf + 16 * b;
      if (d < D && f < F) maps[(size_t)f * D + d] = pw[a][b];
    }
}

cudaError_t launch_srp_power_tc(const float2* xs, const double* tau, const double* freqs_l, float* maps, int D, int M, long long F,
                                cudaStream_t st);

cudaError_t launch_srp(const KernelParams& p, float2* xs, const double* tau, const double* freqs_l, float* maps, int D, int n_hops,
                       cudaStream_t st) {
  const size_t smem1 = sizeof(float2) * (1024 + 8 * kSrpTilePitch);
  cudaError_t e = cudaFuncSetAttribute(srp_spectra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
  if (e != cudaSuccess) return e;
  dim3 g1((n_hops + 1) / 2, p.n_streams);
  srp_spectra_kernel<<<g1, 256, smem1, st>>>(p, xs, n_hops);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const long long F = (long long)p.n_streams * n_hops;
  if (!getenv("BF_SRP_FP32")) return launch_srp_power_tc(xs, tau, freqs_l, maps, D, p.M, F, st);   // tensor-core path (default)
  const size_t smem2 = sizeof(float2) * 2 * 64 * kPad;
  e = cudaFuncSetAttribute(srp_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  if (e != cudaSuccess) return e;
  dim3 g2((unsigned)((F + kTF - 1) / kTF), (unsigned)((D + kTD - 1) / kTD));
  srp_power_kernel<<<g2, 256, smem2, st>>>(xs, tau, freqs_l, maps, D, p.M, F);
  return cudaGetLastError();
}

}   // namespace bf
