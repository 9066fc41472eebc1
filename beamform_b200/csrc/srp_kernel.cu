// Steered-response power sweep (BASELINE config C5): the reference DAS response (das.cpp:41,61-62)
//   y_d[j] = (1/M) sum_i conj(w_{d,i}[j]) X_i[j],   w_{d,i}[j] = exp(-i 2 pi freqs[j] tau_{d,i})   (util.h:136-161)
// evaluated for D look directions per frame; map[s][t][d] = sum_{j=0}^{N-1} |y_d[j]|^2.
//
// Two kernels, 1024-point frames, M <= 64:
//   srp_spectra_kernel     window -> packed FFT of every microphone of a frame pair (same warp-private transform as the
//                          beamforming kernels) -> half spectra + pseudo-bin, split into bf16 hi/lo and written as the
//                          ready-made A-operand images of the tensor-core kernel
//   srp_power_tc_kernel    (srp_tc_kernel.cu) per bin l a complex GEMM Y_l[f][d] = sum_i X_l[f][i] A_l[d][i] on tcgen05
//                          (BF16x3 split, TMEM accumulators), |Y|^2 weighted by the bin's multiplicity (mirror bins share
//                          |y|, SURVEY B-3/B-4 pair excepted) and accumulated over l in registers.
#include <cstdlib>

#include <cuda_bf16.h>

#include "bf_device.h"
#include "fft_reg.cuh"
#include "warp_fft1024.cuh"

namespace bf {

constexpr int kSrpL = 514;   // logical bins of a 1024-point frame
constexpr int kSrpTilePitch = 1026;   // float2 per warp tile in srp_spectra_kernel


// grid = (pairs, streams); frame index f = s*n_hops + t
__global__ void __launch_bounds__(256, 1) srp_spectra_kernel(const KernelParams p, unsigned char* __restrict__ xi, int n_hops, size_t image_bytes, size_t lbo) {
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
  const size_t n_ft = (F + 127) / 128;
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
      warp_fft1024_fwd(v, tile, tw, lane);
#pragma unroll
      for (int k2 = 0; k2 < 32; k2++) tile[k2 * 32 + lane] = v[k2];
    }
    __syncthreads();
    // Output: the A-operand IMAGES of the tensor-core kernel (srp_tc_kernel.cu): per (bin, tile of 128 frames) a bf16 hi
    // tile and a bf16 lo tile in the UMMA core-matrix layout, element (frame row, k) with k = microphone (real parts) and
    // 64 + microphone (imaginary parts).  The eight microphones of a round are one 16-byte core-matrix row, the two frames
    // of the pair are neighbouring rows: each thread (one bin) writes full 16-byte pieces, frame pairs fill 32-byte sectors.
    // A lane PAIR per bin: lane parity = frame of the pair, so the two 16-byte pieces of a store instruction's neighbouring
    // lanes are the two neighbouring rows of one 32-byte sector.
    for (int it = tid; it < 2 * kSrpL; it += 256) {
      const int l = it >> 1, f = it & 1;
      const int j = (l == kSrpL - 1) ? 511 : l;
      __align__(16) __nv_bfloat16 h[2][8];    // [re/im][microphone of the round]
      __align__(16) __nv_bfloat16 lo[2][8];
#pragma unroll
      for (int m = 0; m < 8; m++) {
        float xr = 0.f, xim = 0.f;
        if (r0 + m < M) {
          const float2* zt = tiles + (size_t)m * kSrpTilePitch;
          const float2 a = zt[j], b = zt[(1024 - j) & 1023];
          xr = f ? a.y + b.y : a.x + b.x;
          xim = f ? b.x - a.x : a.y - b.y;
          if (l == kSrpL - 1) xim = -xim;   // pseudo-bin: conj of bin N/2-1 (SURVEY B-4)
        }
        h[0][m] = __float2bfloat16_rn(xr);
        lo[0][m] = __float2bfloat16_rn(xr - __bfloat162float(h[0][m]));
        h[1][m] = __float2bfloat16_rn(xim);
        lo[1][m] = __float2bfloat16_rn(xim - __bfloat162float(h[1][m]));
      }
      if (f == 0 || two) {
        const size_t fr = f0 + f;
        unsigned char* img = xi + ((size_t)l * n_ft + (fr >> 7)) * image_bytes;
        const int row = (int)(fr & 127);
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const size_t off = (size_t)((c * 64 + r0) >> 3) * lbo + (size_t)(row >> 3) * 128 + (size_t)(row & 7) * 16;
          *reinterpret_cast<uint4*>(img + off) = *reinterpret_cast<const uint4*>(h[c]);
          *reinterpret_cast<uint4*>(img + image_bytes / 2 + off) = *reinterpret_cast<const uint4*>(lo[c]);
        }
      }
    }
    __syncthreads();
  }
}

cudaError_t launch_srp_power_tc(const unsigned char* xi, const double* tau, const double* freqs_l, float* maps, int D, int M, long long F,
                                cudaStream_t st);
size_t srp_image_bytes();

// bytes of the operand-image workspace for F frames: [514 bins][ceil(F/128) frame tiles][image]
size_t srp_workspace_bytes(long long F) { return (size_t)kSrpL * (size_t)((F + 127) / 128) * srp_image_bytes(); }

cudaError_t launch_srp(const KernelParams& p, unsigned char* xi, const double* tau, const double* freqs_l, float* maps, int D, int n_hops,
                       cudaStream_t st) {
  const size_t smem1 = sizeof(float2) * (1024 + 8 * kSrpTilePitch);
  cudaError_t e = cudaFuncSetAttribute(srp_spectra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
  if (e != cudaSuccess) return e;
  dim3 g1((n_hops + 1) / 2, p.n_streams);
  const size_t image = srp_image_bytes(), lbo = image / 2 / 16;   // 16 K-cores per tile
  srp_spectra_kernel<<<g1, 256, smem1, st>>>(p, xi, n_hops, image, lbo);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const long long F = (long long)p.n_streams * n_hops;
  return launch_srp_power_tc(xi, tau, freqs_l, maps, D, p.M, F, st);
}

}   // namespace bf
