// Warp-private 1024-point complex FFT for sm_100a: 1024 = 32 x 32, 32 points per lane.
//
//   stage 1: each lane runs a 32-point register FFT over n1 (samples 32*n1 + lane)
//   twiddle: W_1024^{lane*k1} from a [32][32] shared table (conflict-free LDS.64)
//   exchange: ONE shared-memory transpose through a warp-private [32][34] float2 tile
//             (STS.64 rows / LDS.128 columns, both bank-conflict free), __syncwarp only
//   stage 2: each lane (= k1) runs a 32-point register FFT over n2 -> X[k1 + 32*k2]
//
// No block-level barrier is involved, so every warp of a CTA transforms independently.  The two
// real frames of a frame PAIR (t, t+1) ride in the real/imaginary parts of one complex transform
// (z = w*(frame_t + i*frame_{t+1})), which halves the FFT work for real audio; see frames_kernel.
#pragma once
#include "fft_reg.cuh"

namespace bf {

constexpr int kXRow = 34;                 // float2 per exchange-tile row (272 B: 16B aligned, conflict-free)
constexpr int kXTile = 32 * kXRow;        // float2 per warp-private tile (8704 B)

__device__ __forceinline__ constexpr int brev5(int i) { return brev(i, 5); }

// compile-time loop: f(std::integral_constant<int, I>) for I in [B, E)
template <int I>
struct IC { static constexpr int value = I; constexpr operator int() const { return I; } };
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (B < E) {
    f(IC<B>{});
    static_for<B + 1, E>(f);
  }
}

// v: in = stage-1 inputs in bit-reversed slots (v[brev5(n1)] = z[32*n1 + lane]);
//    out = X[lane + 32*k2] in v[k2].  DIR=-1 forward, +1 backward (unnormalised, like FFTW).
template <int DIR>
__device__ __forceinline__ void warp_fft1024(float2 (&v)[32], float2* tile, const float2* __restrict__ tw, int lane) {
  fft_dit<32, DIR>(v);
#pragma unroll
  for (int k1 = 1; k1 < 32; k1++) {
    float2 t = tw[k1 * 32 + lane];
    v[k1] = (DIR < 0) ? cmul(v[k1], t) : cmulc(v[k1], t);
  }
#pragma unroll
  for (int k1 = 0; k1 < 32; k1++) tile[k1 * kXRow + lane] = v[k1];
  __syncwarp();
  const float4* row = reinterpret_cast<const float4*>(tile + lane * kXRow);
  static_for<0, 16>([&](auto q) {
    float4 r = row[q];
    v[brev5(2 * q)] = make_float2(r.x, r.y);
    v[brev5(2 * q + 1)] = make_float2(r.z, r.w);
  });
  __syncwarp();
  fft_dit<32, DIR>(v);
}

// The transform the 1024-point kernels of rounds 1-2 use (das_pairs, sel_pairs, sel_stream, mcra_pairs, srp_spectra): forward only,
// with an 8 KB XOR-swizzled exchange tile.  Element (row k1, col c) lives at float2 index k1*32 + (c ^ ((k1 & 15) << 1)): row stores are
// full 256-byte rows, column loads are LDS.128 whose eight lanes per phase fall into distinct 16-byte bank groups (no padding).
// The two 32-point register passes share ONE copy of the butterfly code (rolled 2-trip loop), so that a kernel's hot loop fits the
// instruction caches; the inverse reuses it through IFFT(x) = swap(FFT(swap(x))), swap = exchange of real and imaginary parts.
// `after_exchange` runs once the exchange step has read the tile: the tile is free from there on (the caller may start a TMA copy
// of the next job's hops into it).
template <class F>
__device__ __forceinline__ void warp_fft1024_fwd(float2 (&v)[32], float2* tile, const float2* __restrict__ tw, int lane, F&& after_exchange) {
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
      after_exchange();
    }
  }
}
__device__ __forceinline__ void warp_fft1024_fwd(float2 (&v)[32], float2* tile, const float2* __restrict__ tw, int lane) {
  warp_fft1024_fwd(v, tile, tw, lane, [] {});
}

// sqrt-Hann window (util.h:201-211: sqrt(0.5 - 0.5 cos(2 pi n/N)) = sin(pi n/N) for 0 <= n < N) at
// n = 32*r + lane, built from the lane's (sin, cos)(pi*lane/N) and compile-time (cos, sin)(pi*r/32):
// two full-rate FP32 instructions, no table traffic.
template <int R>
__device__ __forceinline__ float win1024(float s_l, float c_l) {
  constexpr float cr = (float)cos64(R);   // cos(pi R/32)
  constexpr float sr = (float)sin64(R);   // sin(pi R/32)
  return fmaf(s_l, cr, c_l * sr);
}

}   // namespace bf
