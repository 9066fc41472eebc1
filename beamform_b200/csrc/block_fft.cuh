// Shared-memory Stockham FFT of the CTA-per-stream kernels (generic_kernel.cu, phase_n_kernel.cu): radix-8/16 register
// butterflies (fft_reg.cuh), run-time pass loop, swizzled element index.  See the notes at each piece.
#pragma once
#include <type_traits>

#include "bf_device.h"
#include "fft_reg.cuh"
#include "fft_reg_d.cuh"
#include "warp_fft1024.cuh"

namespace bf {

constexpr int kGenThreads = 256;   // default CTA size of the CTA-per-stream kernels (the in-place passes take the actual size as template argument T)

// sqrt through MUFU.SQRT (~2 ulp; NaN / negative / zero inputs behave as sqrtf).  The IEEE sqrtf is a ~10-instruction sequence
// with a slow-path branch; the per-bin stages take 5-8 square roots per bin and none of them feeds an exact decision without a
// guard band far wider than 2 ulp.
__device__ __forceinline__ float sqrt_fast(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// XOR swizzle of a transform's element index (elements are 8 or 16 bytes): the first Stockham pass scatters with a
// stride of R elements, which without it lands every lane of a warp in the same shared-memory bank group.
__device__ __forceinline__ int swz(int idx) { return idx ^ ((idx >> 4) & 7); }


template <typename V> struct VecOps;
template <> struct VecOps<float2> {
  static __device__ __forceinline__ float2 mul(float2 a, float2 w) { return cmul(a, w); }
  static __device__ __forceinline__ float2 mulc(float2 a, float2 w) { return cmulc(a, w); }
  template <int R, int DIR> static __device__ __forceinline__ void fft(float2* v) { fft_dit<R, DIR>(v); }
};
template <> struct VecOps<double2> {
  static __device__ __forceinline__ double2 mul(double2 a, double2 w) { return cmul_d(a, w); }
  static __device__ __forceinline__ double2 mulc(double2 a, double2 w) { return cmulc_d(a, w); }
  template <int R, int DIR> static __device__ __forceinline__ void fft(double2* v) { fft_dit_d<R, DIR>(v); }
};

// ---- shared-memory Stockham FFT, code-size first ----
// The kernel's hot loop has to live in the instruction caches (8 warps per SM execute it once per frame pair: with
// every pass unrolled into its own copy the 4096-point PhaseMPF kernel was 1.2 MB of SASS and a third of all issue
// slots waited for instruction fetch).  So the pass index is a RUN-TIME loop variable: one copy of the radix-R
// gather/twiddle/butterfly body and one of the scatter per radix and element type.
//   pass with Ns = 1 << sh:  j in [0, NN/R): k = j mod Ns; v[q] = z[j + q*NN/R] * W_{Ns*R}^{k q}; V = DFT_R(v);
//                            z[(j - k)*R + k + q*Ns] = V[q]
template <int NN, int R, int DIR, typename V>
struct Step {
  static constexpr int per = NN / R;   // tasks per transform
  static __device__ __forceinline__ void load(const V* zz, const V* __restrict__ tw, int j, int sh, V (&v)[R]) {
    const int k = j & ((1 << sh) - 1);
    // Twiddles W^q, q = 1..R-1, W = W_NN^m, m = k * per >> sh (q m < NN: W^q = tw[q m]).  With 225 KB of the SM given to
    // shared memory the tables do not stay in L1 and every load pays an L2 round trip, so only a few are loaded.
    // FP32: the power-of-two members W, W^2, W^4, W^8 come from the table (correctly rounded), the others are products of
    // those along the bits of q (at most 3 multiplications deep): ~2 ulp, and only the four table values stay live (a
    // full product tree held 15 twiddles in registers and spilled under the 80-register budget of 3 CTAs per SM).
    // FP64: running product W^q = W^(q-1) * W (two live values; 15 roundings of 1e-16).
    constexpr bool kDouble = sizeof(V) == sizeof(double2);
    constexpr int LOGR = ilog2(R);
    V w[kDouble ? 2 : LOGR];
    if (sh > 0) {   // the first pass has k = 0: unit twiddles
      const int m = k * (per >> sh);
      if constexpr (!kDouble) {
#pragma unroll
        for (int b = 0; b < LOGR; b++) w[b] = __ldg(tw + (m << b));
      } else {
        w[1] = __ldg(tw + m);
        w[0] = w[1];
      }
    }
    static_for<0, R>([&](auto qc) {
      constexpr int q = decltype(qc)::value;
      V a = zz[swz(j + q * per)];
      if (q > 0 && sh > 0) {
        if constexpr (kDouble) {
          a = (DIR < 0) ? VecOps<V>::mul(a, w[0]) : VecOps<V>::mulc(a, w[0]);
          if (q + 1 < R) w[0] = VecOps<V>::mul(w[0], w[1]);
        } else {
          constexpr int b0 = ilog2(q & -q);   // lowest set bit
          V wq = w[b0];
          static_for<b0 + 1, LOGR>([&](auto bc) {
            constexpr int b = decltype(bc)::value;
            if constexpr ((q >> b) & 1) wq = VecOps<V>::mul(wq, w[b]);
          });
          a = (DIR < 0) ? VecOps<V>::mul(a, wq) : VecOps<V>::mulc(a, wq);
        }
      }
      v[brev(q, ilog2(R))] = a;
    });
    VecOps<V>::template fft<R, DIR>(v);
  }
  static __device__ __forceinline__ void store(V* zz, int j, int sh, const V (&v)[R]) {
    const int k = j & ((1 << sh) - 1);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int q = 0; q < R; q++) zz[swz(j0 + (q << sh))] = v[q];
  }
};

// n_pass in-place passes of radix R, Ns = 1 << sh0, then * R per pass, over nfft transforms stored back to back.
// Every read of a round precedes every write: two block barriers per round.
template <int NN, int R, int DIR, typename V, int T = 256>
__device__ __forceinline__ void passes_inplace(V* z, int nfft, const V* __restrict__ tw, int tid, int sh0, int n_pass) {
  constexpr int kGenThreads = T;   // threads of the CTA (shadows the default)
  typedef Step<NN, R, DIR, V> S;
  constexpr int per = S::per;
  constexpr int g = kGenThreads / per > 0 ? kGenThreads / per : 1;   // transforms per round
  static_assert(per <= kGenThreads, "one round must cover a whole transform");
  const int f_local = tid / per, j = tid - f_local * per;
  const int rounds = (nfft + g - 1) / g;
#pragma unroll 1
  for (int it = 0; it < n_pass * rounds; it++) {
    const int pass = it / rounds, f = (it - pass * rounds) * g + f_local;
    const int sh = sh0 + pass * ilog2(R);
    const bool on = f_local < g && f < nfft;
    V v[R];
    V* zz = z + (size_t)f * NN;
    if (on) S::load(zz, tw, j, sh, v);
    __syncthreads();
    if (on) S::store(zz, j, sh, v);
    __syncthreads();
  }
}

template <int NN, int DIR, typename V, int T = 256>
__device__ __forceinline__ void block_fft(V* z, int nfft, const V* __restrict__ tw, int tid) {
  if constexpr (NN == 4096) {
    passes_inplace<NN, 16, DIR, V, T>(z, nfft, tw, tid, 0, 3);
  } else if constexpr (NN == 2048) {
    passes_inplace<NN, 16, DIR, V, T>(z, nfft, tw, tid, 0, 2);
    passes_inplace<NN, 8, DIR, V, T>(z, nfft, tw, tid, 8, 1);
  } else if constexpr (NN == 1024) {
    passes_inplace<NN, 16, DIR, V, T>(z, nfft, tw, tid, 0, 1);
    passes_inplace<NN, 8, DIR, V, T>(z, nfft, tw, tid, 4, 2);
  } else {
    static_assert(NN == 512, "supported frame sizes: 512, 1024, 2048, 4096");
    passes_inplace<NN, 8, DIR, V, T>(z, nfft, tw, tid, 0, 3);
  }
}

// Out-of-place passes of ONE transform, ping-pong between two buffers: writes cannot clobber the pass's own reads,
// so one barrier per pass.  Returns the buffer that holds the result.
template <int NN, int R, int DIR, typename V>
__device__ __forceinline__ V* passes_oop(V* src, V* dst, const V* __restrict__ tw, int tid, int sh0, int n_pass) {
  typedef Step<NN, R, DIR, V> S;
  static_assert(S::per <= kGenThreads, "one round must cover a whole transform");
#pragma unroll 1
  for (int pass = 0; pass < n_pass; pass++) {
    if (tid < S::per) {
      V v[R];
      S::load(src, tw, tid, sh0 + pass * ilog2(R), v);
      S::store(dst, tid, sh0 + pass * ilog2(R), v);
    }
    __syncthreads();
    V* t = src; src = dst; dst = t;
  }
  return src;
}
template <int NN, int DIR, typename V>
__device__ __forceinline__ V* block_fft_oop(V* a, V* b, const V* __restrict__ tw, int tid) {
  if constexpr (NN == 4096) {
    return passes_oop<NN, 16, DIR, V>(a, b, tw, tid, 0, 3);
  } else if constexpr (NN == 2048) {
    V* r = passes_oop<NN, 16, DIR, V>(a, b, tw, tid, 0, 2);
    return passes_oop<NN, 8, DIR, V>(r, r == a ? b : a, tw, tid, 8, 1);
  } else if constexpr (NN == 1024) {
    V* r = passes_oop<NN, 16, DIR, V>(a, b, tw, tid, 0, 1);
    return passes_oop<NN, 8, DIR, V>(r, r == a ? b : a, tw, tid, 4, 2);
  } else {
    return passes_oop<NN, 8, DIR, V>(a, b, tw, tid, 0, 3);
  }
}

// ---- out-of-line entry points ----
// Each heavy stage of the frame-pair loop is its own (non-inlined) function: ptxas then allocates registers per stage
// instead of across the whole loop (inlined, the stages' live ranges pushed the FP64 radix-16 butterflies into spills
// and the stage time moved by 50 % from build to build).  Buffers are named by their byte offset in the dynamic
// shared-memory window, so the accesses stay LDS/STS.
extern __shared__ __align__(16) unsigned char gen_smem_raw[];

template <int NN, int DIR, typename V, int T = 256>
__device__ __noinline__ void block_fft_fn(unsigned z_off, int nfft, const V* __restrict__ tw, int tid) {
  block_fft<NN, DIR, V, T>(reinterpret_cast<V*>(gen_smem_raw + z_off), nfft, tw, tid);
}
// returns the byte offset of the buffer that holds the result
template <int NN, int DIR, typename V>
__device__ __noinline__ unsigned block_fft_oop_fn(unsigned a_off, unsigned b_off, const V* __restrict__ tw, int tid) {
  V* a = reinterpret_cast<V*>(gen_smem_raw + a_off);
  V* r = block_fft_oop<NN, DIR, V>(a, reinterpret_cast<V*>(gen_smem_raw + b_off), tw, tid);
  return r == a ? a_off : b_off;
}
}   // namespace bf
