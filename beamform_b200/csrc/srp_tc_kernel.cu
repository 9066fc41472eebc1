// Steered-response power on the 5th-generation tensor cores (tcgen05 / UMMA, accumulators in TMEM), sm_100a.
//
// Per bin l the sweep is a dense complex contraction  Y_l[f][d] = sum_i X_l[f][i] * A_l[d][i]  (das.cpp:61-62 with the
// look direction swept; A_l[d][i] = conj(w_{d,i}[l]) / M).  It is mapped onto real UMMAs with the complex parts
// concatenated along K (K' = 2 * 64 microphones):
//     rows  (M = 128)  : frames f,                 A-operand row  f      = [ Xr[f][:] |  Xi[f][:] ]
//     cols  (N = 240)  : n <  120  -> Re Y[d=n],   B-operand row  n      = [ Ar[d][:] | -Ai[d][:] ]
//                        n >= 120  -> Im Y[d],     B-operand row  120+d  = [ Ai[d][:] |  Ar[d][:] ]
// so one thread of the epilogue (TMEM lane = frame) sees Re and Im of every direction of its frame.
// Precision: north_star wants 1e-4 on the maps; BF16 alone gives 4e-3, so both operands are split hi + lo and three
// UMMAs (hi*hi + hi*lo + lo*hi) accumulate into the same FP32 TMEM tile (error ~2^-17 per term).
//
// One CTA = 128 frames x 120 directions, looping over the 514 logical bins:
//   all threads   build the two operand tiles in shared memory (spectra from XS -> bf16 hi/lo; steering generated from
//                 the delay table with the phase reduced in double), in the canonical no-swizzle K-major core-matrix
//                 layout (8 rows x 16 bytes per core matrix; K-core pitch padded by 16 B against bank conflicts)
//   thread 0      issues 8 (K steps of 16) x 3 (split) tcgen05.mma, then tcgen05.commit -> mbarrier
//   all 8 warps   wait, tcgen05.ld their quarter of the lanes (warps 0-3: directions 0-59, warps 4-7: 60-119),
//                 power[d] += weight_l * (Re^2 + Im^2) in registers
// This first tensor-core version is synchronous per bin (no overlap of operand generation, UMMA and epilogue).
#include <cuda_bf16.h>

#include "async_copy.cuh"
#include "bf_device.h"

namespace bf {

constexpr int kTcM = 128;                 // frames per CTA (UMMA M)
constexpr int kTcD = 120;                 // directions per CTA
constexpr int kTcN = 2 * kTcD;            // UMMA N
constexpr int kTcK = 128;                 // 2 * 64 microphones
constexpr int kLboA = (kTcM / 8) * 128 + 16;   // bytes between K-cores (padded)
constexpr int kLboB = (kTcN / 8) * 128 + 16;
constexpr int kTileA = (kTcK / 8) * kLboA;      // bytes per A tile (hi or lo)
constexpr int kTileB = (kTcK / 8) * kLboB;
constexpr int kSrpBins = 514;

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SM100 shared-memory matrix descriptor: start address [0,14), leading byte offset [16,30), stride byte offset
  // [32,46) (all >> 4), version [46,48) = 1, layout type [61,64) = 0 (no swizzle)
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}

// bounded wait: a UMMA that never completes must end in a trap, not in a hung GPU
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
  for (unsigned spin = 0; spin < (1u << 20); spin++) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

// element (row, k) of an operand tile: K-core (k/8) | row group (row/8) | row in group | element in the 16-byte row
__device__ __forceinline__ uint32_t tile_off(int row, int k, int lbo) { return (uint32_t)((k >> 3) * lbo + (row >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2); }

__device__ __forceinline__ void split_store2(unsigned char* hi, unsigned char* lo, uint32_t off, float a, float b) {
  const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
  const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah)), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
  __nv_bfloat162 h2, l2;
  h2.x = ah; h2.y = bh; l2.x = al; l2.y = bl;
  *reinterpret_cast<__nv_bfloat162*>(hi + off) = h2;
  *reinterpret_cast<__nv_bfloat162*>(lo + off) = l2;
}

// grid = (ceil(D/120), ceil(F/128)); block = 256; maps[f][d]
__global__ void __launch_bounds__(256, 1) srp_power_tc_kernel(const float2* __restrict__ xs, const double* __restrict__ tau,
                                                               const double* __restrict__ freqs_l, float* __restrict__ maps, int D, int M,
                                                               long long F) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* a_hi = smem_raw;
  unsigned char* a_lo = a_hi + kTileA;
  unsigned char* b_hi = a_lo + kTileA;
  unsigned char* b_lo = b_hi + kTileB;
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_lo + kTileB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // direction tile fastest: the CTAs that sweep the same 128 frames run together and share the spectra through L2
  const long long fbase = (long long)blockIdx.y * kTcM;
  const int dbase = blockIdx.x * kTcD;
  const float invM = 1.0f / (float)M;

  if (tid == 0) mbar_init(bar, 1);
  mbar_fence_init();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // instruction descriptor (kind::f16): D = F32 [4,6) = 1, A = B = BF16 [7,10) = [10,13) = 1, both K-major, N>>3 at [17,23), M>>4 at [24,29)
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
  const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), b_hi_s = smem_u32(b_hi), b_lo_s = smem_u32(b_lo);

  // zero the operand tiles once: padding rows (frames beyond F, directions beyond D, microphones beyond M) stay zero
  for (int i = tid; i < (2 * kTileA + 2 * kTileB) / 16; i += 256) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  float pw[60];
#pragma unroll
  for (int j = 0; j < 60; j++) pw[j] = 0.f;
  const int dsub = (warp >> 2) * 60;                                  // this warp's directions: dsub .. dsub+59
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;       // this warp's quarter of the TMEM lanes (= frames)

  // this thread's 16 spectra chunks (frame m, microphones i, i+1) and 15 steering items (direction d, microphones i, i+1)
  // keep their delays in registers; the spectra of the NEXT bin are prefetched into registers behind the UMMAs
  double tau_r[15][2];
#pragma unroll
  for (int q = 0; q < 15; q++) {
    const int c = tid + 256 * q, d = c >> 5, i = (c & 31) * 2;
#pragma unroll
    for (int u = 0; u < 2; u++) tau_r[q][u] = (dbase + d < D && i + u < M) ? tau[(size_t)(dbase + d) * M + i + u] : 0.0;
  }
  float4 xr[16];
  auto load_xs = [&](int l) {
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const int c = tid + 256 * q, m = c >> 5, i = (c & 31) * 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (fbase + m < F && i < M) {
        const float2* src = xs + ((size_t)l * F + fbase + m) * M + i;
        const float2 x0 = __ldg(src);
        v.x = x0.x; v.y = x0.y;
        if (i + 1 < M) { const float2 x1 = __ldg(src + 1); v.z = x1.x; v.w = x1.y; }
      }
      xr[q] = v;
    }
  };
  load_xs(0);

  for (int l = 0; l < kSrpBins; l++) {
    // ---- A operand: spectra of 128 frames, [Xr | Xi], bf16 hi/lo ----
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const int c = tid + 256 * q, m = c >> 5, i = (c & 31) * 2;
      split_store2(a_hi, a_lo, tile_off(m, i, kLboA), xr[q].x, xr[q].z);            // Xr[i], Xr[i+1]
      split_store2(a_hi, a_lo, tile_off(m, 64 + i, kLboA), xr[q].y, xr[q].w);       // Xi[i], Xi[i+1]
    }
    // ---- B operand: steering A_l[d][i] = exp(+i 2 pi f_l tau_{d,i}) / M (conj of das.cpp:41): the phase is reduced to
    //      [-1/2, 1/2] turn in double, then MUFU sin/cos (abs. error ~1e-6, two orders below the 1e-4 budget) ----
    const double fl = freqs_l[l];
#pragma unroll
    for (int q = 0; q < 15; q++) {
      const int c = tid + 256 * q, d = c >> 5, i = (c & 31) * 2;
      float ar[2], ai[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const double turns = fl * tau_r[q][u];
        const float ang = 6.283185307179586f * (float)(turns - rint(turns));
        float sn, cs;
        __sincosf(ang, &sn, &cs);
        const bool on = dbase + d < D && i + u < M;
        ar[u] = on ? cs * invM : 0.f;
        ai[u] = on ? sn * invM : 0.f;
      }
      // every value sits in two rows (Re row [ Ar | -Ai ], Im row [ Ai | Ar ]): split into bf16 hi/lo once, negate by sign bit
      __nv_bfloat162 rh, rl, ih, il;
      rh.x = __float2bfloat16_rn(ar[0]); rh.y = __float2bfloat16_rn(ar[1]);
      rl.x = __float2bfloat16_rn(ar[0] - __bfloat162float(rh.x)); rl.y = __float2bfloat16_rn(ar[1] - __bfloat162float(rh.y));
      ih.x = __float2bfloat16_rn(ai[0]); ih.y = __float2bfloat16_rn(ai[1]);
      il.x = __float2bfloat16_rn(ai[0] - __bfloat162float(ih.x)); il.y = __float2bfloat16_rn(ai[1] - __bfloat162float(ih.y));
      const uint32_t rhu = *reinterpret_cast<uint32_t*>(&rh), rlu = *reinterpret_cast<uint32_t*>(&rl);
      const uint32_t ihu = *reinterpret_cast<uint32_t*>(&ih), ilu = *reinterpret_cast<uint32_t*>(&il);
      const uint32_t o_re = tile_off(d, i, kLboB), o_im = tile_off(kTcD + d, i, kLboB);
      const uint32_t kh = (uint32_t)(8 * kLboB);   // + 64 along K: eight K-cores further
      *reinterpret_cast<uint32_t*>(b_hi + o_re) = rhu;                 *reinterpret_cast<uint32_t*>(b_lo + o_re) = rlu;
      *reinterpret_cast<uint32_t*>(b_hi + o_re + kh) = ihu ^ 0x80008000u; *reinterpret_cast<uint32_t*>(b_lo + o_re + kh) = ilu ^ 0x80008000u;
      *reinterpret_cast<uint32_t*>(b_hi + o_im) = ihu;                 *reinterpret_cast<uint32_t*>(b_lo + o_im) = ilu;
      *reinterpret_cast<uint32_t*>(b_hi + o_im + kh) = rhu;            *reinterpret_cast<uint32_t*>(b_lo + o_im + kh) = rlu;
    }
    fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
    __syncthreads();
    // ---- UMMA: 8 K-steps x (hi*hi + hi*lo + lo*hi) into one FP32 accumulator ----
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < kTcK / 16; ks++) {
        const uint32_t ka = (uint32_t)(ks * 2 * kLboA), kb = (uint32_t)(ks * 2 * kLboB);
        const uint64_t ah = umma_desc(a_hi_s + ka, kLboA, 128), al = umma_desc(a_lo_s + ka, kLboA, 128);
        const uint64_t bh = umma_desc(b_hi_s + kb, kLboB, 128), bl = umma_desc(b_lo_s + kb, kLboB, 128);
        umma_bf16(tmem_base, ah, bh, idesc, ks > 0 ? 1u : 0u);
        umma_bf16(tmem_base, ah, bl, idesc, 1u);
        umma_bf16(tmem_base, al, bh, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    if (l + 1 < kSrpBins) load_xs(l + 1);   // in flight behind the UMMAs and the epilogue
    mbar_wait_or_trap(bar, (uint32_t)(l & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: |Y|^2 weighted by the bin's multiplicity (mirror bins share |y|; 0, N/2-1, N/2 and the pseudo-bin count once) ----
    const float wgt = (l == 0 || l >= 511) ? 1.0f : 2.0f;
#pragma unroll
    for (int j = 0; j < 60; j += 4) {
      float re[4], im[4];
      tmem_ld4(tmem_base + lane_base + (uint32_t)(dsub + j), re);
      tmem_ld4(tmem_base + lane_base + (uint32_t)(kTcD + dsub + j), im);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int u = 0; u < 4; u++) pw[j + u] = fmaf(wgt, fmaf(re[u], re[u], im[u] * im[u]), pw[j + u]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // accumulator drained and operand tiles free for the next bin
  }
  {
    const long long f = fbase + (warp & 3) * 32 + lane;
    if (f < F) {
#pragma unroll
      for (int j = 0; j < 60; j++)
        if (dbase + dsub + j < D) maps[(size_t)f * D + dbase + dsub + j] = pw[j];
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
}

cudaError_t launch_srp_power_tc(const float2* xs, const double* tau, const double* freqs_l, float* maps, int D, int M, long long F,
                                cudaStream_t st) {
  const size_t smem = 2 * (size_t)kTileA + 2 * (size_t)kTileB + 64;
  cudaError_t e = cudaFuncSetAttribute(srp_power_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)((D + kTcD - 1) / kTcD), (unsigned)((F + kTcM - 1) / kTcM));
  srp_power_tc_kernel<<<grid, 256, smem, st>>>(xs, tau, freqs_l, maps, D, M, F);
  return cudaGetLastError();
}

}   // namespace bf
