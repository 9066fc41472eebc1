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
// One CTA = 128 frames x 120 directions, looping over the 514 logical bins.  Operand tiles use the canonical no-swizzle
// K-major core-matrix layout (8 rows x 16 bytes per core matrix; K-core pitch padded by 16 B against bank conflicts).
//   A (spectra)   srp_spectra_kernel writes them ONCE, already split into bf16 hi/lo and already in this layout ("images"),
//                 so a single TMA bulk copy (66 KB, mbarrier) per bin fills the tile: no CUDA-core work, issued as soon as
//                 the previous bin's UMMAs have read the tile, landing behind the steering generation
//   B (steering)  all threads; the phasors advance from bin to bin by one complex rotation (exact re-derivation every 16
//                 bins), then bf16 hi/lo split into the tile
//   thread 0      8 (K steps of 16) x 3 (split) tcgen05.mma into one of TWO TMEM accumulators (bin parity), tcgen05.commit
//   all 8 warps   epilogue of the PREVIOUS bin (tcgen05.ld of the other accumulator, power[d] += weight_l (Re^2 + Im^2) in
//                 registers) while the tensor pipe works on the current one
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

// grid = (ceil(D/120), ceil(F/128)); block = 256; maps[f][d].  xi: operand images written by srp_spectra_kernel,
// [bin][frame tile][ hi tile | lo tile ] in exactly the shared-memory layout (srp_image_bytes each).
__global__ void __launch_bounds__(256, 1) srp_power_tc_kernel(const unsigned char* __restrict__ xi, const double* __restrict__ tau,
                                                               const double* __restrict__ freqs_l, float* __restrict__ maps, int D, int M,
                                                               long long F) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* a_hi = smem_raw;
  unsigned char* a_lo = a_hi + kTileA;
  unsigned char* b_hi = a_lo + kTileA;
  unsigned char* b_lo = b_hi + kTileB;
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(b_lo + kTileB);   // [2]: UMMAs of an even / odd bin complete
  uint64_t* bar_a = bar_mma + 2;                                    // A image of the current bin has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_a + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // direction tile fastest: the CTAs that sweep the same 128 frames run together and share the spectra through L2
  const long long fbase = (long long)blockIdx.y * kTcM;
  const long long n_ft = (F + kTcM - 1) / kTcM;
  const int dbase = blockIdx.x * kTcD;
  const float invM = 1.0f / (float)M;

  if (tid == 0) { mbar_init(&bar_mma[0], 1); mbar_init(&bar_mma[1], 1); mbar_init(bar_a, 1); }
  mbar_fence_init();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // instruction descriptor (kind::f16): D = F32 [4,6) = 1, A = B = BF16 [7,10) = [10,13) = 1, both K-major, N>>3 at [17,23), M>>4 at [24,29)
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
  const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), b_hi_s = smem_u32(b_hi), b_lo_s = smem_u32(b_lo);

  // A operand: one TMA bulk copy per bin brings the ready-made image (bf16 hi | lo, core-matrix layout) of this CTA's 128 frames
  auto load_a = [&](int l) {
    mbar_expect_tx(bar_a, 2u * kTileA);
    bulk_g2s(a_hi, xi + ((size_t)l * n_ft + blockIdx.y) * (2u * kTileA), 2u * kTileA, bar_a);
  };
  if (tid == 0) load_a(0);
  // zero the B tiles once: padding rows (directions beyond D, microphones beyond M) stay zero
  for (int i = tid; i < (2 * kTileB) / 16; i += 256) reinterpret_cast<uint4*>(b_hi)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  float pw[60];
#pragma unroll
  for (int j = 0; j < 60; j++) pw[j] = 0.f;
  const int dsub = (warp >> 2) * 60;                                  // this warp's directions: dsub .. dsub+59
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;       // this warp's quarter of the TMEM lanes (= frames)

  // B operand: steering A_l[d][i] = exp(+i 2 pi f_l tau_{d,i}) / M (conj of das.cpp:41).  The frequency grid is uniform
  // (util.h:194-197) except for the last three logical bins (SURVEY B-1, B-2, B-4), so the phasor of bin l+1 is the phasor
  // of bin l times a fixed rotation; it is re-derived exactly (phase reduced in double, as before) every 16 bins and for the
  // irregular bins, which bounds the drift of the FP32 recurrence at ~1e-6.  This thread owns 15 items (direction d,
  // microphones i, i+1).
  float2 cur[15][2], rot[15][2];
  const double df = freqs_l[1] - freqs_l[0];
  auto exact = [&](int l) {
    const double fl = freqs_l[l];
#pragma unroll
    for (int q = 0; q < 15; q++) {
      const int c = tid + 256 * q, d = c >> 5, i = (c & 31) * 2;
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const bool on = dbase + d < D && i + u < M;
        const double t = on ? tau[(size_t)(dbase + d) * M + i + u] : 0.0;
        const double turns = fl * t;
        float sn, cs;
        __sincosf(6.283185307179586f * (float)(turns - rint(turns)), &sn, &cs);
        cur[q][u] = on ? make_float2(cs, sn) : make_float2(0.f, 0.f);
      }
    }
  };
#pragma unroll
  for (int q = 0; q < 15; q++) {
    const int c = tid + 256 * q, d = c >> 5, i = (c & 31) * 2;
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const bool on = dbase + d < D && i + u < M;
      const double turns = on ? df * tau[(size_t)(dbase + d) * M + i + u] : 0.0;
      float sn, cs;
      sincospif(2.0f * (float)(turns - rint(turns)), &sn, &cs);
      rot[q][u] = make_float2(cs, sn);
    }
  }

  auto epilogue = [&](int l) {   // |Y|^2 of bin l, weighted by its multiplicity (mirror bins share |y|; 0, N/2-1, N/2 and the pseudo-bin count once)
    const float wgt = (l == 0 || l >= 511) ? 1.0f : 2.0f;
    const uint32_t acc = tmem_base + lane_base + (uint32_t)((l & 1) * kTcN);
#pragma unroll
    for (int j = 0; j < 60; j += 4) {
      float re[4], im[4];
      tmem_ld4(acc + (uint32_t)(dsub + j), re);
      tmem_ld4(acc + (uint32_t)(kTcD + dsub + j), im);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int u = 0; u < 4; u++) pw[j + u] = fmaf(wgt, fmaf(re[u], re[u], im[u] * im[u]), pw[j + u]);
    }
  };

  for (int l = 0; l < kSrpBins; l++) {
    // ---- steering of bin l (the B tiles are free: the UMMAs of bin l-1 were waited for below) ----
    if ((l & 15) == 0 || l >= 511) {
      exact(l);
    } else {
#pragma unroll
      for (int q = 0; q < 15; q++)
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const float2 c = cur[q][u], r = rot[q][u];
          cur[q][u] = make_float2(fmaf(c.x, r.x, -c.y * r.y), fmaf(c.x, r.y, c.y * r.x));
        }
    }
#pragma unroll
    for (int q = 0; q < 15; q++) {
      const int c = tid + 256 * q, d = c >> 5, i = (c & 31) * 2;
      const float ar0 = cur[q][0].x * invM, ai0 = cur[q][0].y * invM, ar1 = cur[q][1].x * invM, ai1 = cur[q][1].y * invM;
      // every value sits in two rows (Re row [ Ar | -Ai ], Im row [ Ai | Ar ]): split into bf16 hi/lo once, negate by sign bit
      __nv_bfloat162 rh, rl, ih, il;
      rh.x = __float2bfloat16_rn(ar0); rh.y = __float2bfloat16_rn(ar1);
      rl.x = __float2bfloat16_rn(ar0 - __bfloat162float(rh.x)); rl.y = __float2bfloat16_rn(ar1 - __bfloat162float(rh.y));
      ih.x = __float2bfloat16_rn(ai0); ih.y = __float2bfloat16_rn(ai1);
      il.x = __float2bfloat16_rn(ai0 - __bfloat162float(ih.x)); il.y = __float2bfloat16_rn(ai1 - __bfloat162float(ih.y));
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
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();       // B tiles complete; every warp has drained the accumulator of bin l-2 (epilogue of the previous trip)
    // ---- UMMA: 8 K-steps x (hi*hi + hi*lo + lo*hi) into the accumulator of this bin's parity ----
    if (tid == 0) {
      mbar_wait_or_trap(bar_a, (uint32_t)(l & 1));   // the A image of bin l has landed
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_base + (uint32_t)((l & 1) * kTcN);
#pragma unroll
      for (int ks = 0; ks < kTcK / 16; ks++) {
        const uint32_t ka = (uint32_t)(ks * 2 * kLboA), kb = (uint32_t)(ks * 2 * kLboB);
        const uint64_t ah = umma_desc(a_hi_s + ka, kLboA, 128), al = umma_desc(a_lo_s + ka, kLboA, 128);
        const uint64_t bh = umma_desc(b_hi_s + kb, kLboB, 128), bl = umma_desc(b_lo_s + kb, kLboB, 128);
        umma_bf16(acc, ah, bh, idesc, ks > 0 ? 1u : 0u);
        umma_bf16(acc, ah, bl, idesc, 1u);
        umma_bf16(acc, al, bh, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma[l & 1])) : "memory");
    }
    // ---- the epilogue of the PREVIOUS bin runs on the CUDA cores while the tensor pipe works on this one ----
    if (l > 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      epilogue(l - 1);
    }
    mbar_wait_or_trap(&bar_mma[l & 1], (uint32_t)((l >> 1) & 1));   // UMMAs of bin l done: A and B tiles are free
    if (tid == 0 && l + 1 < kSrpBins) load_a(l + 1);                // lands behind the next bin's steering generation
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  epilogue(kSrpBins - 1);
  {
    const long long f = fbase + (warp & 3) * 32 + lane;
    if (f < F) {
#pragma unroll
      for (int j = 0; j < 60; j++)
        if (dbase + dsub + j < D) maps[(size_t)f * D + dbase + dsub + j] = pw[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

size_t srp_image_bytes() { return 2 * (size_t)kTileA; }

cudaError_t launch_srp_power_tc(const unsigned char* xi, const double* tau, const double* freqs_l, float* maps, int D, int M, long long F,
                                cudaStream_t st) {
  const size_t smem = 2 * (size_t)kTileA + 2 * (size_t)kTileB + 64;
  cudaError_t e = cudaFuncSetAttribute(srp_power_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)((D + kTcD - 1) / kTcD), (unsigned)((F + kTcM - 1) / kTcM));
  srp_power_tc_kernel<<<grid, 256, smem, st>>>(xi, tau, freqs_l, maps, D, M, F);
  return cudaGetLastError();
}

}   // namespace bf
