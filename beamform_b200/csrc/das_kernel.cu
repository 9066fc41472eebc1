// Fused delay-and-sum kernel for 1024-point frames, sm_100a: every warp is an independent worker.
//
//   reference path replaced (citations /root/reference/beamform/src/): util.h:217-242 (window + framing),
//   das.cpp:47-70 (M forward FFTs, per-bin weight-and-sum, inverse FFT), util.h:244-253,289-314 (synthesis
//   window, 50 % overlap-add).
//
// Work = the global sequence of frame PAIRS (stream-major); it is cut into gridDim*warps equal contiguous
// ranges, one per warp.  For each pair (frames t, t+1 of one stream) a warp
//   1. receives hops t-1..t+1 of microphone i in its private shared-memory tile (TMA bulk copy, mbarrier;
//      the copy for microphone i+1 is in flight while microphone i is transformed),
//   2. windows them into z = 0.5*w*(frame_t + i*frame_{t+1}) and runs a warp-private 1024-point FFT
//      (32 points per lane in registers, one swizzled shared-memory transpose),
//   3. accumulates G += ceff_i .* Z_i in registers (ceff = Hermitian-ised conj(w)/M, host double),
//   4. after the last microphone runs the inverse FFT straight from the accumulator registers,
//      applies the synthesis window and overlap-adds with the previous frame's tail, all in registers.
// No block-level barrier exists after start-up; spectra never leave the SM; every input hop is fetched from
// HBM once (its second use, as "previous hop" of the next pair, hits L2) and every output sample is written once.
// A range that starts inside a stream first recomputes the pair before it (no stores) to obtain the OLA tail.
#include <cstdio>
#include <cstdlib>

#include "async_copy.cuh"
#include "bf_device.h"
#include "fft_reg.cuh"
#include "tmem.cuh"
#include "warp_fft1024.cuh"

namespace bf {

constexpr int kTileF2 = 1024;   // float2 per warp tile (8 KB, XOR-swizzled: no padding)

// (the transform itself is warp_fft1024_fwd in warp_fft1024.cuh)

struct PairPos {
  int s;      // stream (global index)
  int t;      // first hop of the pair inside this launch's arrays
  bool two;   // frame t+1 exists
};

// kTiles = 2: the next job's hops land in the second tile while the current job transforms (round 1).  kTiles = 1: they land in
// the SAME tile as soon as the transform's exchange step has read it (behind the second FFT pass and the accumulation), which
// halves the shared memory per warp (12 warps fit beside the weights; measured slower: the kernel is bound by the issue rate of
// the butterflies, not by latency, and 12 warps leave 168 registers per thread).
// kTmem: the per-pair accumulator G (32 complex values per lane) lives in tensor memory instead of registers (tcgen05.ld / st,
// 64 columns per warp in its lane quarter): the kernel then fits 128 registers per thread and SIXTEEN warps per SM, which fill
// the issue slots two warps per scheduler leave empty (round 1: 64 % issue utilisation, stalls "wait" and "short scoreboard").
template <int kWarps, int kTiles, bool kTmem>
__global__ void __launch_bounds__(kWarps * 32, 1) das_pairs_kernel(const KernelParams p, const int ceff_in_smem) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* tw = reinterpret_cast<float2*>(smem_raw);                         // [32][32]
  float2* tiles = tw + 1024;                                                  // [kWarps][kTiles][1024]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)kWarps * kTiles * kTileF2);   // [kWarps][kTiles]
  float* tails = reinterpret_cast<float*>(bars + kWarps * kTiles);          // [kWarps][512] OLA tails
  float2* ceff_s = reinterpret_cast<float2*>(tails + kWarps * 512);         // [M][1024] when it fits
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = p.M;
  constexpr int H = 512;

  for (int i = tid; i < 1024; i += blockDim.x) {
    const int k1 = i >> 5, l = i & 31;
    float sn, cs;
    sincospif(-2.0f * (float)((k1 * l) & 1023) / 1024.0f, &sn, &cs);
    tw[i] = make_float2(cs, sn);
  }
  if (ceff_in_smem)
    for (int i = tid; i < M * 1024; i += blockDim.x) ceff_s[i] = p.das_ceff[i];
  if (lane == 0)
    for (int b = 0; b < kTiles; b++) mbar_init(&bars[warp * kTiles + b], 1);
  mbar_fence_init();
  if (kTmem && warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (kTmem) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (kTmem) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tacc = kTmem ? tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64) : 0u;   // this warp's 64 columns
  const float2* ceff = ceff_in_smem ? ceff_s : p.das_ceff;
  float2* mytile = tiles + (size_t)warp * kTiles * kTileF2;
  uint64_t* mybar = bars + warp * kTiles;

  double sd, cd;
  sincospi((double)lane / 1024.0, &sd, &cd);
  const float s_l = (float)(0.5 * sd), c_l = (float)(0.5 * cd);                           // analysis window * 0.5
  const float s_o = (float)(2.0 * sd * p.out_scale), c_o = (float)(2.0 * cd * p.out_scale);   // synthesis window * 2/N

  // ---- this warp's range of the global pair sequence ----
  const int nh = p.hop_end - p.hop_begin;
  const int Tp = (nh + 1) >> 1;
  const long long total = (long long)p.n_streams * Tp;
  const long long W = (long long)gridDim.x * kWarps, w = (long long)blockIdx.x * kWarps + warp;
  const long long g_begin = total * w / W, g_end = total * (w + 1) / W;
  if (g_begin < g_end) {
  const bool warm = (g_begin % Tp) != 0;
  const long long g_first = warm ? g_begin - 1 : g_begin;

  // position bookkeeping without per-job 64-bit divisions: (stream, pair-in-stream) advance incrementally
  auto make_pos = [&](int sl, int q) {
    PairPos pp;
    pp.s = sl + p.stream_begin;
    pp.t = p.hop_begin + 2 * q;
    pp.two = pp.t + 1 < p.hop_end;
    return pp;
  };
  int cur_sl = (int)(g_first / Tp), cur_q = (int)(g_first - (long long)cur_sl * Tp);
  // TMA stage-in of hops t-1..t+1 of one microphone: contiguous in the input array except at stream start,
  // where hop -1 comes from the per-stream state (util.h:275-277: zeros before the first call).
  auto issue = [&](const PairPos pp, int ch, int b) {
    if (lane != 0) return;
    const float* base = p.in + (size_t)pp.s * p.in_stream_stride + (size_t)ch * p.in_mic_stride;
    float* dst = reinterpret_cast<float*>(mytile + (size_t)b * kTileF2);
    const uint32_t nb = (pp.two ? 2u : 1u) * H * 4u;
    mbar_expect_tx(&mybar[b], nb + H * 4u);
    const float* prev = (pp.t - 1 < 0) ? p.prev_hop + ((size_t)pp.s * M + ch) * H : base + (size_t)(pp.t - 1) * H;
    bulk_g2s(dst, prev, H * 4u, &mybar[b]);
    bulk_g2s(dst + H, base + (size_t)pp.t * H, nb, &mybar[b]);
  };

  // OLA tail (second half of the last synthesised frame, util.h:301-302): per-warp shared-memory line between
  // pairs, loaded from / persisted to the per-stream state at stream boundaries.
  float* mytail = tails + warp * H;
  unsigned job = 0;   // jobs alternate between the two tiles; barrier parity = (job >> 1) & 1
  issue(make_pos(cur_sl, cur_q), 0, 0);

  for (long long g = g_first; g < g_end; g++) {
    const PairPos pp = make_pos(cur_sl, cur_q);
    if (++cur_q == Tp) { cur_q = 0; cur_sl++; }   // (cur_sl, cur_q) now name the NEXT pair
    const bool write = g >= g_begin;
    if (pp.t == p.hop_begin) {   // stream start: tail of the previous call (zeros initially, util.h:285)
#pragma unroll
      for (int m2 = 0; m2 < 16; m2++) mytail[32 * m2 + lane] = p.tail[(size_t)pp.s * H + 32 * m2 + lane];
    }
    const bool last = pp.t + 2 >= p.hop_end;   // stream end: persist the tail as state for the next call
    float2 acc[kTmem ? 1 : 32];
    if (!kTmem) {
#pragma unroll
      for (int k = 0; k < (kTmem ? 1 : 32); k++) acc[k] = make_float2(0.f, 0.f);
    }

    // M forward jobs and one inverse job run through the same transform code (ch == M: inverse)
#pragma unroll 1
    for (int ch = 0; ch <= M; ch++) {
      const bool fwd = ch < M;
      float2 v[32];
      float2* tile;
      const PairPos pnext = make_pos(cur_sl, cur_q);
      if (fwd) {
        const int b = (kTiles == 2) ? (int)(job & 1) : 0;
        if (kTiles == 2) {
          // the other tile is free (its transform finished in the previous job): start the next stage-in
          fence_proxy_async();
          __syncwarp();
          if (ch + 1 < M) issue(pp, ch + 1, b ^ 1);
          else if (g + 1 < g_end) issue(pnext, 0, b ^ 1);
        }
        tile = mytile + (size_t)b * kTileF2;
        const float* stg = reinterpret_cast<const float*>(tile);
        mbar_wait(&mybar[b], (kTiles == 2) ? ((job >> 1) & 1) : (job & 1));
        static_for<0, 16>([&](auto r) {
          const float a = stg[32 * r + lane], bb = stg[512 + 32 * r + lane];
          const float c = pp.two ? stg[1024 + 32 * r + lane] : 0.0f;
          const float w0 = win1024<r>(s_l, c_l);        // 0.5 * w[32r + lane]
          const float w1 = win1024<r + 16>(s_l, c_l);   // 0.5 * w[32r + lane + 512]
          v[brev5(r)] = make_float2(a * w0, bb * w0);
          v[brev5(r + 16)] = make_float2(bb * w1, c * w1);
        });
        __syncwarp();   // staged samples consumed: the tile becomes the exchange buffer
        job++;
      } else {
        // inverse straight from the accumulators (stage 1 wants g[32*n1 + lane] in slot brev5(n1)), parts swapped
        tile = mytile + ((kTiles == 2) ? (size_t)((job - 1) & 1) * kTileF2 : 0);
        if constexpr (kTmem) {
          static_for<0, 4>([&](auto cc) {
            uint32_t r[16];
            tmem_ld16(tacc + 16u * cc, r);
            tmem_wait_ld();
            static_for<0, 8>([&](auto u) { v[brev5(8 * cc + u)] = make_float2(__uint_as_float(r[2 * u + 1]), __uint_as_float(r[2 * u])); });
          });
        } else {
          static_for<0, 32>([&](auto n1) { v[brev5(n1)] = make_float2(acc[kTmem ? 0 : (int)n1].y, acc[kTmem ? 0 : (int)n1].x); });
        }
        __syncwarp();
      }
      warp_fft1024_fwd(v, tile, tw, lane, [&]() {
        if (kTiles == 1) {
          // one tile: the next job's hops may land once this transform no longer needs the tile.  The job before the inverse
          // issues nothing (the inverse exchanges through the tile too); the inverse issues the next pair's first microphone.
          const bool nxt_mic = ch + 1 < M, nxt_pair = (ch == M) && (g + 1 < g_end);
          if (nxt_mic || nxt_pair) {
            fence_proxy_async();
            if (nxt_mic) issue(pp, ch + 1, 0); else issue(pnext, 0, 0);
          }
        }
      });
      if (fwd) {
        // das.cpp:60-63 on the packed spectrum: G[j] += ceff_i[j] * Z_i[j],  j = lane + 32*k2
        const float2* cw = ceff + (size_t)ch * 1024 + lane;
        if constexpr (kTmem) {
          static_for<0, 4>([&](auto cc) {   // 8 bins (16 columns) at a time: load the running sums, accumulate, store back
            uint32_t r[16];
            if (ch > 0) {
              tmem_ld16(tacc + 16u * cc, r);
              tmem_wait_ld();
            } else {
#pragma unroll
              for (int u = 0; u < 16; u++) r[u] = 0u;
            }
            static_for<0, 8>([&](auto u) {
              constexpr int k2 = 8 * cc + u;
              const float2 c = cw[32 * k2];
              float ax = __uint_as_float(r[2 * u]), ay = __uint_as_float(r[2 * u + 1]);
              ax = fmaf(v[k2].x, c.x, ax); ax = fmaf(-v[k2].y, c.y, ax);
              ay = fmaf(v[k2].x, c.y, ay); ay = fmaf(v[k2].y, c.x, ay);
              r[2 * u] = __float_as_uint(ax); r[2 * u + 1] = __float_as_uint(ay);
            });
            tmem_st16(tacc + 16u * cc, r);
          });
          tmem_wait_st();
        } else {
#pragma unroll
          for (int k2 = 0; k2 < (kTmem ? 0 : 32); k2++) {
            const float2 c = cw[32 * k2];
            acc[k2].x = fmaf(v[k2].x, c.x, acc[k2].x); acc[k2].x = fmaf(-v[k2].y, c.y, acc[k2].x);
            acc[k2].y = fmaf(v[k2].x, c.y, acc[k2].y); acc[k2].y = fmaf(v[k2].y, c.x, acc[k2].y);
          }
        }
      } else {
        // v = swap(IFFT(G)): frame t in .y, frame t+1 in .x; synthesis window + overlap-add (util.h:244-253,301-302)
        float* o0 = p.out + (size_t)pp.s * p.out_stream_stride + (size_t)pp.t * H;
        static_for<0, 16>([&](auto m2) {
          const float w0 = win1024<m2>(s_o, c_o);
          const float w1 = win1024<m2 + 16>(s_o, c_o);
          const float y0a = v[m2].y * w0, y0b = v[m2 + 16].y * w1;   // frame t: first / second half
          const float y1a = v[m2].x * w0, y1b = v[m2 + 16].x * w1;   // frame t+1
          if (write) o0[32 * m2 + lane] = mytail[32 * m2 + lane] + y0a;
          if (pp.two && write) o0[H + 32 * m2 + lane] = y0b + y1a;
          const float nt = pp.two ? y1b : y0b;
          mytail[32 * m2 + lane] = nt;
          if (last && write) p.tail_out[(size_t)pp.s * H + 32 * m2 + lane] = nt;
        });
      }
    }
  }
  }   // g_begin < g_end
  if (kTmem) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "n"(256) : "memory");
  }
}

static int das_pick_warps(int M, int* ceff_in_smem, size_t* smem, int* tiles) {
  const size_t cap = 232448 - 1024;   // 227 KB dynamic limit, minus slack
  const size_t fixed = 1024 * sizeof(float2);
  const size_t ceff = (size_t)M * 1024 * sizeof(float2);
  // measured on B200 (C1, audio-s/s): 8x2 2.597 M, 8x1 2.605 M, 12x1 2.229 M (168 registers, spills), 16x1 with the accumulators
  // in tensor memory 2.373 M (128 registers): the kernel is bound by the issue rate of the butterflies, more warps do not help
  int want_w = 8, want_t = 1;
  if (const char* e = getenv("BF_DAS_CFG")) sscanf(e, "%dx%d", &want_w, &want_t);   // tuning knob: 8x2, 8x1, 12x1, 16x1 (16 warps: accumulators in tensor memory)
  if (!((want_w == 8 || want_w == 12 || want_w == 16) && (want_t == 1 || want_t == 2)) || (want_w > 8 && want_t == 2)) { want_w = 8; want_t = 1; }
  const size_t per_warp = (size_t)want_t * kTileF2 * sizeof(float2) + (size_t)want_t * sizeof(uint64_t) + 512 * sizeof(float);
  *tiles = want_t;
  *ceff_in_smem = (fixed + want_w * per_warp + ceff <= cap) ? 1 : 0;   // large arrays: weights stay in global memory (L2-resident)
  *smem = fixed + want_w * per_warp + (*ceff_in_smem ? ceff : 0);
  return want_w;
}

// true when the bulk-copy alignment rules hold (16-byte aligned base, strides multiples of 4 floats)
bool das_pairs_supported(const KernelParams& p) {
  if (getenv("BF_DAS_OLD")) return false;
  return p.H == 512 && (reinterpret_cast<uintptr_t>(p.in) & 15) == 0 && (p.in_stream_stride & 3) == 0 &&
         (p.in_mic_stride & 3) == 0;
}

template <int kWarps, int kTiles, bool kTmem>
static cudaError_t launch_das_pairs_t(const KernelParams& p, cudaStream_t st, int ctas, size_t smem, int in_smem) {
  cudaError_t e = cudaFuncSetAttribute(das_pairs_kernel<kWarps, kTiles, kTmem>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  das_pairs_kernel<kWarps, kTiles, kTmem><<<ctas, kWarps * 32, smem, st>>>(p, in_smem);
  return cudaGetLastError();
}

// one CTA per SM
cudaError_t launch_das_pairs(const KernelParams& p, cudaStream_t st, int sm_count) {
  int in_smem = 0, tiles = 1;
  size_t smem = 0;
  const int warps = das_pick_warps(p.M, &in_smem, &smem, &tiles);
  const int nh = p.hop_end - p.hop_begin;
  const long long total = (long long)p.n_streams * ((nh + 1) / 2);
  long long ctas = (total + warps - 1) / warps;
  if (ctas > sm_count) ctas = sm_count;
  if (ctas < 1) ctas = 1;
  if (warps == 16) return launch_das_pairs_t<16, 1, true>(p, st, (int)ctas, smem, in_smem);
  if (warps == 12) return launch_das_pairs_t<12, 1, false>(p, st, (int)ctas, smem, in_smem);
  if (tiles == 1) return launch_das_pairs_t<8, 1, false>(p, st, (int)ctas, smem, in_smem);
  return launch_das_pairs_t<8, 2, false>(p, st, (int)ctas, smem, in_smem);
}

}   // namespace bf
