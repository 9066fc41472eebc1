// Pipelined kernel for mvdr (BASELINE config C2), 1024-point frames, M <= 8, P <= 10, sm_100a.
//
//   reference path replaced (citations /root/reference/beamform/src/): util.h:217-253,289-314 (window, framing,
//   OLA) and apply_weights of mvdr.cpp:62-115, lcmv.cpp:88-140 (history matrices past_ffts of mvdr.cpp:228-243).
//
// One CTA (16 warps, one per SM) owns one stream.  The work of a frame pair (t, t+1) is feed-forward (spectra ->
// gate -> per-bin solves -> inverse), so the warps are specialised and coupled only by mbarriers; consecutive pairs
// overlap and no CTA-wide barrier exists after start-up:
//   warps 0-7   microphone m: hops t-1..t+1 arrive by TMA bulk copy (issued one pair ahead), window, packed 1024-point
//               FFT in registers, even/odd separation into X_t, X_{t+1} by warp shuffles, magnitudes for the gate,
//               history append, staging of the selected bins' histories for the solvers
//   warps 8-14  solvers: a batch of 16 selected bins per warp, lanes 0-15 frame t, lanes 16-31 frame t+1; the P-1
//               history frames the two frames share are summed once (half per lane + one shfl.xor), Cholesky, weights
//   warp 15     Hermitian assembly, inverse FFT, synthesis window, overlap-add, output
// The history of every in-band bin (mvdr.cpp:99-101: P frames x M microphones, 217 KB per stream) lives in TENSOR
// MEMORY: lane = 32*(m%4) + bin%32 (the lane quarter a warp may address is warp%4, and microphone warp m holds bins
// == lane mod 32 after its FFT), column = (m/4)*242 + ((bin/32)*11 + frame%(P+1))*2 + {re, im}.  Every cell is read
// and written by one warp only, so tcgen05.st / tcgen05.ld need no cross-warp ordering; the ring never touches L2 or
// HBM except for one load at the start and one save at the end of a launch (state carried between calls, same
// layout as sel_pairs_kernel's global ring).  Shapes outside (in-band bins >= 352, P > 10) run sel_pairs_kernel.
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "async_copy.cuh"
#include "bf_device.h"
#include "fft_reg.cuh"
#include "warp_fft1024.cuh"
#include "phase_b_select.cuh"
#include "sel_solve.cuh"
#include "tmem.cuh"

namespace bf {

constexpr int kSsMicWarps = 8;
// solver warps == staging batches in flight (warp w owns batch slot w): 16 warps x 128 registers.  (lcmv solves in FP64: measured on
// B200 it is slower here than in sel_pairs_kernel -- 104 k audio-s/s with 7 solvers at 128 registers (spills), 86 k with 3 solvers at 168
// registers, against 123 k -- so capi.cu keeps lcmv on sel_pairs_kernel; the template stays generic.)
template <int ALGO> struct SsCfg { static constexpr int kSolvers = 7; static constexpr int kThreads = (kSsMicWarps + kSolvers + 1) * 32; };
constexpr int kSsMaxSolvers = 7;
constexpr int kSsBlocks = 11;                            // 32-bin blocks with on-chip history: logical bins 0..351
constexpr int kSsBins = kSsBlocks * 32;
constexpr int kSsSlots = 11;                             // ring slots per bin in tensor memory (P + 1 <= 11)
constexpr int kSsHalfCols = kSsBlocks * kSsSlots * 2;    // 242 columns per microphone of a lane quarter
constexpr int kSsItems = 16;                             // bins per solver batch
constexpr int kSsEntries = kSsSlots + 1;                 // staged per item: the 11 ring slots + X_{t+1}
constexpr int kSsItemF2 = kSsEntries * 8 + 2;            // 98 float2 = 784 B: 16-byte multiple, odd multiple of 16 B (LDS.128 conflict-free)
constexpr int kSsKY = 3;                                 // output-spectrum buffers in flight
constexpr int kSsYDoneCount = kSsBins / kSsItems + 1;    // arrivals that complete y_done: every batch of the pair (<= 22) + microphone warp 0 with the remainder

struct SsBatch {
  float2 item[kSsItems][kSsItemF2];   // [item][entry*8 + mic]
  unsigned short bin[kSsItems];
  unsigned char flags[kSsItems];      // bit 0: frame t selected, bit 1: frame t+1 selected
  int n_items, ybuf, t, two;
};

struct SsShared {
  float2 tw[1024];
  float2 tile[kSsMicWarps + 1][1024]; // per microphone warp: TMA landing zone, then FFT exchange tile; [8]: the inverse warp's exchange tile
  float mags[8][2][kSsBins];          // |X_m[l]| of both frames; after the gate microphone m parks X_{t+1} in its own rows (352 float2)
  float ygain[kSsBins];               // default output = ygain * X_0: 0.01 inside the band (mvdr.cpp:96), 1 for mvdr's bin 0, else 0
  float tail[512];                    // overlap-add tail (inverse warp only)
  float2 y[kSsKY][2][kL1K];
  SsBatch batch[kSsMaxSolvers];
  float sqrtE[2][8];
  unsigned masks[2][kSsBlocks + 1];
  int nonfinite[kSsKY][2];
  int total_batches;
  short sel_slot[kSsBins];
  unsigned char inband[kSsBins];
  uint64_t tma_bar[kSsMicWarps], empty[kSsMaxSolvers], y_done[kSsKY], y_free[kSsKY];
  uint32_t tmem_slot;
};

// try_wait with a suspend-time hint: the warp is parked by the hardware until the phase completes or `ns` nanoseconds have
// passed (no polling instructions in between: an idle solver must not eat the issue slots of the working warps)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity, uint32_t ns = 2000u) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// every wait of this kernel is bounded: a protocol error traps (the launch fails) instead of hanging the GPU
__device__ __forceinline__ void ss_watchdog(unsigned& spins, long long& t0) {
  if ((++spins & 255u) == 0) {
    const long long now = clock64();
    if (t0 == 0) t0 = now;
    else if (now - t0 > (1ll << 35)) __trap();   // ~17 s at 1.9 GHz without progress
  }
}
__device__ __forceinline__ void ss_wait(uint64_t* bar, uint32_t parity) {
  unsigned spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) ss_watchdog(spins, t0);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, uint32_t cnt) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(cnt) : "memory");
}
// Solver wake-up goes through hardware named barriers (a blocked warp issues nothing, unlike an mbarrier poll loop): the
// eight microphone warps bar.arrive on barrier 3 + slot once the batch in that slot is complete, the slot's solver bar.syncs.
constexpr int kSsBarBase = 3, kSsBarThreads = (kSsMicWarps + 1) * 32;
__device__ __forceinline__ void ss_publish(int slot) {
  __syncwarp();
  // bar.arrive orders this thread's earlier shared-memory writes before the barrier completes (no MEMBAR needed: it cost
  // ~110 cycles per call on the microphone warps' critical path)
  asm volatile("bar.arrive %0, %1;" ::"r"(kSsBarBase + slot), "r"(kSsBarThreads) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }


__device__ __noinline__ bool ss_gate_fp64(const KernelParams& p, int s, int t, int l, int f, int lane) { return gate_fp64(p, s, t, l, f, lane); }

// One solver batch: lanes 0-15 carry frame t of items 0-15, lanes 16-31 frame t+1.  Staged entries are ring slots
// sigma = frame % Dt (Dt = P + 1): slot sx holds X_t, slot (sx + 1) % Dt holds X_{t-P}; the other P-1 slots are the frames
// both histories contain (mvdr.cpp:87, :239-243: R_t over t-P..t-1, R_{t+1} over t-P+1..t).
// gss (gss.cpp:118-137): one lane per staged bin runs the frames of the pair in order (the separation matrix W of the bin is a
// recursion over the selected frames; it lives in global memory, L2-resident, bin index fastest).  Staged entries 0 / 1 = X_t / X_{t+1}.
__device__ __noinline__ void ss_solve_batch_gss(const KernelParams& p, SsShared& sh, const SsBatch& bt, int lane, int s) {
  if (lane >= bt.n_items || lane >= kSsItems) return;
  const int l = bt.bin[lane];
  float2* Wg = p.gss_w + (size_t)s * BF_GSS_ROWS * p.M * p.Lsel + sh.sel_slot[l];
  const float2* steer_l = p.steer + (size_t)l * p.C * p.M;
  const float4* it = reinterpret_cast<const float4*>(&bt.item[lane][0]);
  for (int f = 0; f < 2; f++) {
    if (!((bt.flags[lane] >> f) & 1)) continue;
    float2 x[8];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const float4 r = it[f * 4 + q];
      x[2 * q] = make_float2(r.x, r.y);
      x[2 * q + 1] = make_float2(r.z, r.w);
    }
    const float2 yv = gss_item(p, Wg, (size_t)p.Lsel, x, steer_l);
    sh.y[bt.ybuf][f][l] = yv;
    if (!(isfinite(yv.x) && isfinite(yv.y))) sh.nonfinite[bt.ybuf][f] = 1;
  }
}

template <int ALGO>
__device__ __noinline__ void ss_solve_batch(const KernelParams& p, SsShared& sh, const SsBatch& bt, int lane) {
  typedef typename std::conditional<ALGO == ALGO_MVDR, float, double>::type T;
  const int i = lane & 15, f = lane >> 4;
  const int Dt = p.P + 1;
  const bool on = i < bt.n_items;
  const int l = on ? (int)bt.bin[i] : 0;
  const bool sel = on && ((bt.flags[i] >> f) & 1) != 0;
  const int sx = (p.frame_index0 + (bt.t - p.hop_begin)) % Dt;
  const float4* it = reinterpret_cast<const float4*>(&bt.item[i][0]);
  auto load_entry = [&](int e, float2 (&h)[8]) {
    const float4* src = it + e * 4;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const float4 r = src[q];
      h[2 * q] = make_float2(r.x, r.y);
      h[2 * q + 1] = make_float2(r.z, r.w);
    }
  };
  HermLower<8, T> A;
#pragma unroll
  for (int k = 0; k < 8; k++) A.dg[k] = T(0);
#pragma unroll
  for (int k = 0; k < 28; k++) A.lo[k] = mk<T>(T(0), T(0));
  const int n_sh = p.P - 1, n_half = (n_sh + 1) >> 1;
  int sg = sx + 2 + (f ? n_half : 0);
  if (sg >= Dt) sg -= Dt;
  if (sg >= Dt) sg -= Dt;
  const int sold = (sx + 1 == Dt) ? 0 : sx + 1;
#pragma unroll 1
  for (int j = 0; j <= n_half; j++) {   // n_half shared frames per lane, then (after the butterfly) the one frame only this lane's history holds
    bool go;
    int e;
    if (j < n_half) {
      go = f == 0 || n_half + j < n_sh;
      e = sg;
      if (++sg == Dt) sg = 0;
    } else {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; k++) A.dg[k] += __shfl_xor_sync(0xffffffffu, A.dg[k], 16);
#pragma unroll
      for (int k = 0; k < 28; k++) {
        A.lo[k].x += __shfl_xor_sync(0xffffffffu, A.lo[k].x, 16);
        A.lo[k].y += __shfl_xor_sync(0xffffffffu, A.lo[k].y, 16);
      }
      go = sel;
      e = f ? sx : sold;   // X_t for frame t+1, X_{t-P} for frame t
    }
    if (go) {
      float2 hf[8];
      load_entry(e, hf);
      cov_rank1<8, T>(A, hf);
    }
  }
  if (sel) {
    float2 x[8];
    load_entry(f ? kSsSlots : sx, x);    // this lane's own frame
    T invd[8];
    chol_in_place<8, T>(p, A, invd);
    const float2* steer_l = p.steer + (size_t)l * p.C * p.M;
    const float2 yv = (ALGO == ALGO_MVDR) ? mvdr_finish<8, T>(p, A, invd, x, steer_l) : lcmv_finish<8, T>(p, A, invd, x, steer_l);
    sh.y[bt.ybuf][f][l] = yv;
    if (!(isfinite(yv.x) && isfinite(yv.y))) sh.nonfinite[bt.ybuf][f] = 1;   // cold start (SURVEY B-10): the whole frame turns NaN
  }
}

template <int ALGO>
__global__ void __launch_bounds__(SsCfg<ALGO>::kThreads, 1) sel_stream_kernel(const __grid_constant__ KernelParams p, const int use_tma) {
  constexpr int kSsSolvers = SsCfg<ALGO>::kSolvers, kSsThreads = SsCfg<ALGO>::kThreads;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SsShared& sh = *reinterpret_cast<SsShared*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = p.M, D = p.ring_depth, Dt = p.P + 1;
  constexpr int H = 512;
  // persistent CTAs: CTA c owns streams c, c + gridDim.x, ...; the pipeline keeps running across stream boundaries (no drain)
  const int s_first = blockIdx.x + p.stream_begin, s_end = p.stream_begin + p.n_streams;
  const int nh = p.hop_end - p.hop_begin;
  const int npairs = (nh + 1) >> 1;
  const int sig0 = p.frame_index0 % Dt;   // tensor-memory ring slot of the launch's first frame

  for (int i = tid; i < 1024; i += kSsThreads) {
    const int k1 = i >> 5, l = i & 31;
    float sn, cs;
    sincospif(-2.0f * (float)((k1 * l) & 1023) / 1024.0f, &sn, &cs);
    sh.tw[i] = make_float2(cs, sn);
  }
  for (int i = tid; i < kSsBins; i += kSsThreads) {
    const bool inb = p.inband[i] != 0 && !(ALGO == ALGO_MVDR && i == 0);
    sh.sel_slot[i] = (short)p.sel_slot[i];
    sh.inband[i] = inb ? 1 : 0;
    sh.ygain[i] = inb ? 0.01f : ((ALGO == ALGO_MVDR && i == 0) ? 1.0f : 0.0f);
  }
  for (int i = tid; i < kSsKY * 2 * kL1K; i += kSsThreads) (&sh.y[0][0][0])[i] = make_float2(0.f, 0.f);   // bins outside the band stay 0 (mvdr.cpp:103)
  if (tid < kSsKY) { sh.nonfinite[tid][0] = 0; sh.nonfinite[tid][1] = 0; }
  if (tid == 0) {
    sh.total_batches = -1;
    for (int i = 0; i < kSsMicWarps; i++) mbar_init(&sh.tma_bar[i], 1);
    for (int i = 0; i < kSsSolvers; i++) mbar_init(&sh.empty[i], 1);
    for (int i = 0; i < kSsKY; i++) { mbar_init(&sh.y_done[i], kSsYDoneCount); mbar_init(&sh.y_free[i], 1); }
  }
  mbar_fence_init();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = sh.tmem_slot;

  const bool is_mic = warp < kSsMicWarps, is_inv = warp == kSsMicWarps + kSsSolvers;
  if (is_mic || is_inv) {
    // ================================================================== transform warps: 8 microphones + the inverse warp.
    // They share ONE copy of the FFT code (IFFT(G) = swap(FFT(swap(G)))): the instruction caches hold the hot loops of
    // three roles at once, and a second inlined transform is 9 KB of them.
    const int m = is_mic ? warp : 0;
    const bool have = is_mic && m < M;
    const uint32_t tm = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) & 1) * kSsHalfCols);
    float2* tile = sh.tile[is_mic ? warp : kSsMicWarps];
    float* stage = reinterpret_cast<float*>(tile);
    float2* xpark = reinterpret_cast<float2*>(&sh.mags[m][0][0]);   // X_{t+1} of this microphone between the gate and the staging
    double sd, cd;
    sincospi((double)lane / 1024.0, &sd, &cd);
    // analysis window * 0.5 (microphones) / synthesis window * out_amp / N (inverse)
    const float s_w = is_mic ? (float)(0.5 * sd) : (float)(sd * p.out_scale), c_w = is_mic ? (float)(0.5 * cd) : (float)(cd * p.out_scale);

    auto issue = [&](int sx, int t) {   // hops t-1..t+1 of this microphone of stream sx -> tile (hop -1 = per-stream state, util.h:275-277)
      const float* base = p.in + (size_t)sx * p.in_stream_stride + (size_t)m * p.in_mic_stride;
      const bool two = t + 1 < p.hop_end;
      const uint32_t nb = (two ? 2u : 1u) * H * 4u;
      mbar_expect_tx(&sh.tma_bar[m], nb + H * 4u);
      const float* prev = (t - 1 < 0) ? p.prev_hop + ((size_t)sx * M + m) * H : base + (size_t)(t - 1) * H;
      bulk_g2s(stage, prev, H * 4u, &sh.tma_bar[m]);
      bulk_g2s(stage + H, base + (size_t)t * H, nb, &sh.tma_bar[m]);
    };
    if (have && use_tma && npairs > 0 && lane == 0 && s_first < s_end) issue(s_first, p.hop_begin);

    int seq0 = 0;   // staging batches published so far (identical in every microphone warp)
    int gp0 = 0;    // frame pairs this CTA has been through before the current stream
#pragma unroll 1
    for (int s = s_first; s < s_end; s += gridDim.x, gp0 += npairs) {
    const float* in_s = p.in + (size_t)s * p.in_stream_stride + (size_t)m * p.in_mic_stride;
    float2* hist_s = p.hist + (size_t)s * D * M * p.Lsel + (size_t)m * p.Lsel;   // + slot*M*Lsel + sel_slot[l]
    if (is_inv)
      for (int i = lane; i < H; i += 32) sh.tail[i] = p.tail[(size_t)s * H + i];
    if (is_mic && ALGO != ALGO_GSS) {
      // ---- history of the previous calls: global ring (slot = frame % D) -> tensor memory (slot = frame % Dt) ----
#pragma unroll 1
      for (int j = 1; j <= p.P; j++) {
        int g = p.ring_slot0 - j;
        if (g < 0) g += D;
        int sg = sig0 - j;
        if (sg < 0) sg += Dt;
        const float2* src = hist_s + (size_t)g * M * p.Lsel;
        float2 hv[kSsBlocks];
#pragma unroll
        for (int k2 = 0; k2 < kSsBlocks; k2++) {
          const int ss = sh.sel_slot[k2 * 32 + lane];
          hv[k2] = (have && ss >= 0) ? src[ss] : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int k2 = 0; k2 < kSsBlocks; k2++) tmem_st2(tm + (uint32_t)((k2 * kSsSlots + sg) * 2), hv[k2].x, hv[k2].y);
      }
      tmem_wait_st();
    }

#pragma unroll 1
    for (int ip = 0; ip < npairs; ip++) {
      const int gp = gp0 + ip;   // pair counter of the CTA: buffer rotation and barrier parities run across streams
      const int t = p.hop_begin + 2 * ip;
      const bool two = t + 1 < p.hop_end;
      const int ky = gp % kSsKY;
      float2 v[32];
      float e0 = 0.f, e1 = 0.f;
      bool z0 = false, z1 = false;
      if (is_mic) {
        // ---------------------------------------------------------------- hops -> windowed packed frame pair
        if (have) {
          if (use_tma) {
            ss_wait(&sh.tma_bar[m], gp & 1);
          } else {   // unaligned caller buffers: plain warp copy, no prefetch
            const float* prev = (t - 1 < 0) ? p.prev_hop + ((size_t)s * M + m) * H : in_s + (size_t)(t - 1) * H;
            for (int i = lane; i < H; i += 32) {
              stage[i] = prev[i];
              stage[H + i] = in_s[(size_t)t * H + i];
              stage[2 * H + i] = two ? in_s[(size_t)(t + 1) * H + i] : 0.f;
            }
            __syncwarp();
          }
          static_for<0, 16>([&](auto r) {
            const float a = stage[32 * r + lane], bb = stage[512 + 32 * r + lane];
            const float c = two ? stage[1024 + 32 * r + lane] : 0.0f;
            const float w0 = win1024<r>(s_w, c_w);
            const float w1 = win1024<r + 16>(s_w, c_w);
            v[brev5(r)] = make_float2(a * w0, bb * w0);
            v[brev5(r + 16)] = make_float2(bb * w1, c * w1);
          });
        } else {
#pragma unroll
          for (int r = 0; r < 32; r++) v[r] = make_float2(0.f, 0.f);
        }
        __syncwarp();   // staged samples consumed: the tile becomes the exchange buffer
        // sqrt of the windowed frame energies: scale of the FP32 FFT's absolute error (gate guard band)
#pragma unroll
        for (int r = 0; r < 32; r++) { e0 = fmaf(v[r].x, v[r].x, e0); e1 = fmaf(v[r].y, v[r].y, e1); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { e0 += __shfl_xor_sync(0xffffffffu, e0, o); e1 += __shfl_xor_sync(0xffffffffu, e1, o); }
        if (lane == 0) { sh.sqrtE[0][m] = 2.0f * sqrtf(e0); sh.sqrtE[1][m] = 2.0f * sqrtf(e1); }
      } else {
        // ---------------------------------------------------------------- output spectra of the pair -> G = Yh_t + i Yh_{t+1}
        ss_wait(&sh.y_done[ky], (uint32_t)(gp / kSsKY) & 1u);   // defaults written and every batch of the pair solved
        // one inf/NaN bin makes the reference's whole inverse frame NaN; the two frames of a pair share one complex transform
        // here, so a poisoned frame is left out of G and re-poisoned at the output without touching its partner
        z0 = *reinterpret_cast<volatile int*>(&sh.nonfinite[ky][0]) != 0;
        z1 = *reinterpret_cast<volatile int*>(&sh.nonfinite[ky][1]) != 0;
        if (z0 || z1 || !two) {   // rare: drop the frame from the shared transform (every in-band bin is rewritten for the next pair anyway)
          for (int l = lane; l < kL1K; l += 32) {
            if (z0) sh.y[ky][0][l] = make_float2(0.f, 0.f);
            if (z1 || !two) sh.y[ky][1][l] = make_float2(0.f, 0.f);
          }
          __syncwarp();
        }
        static_for<0, 32>([&](auto n1c) {
          constexpr int n1 = decltype(n1c)::value;
          const int j = n1 * 32 + lane;
          const bool mir = (n1 > 16) || (n1 == 16 && lane != 0);   // bins above N/2: conjugate of bin N - j
          const int l = mir ? 1024 - j : j;
          float2 y0 = sh.y[ky][0][l], y1 = sh.y[ky][1][l];
          if constexpr (n1 == 0 || n1 == 16) {
            if (lane == 0) { y0.y = 0.f; y1.y = 0.f; }   // Re(): self-conjugate bins 0 and N/2
          }
          if constexpr (n1 == 15 || n1 == 16) {           // Hermitian part of the pair (N/2-1, N/2+1); the pseudo-bin is 0 on this path
            if ((n1 == 15 && lane == 31) || (n1 == 16 && lane == 1)) { y0.x *= 0.5f; y0.y *= 0.5f; y1.x *= 0.5f; y1.y *= 0.5f; }
          }
          // G = Yh_t + i Yh_{t+1} (j <= N/2), conj(Yh_t) + i conj(Yh_{t+1}) (mirror); parts swapped: IFFT(G) = swap(FFT(swap(G)))
          const float2 g = mir ? make_float2(y0.x + y1.y, y1.x - y0.y) : make_float2(y0.x - y1.y, y0.y + y1.x);
          v[brev5(n1)] = make_float2(g.y, g.x);
        });
        __syncwarp();
        if (lane == 0) {   // the spectrum buffer goes back to microphone warp 0
          sh.nonfinite[ky][0] = 0; sh.nonfinite[ky][1] = 0;
          __threadfence_block();
          mbar_arrive(&sh.y_free[ky]);
        }
      }
      // ------------------------------------------------------------------ the transform (one code copy for both roles)
      warp_fft1024_fwd(v, tile, sh.tw, lane, [&]() {
        // the exchange tile is free until the next pair: its hops start to arrive now, behind the second FFT pass,
        // the gate and the staging of this pair
        if (have && use_tma && lane == 0 && (ip + 1 < npairs || s + (int)gridDim.x < s_end)) {
          fence_proxy_async();
          if (ip + 1 < npairs) issue(s, t + 2); else issue(s + (int)gridDim.x, p.hop_begin);   // ... or the first pair of the CTA's next stream
        }
      });
      if (is_inv) {
        // ---------------------------------------------------------------- synthesis window, overlap-add (util.h:244-253,301-302)
        float* o0 = p.out + (size_t)s * p.out_stream_stride + (size_t)t * H;
        const float bad0 = z0 ? __int_as_float(0x7fc00000) : 0.f;
        const float bad1 = z1 ? __int_as_float(0x7fc00000) : 0.f;
        static_for<0, 16>([&](auto m2c) {
          constexpr int m2 = decltype(m2c)::value;
          const float w0 = win1024<m2>(s_w, c_w);
          const float w1 = win1024<m2 + 16>(s_w, c_w);
          const float y0a = v[m2].y * w0 + bad0, y0b = v[m2 + 16].y * w1 + bad0;   // frame t: first / second half
          const float y1a = v[m2].x * w0 + bad1, y1b = v[m2 + 16].x * w1 + bad1;   // frame t+1
          o0[32 * m2 + lane] = sh.tail[32 * m2 + lane] + y0a;
          if (two) o0[H + 32 * m2 + lane] = y0b + y1a;
          sh.tail[32 * m2 + lane] = two ? y1b : y0b;
        });
        continue;
      }
      // ================================================================== microphone warps only from here
      int sx = sig0 + 2 * ip;
      sx %= Dt;                                     // ring slot of frame t
      int sx1 = (sx + 1 == Dt) ? 0 : sx + 1;        // ring slot of frame t+1 (it still holds frame t-P)
      if (ALGO == ALGO_GSS) { sx = 0; sx1 = 1; }    // gss keeps no history: slots 0 / 1 just park X_t / X_{t+1} until the staging
      // v[k2] = Z[32*k2 + lane], Z = FFT(0.5*w*(x_t + i x_{t+1})).  X_t[l] = Z[l] + conj(Z[N-l]), X_{t+1}[l] = -i (Z[l] - conj(Z[N-l]));
      // Z[N-l] sits in lane (32 - lane) % 32, register 31 - k2 (lane 0: its own register (32 - k2) % 32).
      if (m == 0) {
        if (gp >= kSsKY) ss_wait(&sh.y_free[ky], (uint32_t)((gp / kSsKY) - 1) & 1u);
        // a non-finite input sample of microphone 0 makes every default output of the frame non-finite (SURVEY B-10)
        if (!isfinite(e0)) sh.nonfinite[ky][0] = 1;
        if (two && !isfinite(e1)) sh.nonfinite[ky][1] = 1;
      }
      float2 x1[kSsBlocks];
      {
        const int src_lane = (32 - lane) & 31;
        static_for<0, kSsBlocks>([&](auto k2c) {
          constexpr int k2 = decltype(k2c)::value;
          const int l = k2 * 32 + lane;
          const float2 a = v[k2];
          float2 b;
          b.x = __shfl_sync(0xffffffffu, v[31 - k2].x, src_lane);
          b.y = __shfl_sync(0xffffffffu, v[31 - k2].y, src_lane);
          if (lane == 0) b = v[(32 - k2) & 31];
          const float2 x0 = make_float2(a.x + b.x, a.y - b.y);
          const float2 xb = make_float2(a.y + b.y, b.x - a.x);
          x1[k2] = xb;
          sh.mags[m][0][l] = sqrt_approx(fmaf(x0.x, x0.x, x0.y * x0.y));
          sh.mags[m][1][l] = sqrt_approx(fmaf(xb.x, xb.x, xb.y * xb.y));
          tmem_st2(tm + (uint32_t)((k2 * kSsSlots + sx) * 2), x0.x, x0.y);   // history append of frame t (mvdr.cpp:99-101)
          if (m == 0) {   // default outputs (overwritten by the solvers where a bin is selected)
            const float g = sh.ygain[l];
            sh.y[ky][0][l] = make_float2(g * x0.x, g * x0.y);
            sh.y[ky][1][l] = make_float2(g * xb.x, g * xb.y);
          }
        });
      }
      named_bar_sync(1, kSsMicWarps * 32);   // magnitudes of all microphones complete
      // ---------------------------------------------------------------- gate (mvdr.cpp:79-85), FP64 re-decision inside the guard band
      {
        float es0 = 0.f, es1 = 0.f;
        for (int ch = 0; ch < M; ch++) { es0 += sh.sqrtE[0][ch]; es1 += sh.sqrtE[1][ch]; }
        const float g0 = 2.0e-5f * es0 + 1.0e-6f * p.thr_mag, g1 = 2.0e-5f * es1 + 1.0e-6f * p.thr_mag;
#pragma unroll 1
        for (int k2 = m; k2 < kSsBlocks; k2 += kSsMicWarps) {
          const int l = k2 * 32 + lane;
          const bool inb = sh.inband[l] != 0;
          float st0 = 0.f, st1 = 0.f;
#pragma unroll
          for (int ch = 0; ch < 8; ch++)
            if (ch < M) { st0 += sh.mags[ch][0][l]; st1 += sh.mags[ch][1][l]; }
          bool f0 = false, f1 = false, r0 = false, r1 = false;
          if (inb) {
            if (fabsf(st0 - p.thr_mag) <= g0) r0 = true;
            else if (st0 > p.thr_mag) f0 = true;
            if (two) {
              if (fabsf(st1 - p.thr_mag) <= g1) r1 = true;
              else if (st1 > p.thr_mag) f1 = true;
            }
          }
          unsigned rm = __ballot_sync(0xffffffffu, r0) | (__ballot_sync(0xffffffffu, r1) << 0) * 0u;
          rm = __ballot_sync(0xffffffffu, r0);
          unsigned rm1 = __ballot_sync(0xffffffffu, r1);
#pragma unroll 1
          while (rm | rm1) {   // rare: exact double DFT of that bin for every microphone, warp-cooperative
            const int f = rm ? 0 : 1;
            unsigned& mk = rm ? rm : rm1;
            const int b = __ffs(mk) - 1;
            mk &= mk - 1;
            const bool sel = ss_gate_fp64(p, s, t, k2 * 32 + b, f, lane);
            if (lane == b) { if (f) f1 = sel; else f0 = sel; }
          }
          const unsigned m0 = __ballot_sync(0xffffffffu, f0), m1 = __ballot_sync(0xffffffffu, f1);
          if (lane == 0) { sh.masks[0][k2] = m0; sh.masks[1][k2] = m1; }
        }
      }
      named_bar_sync(2, kSsMicWarps * 32);   // selection masks complete; nobody reads the magnitudes any more
      if (p.capture) {   // diagnostics: one byte per FFT bin and frame, bit 0 = selected
        for (int k2 = m; k2 <= 16; k2 += kSsMicWarps) {
          const int l = k2 * 32 + lane;
          if (l > 512) continue;
          for (int f = 0; f < (two ? 2 : 1); f++) {
            unsigned char* cap = p.capture + (size_t)s * p.capture_stream_stride + (size_t)(t + f) * p.N;
            const unsigned char fl = (k2 < kSsBlocks) ? (unsigned char)((sh.masks[f][k2] >> lane) & 1u) : (unsigned char)0;
            cap[l] = fl;
            if (l > 0 && l < 511) cap[1024 - l] = fl;
            if (l == 511) cap[513] = 0;   // the pseudo-bin is outside the band on this path
          }
        }
      }
      // ---------------------------------------------------------------- staging of the selected bins for the solvers
      unsigned emask = 0;   // lane k2: bins of block k2 selected in either frame
      if (lane < kSsBlocks) emask = sh.masks[0][lane] | sh.masks[1][lane];
      const int cnt = __popc(emask);
      int pre = cnt;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const int nbr = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += nbr;
      }
      const int n_items = __shfl_sync(0xffffffffu, pre, kSsBlocks - 1);
      const int base_excl = pre - cnt;
      const int nbat = (n_items + kSsItems - 1) / kSsItems;
      // defaults written, batch count known: this arrival stands for every batch the pair does NOT have (the solvers arrive once per batch)
      if (m == 0 && lane == 0) mbar_arrive_cnt(&sh.y_done[ky], (uint32_t)(kSsYDoneCount - nbat));
      int cur_q = -1;
      // Blocks without a selected bin (most of them) only append X_{t+1} to the history, straight from the registers; the
      // blocks with selected bins park X_{t+1} in this microphone's (now idle) magnitude rows and go through the rolled
      // staging loop below, which reads the history BEFORE the append overwrites the slot of frame t-P.
      unsigned bmask = __ballot_sync(0xffffffffu, emask != 0u);   // bit k2: block k2 has a bin selected in either frame
      static_for<0, kSsBlocks>([&](auto k2c) {
        constexpr int k2 = decltype(k2c)::value;
        if ((bmask >> k2) & 1u) xpark[k2 * 32 + lane] = x1[k2];
        else if (two && ALGO != ALGO_GSS) tmem_st2(tm + (uint32_t)((k2 * kSsSlots + sx1) * 2), x1[k2].x, x1[k2].y);
      });
      __syncwarp();
      tmem_wait_st();   // the appends of frame t (issued before the gate) must have landed before the staging reads them back
#pragma unroll 1
      while (bmask) {
        const int k2 = __ffs(bmask) - 1;
        bmask &= bmask - 1;
        const unsigned em = __shfl_sync(0xffffffffu, emask, k2);
        const float2 xn = xpark[k2 * 32 + lane];
        {
          const int bbase = __shfl_sync(0xffffffffu, base_excl, k2);
          uint32_t r[24];
          tmem_ld16(tm + (uint32_t)(k2 * kSsSlots * 2), r);
          tmem_ld8(tm + (uint32_t)(k2 * kSsSlots * 2 + 16), r + 16);
          tmem_wait_ld();
          const bool mine = ((em >> lane) & 1u) != 0;
          const int idx = bbase + __popc(em & ((1u << lane) - 1u));
          const int q_lo = bbase / kSsItems, q_hi = (bbase + __popc(em) - 1) / kSsItems;
#pragma unroll 1
          for (int q = q_lo; q <= q_hi; q++) {
            const int n = seq0 + q, bslot = n % kSsSolvers;
            SsBatch& bt = sh.batch[bslot];
            if (q > cur_q) {
              if (cur_q >= 0) ss_publish((seq0 + cur_q) % kSsSolvers);   // the previous batch is complete for this warp
              const int round = n / kSsSolvers;
              if (round >= 1) ss_wait(&sh.empty[bslot], (uint32_t)(round - 1) & 1u);
              if (m == 0 && lane == 0) {
                bt.n_items = min(kSsItems, n_items - q * kSsItems);
                bt.ybuf = ky; bt.t = t; bt.two = two ? 1 : 0;
              }
              cur_q = q;
            }
            if (mine && idx / kSsItems == q) {
              float2* dst = &bt.item[idx % kSsItems][m];
              if (ALGO == ALGO_GSS) {
                dst[0] = make_float2(__uint_as_float(r[0]), __uint_as_float(r[1]));   // X_t (slot 0)
                dst[8] = xn;                                                              // X_{t+1}
              } else {
#pragma unroll
                for (int e = 0; e < kSsSlots; e++) dst[e * 8] = make_float2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
                dst[kSsSlots * 8] = xn;
              }
              if (m == 0) {
                bt.bin[idx % kSsItems] = (unsigned short)(k2 * 32 + lane);
                bt.flags[idx % kSsItems] = (unsigned char)(((sh.masks[0][k2] >> lane) & 1u) | (((sh.masks[1][k2] >> lane) & 1u) << 1));
              }
            }
          }
        }
        if (two && ALGO != ALGO_GSS) tmem_st2(tm + (uint32_t)((k2 * kSsSlots + sx1) * 2), xn.x, xn.y);   // history append of frame t+1
      }
      if (cur_q >= 0) ss_publish((seq0 + cur_q) % kSsSolvers);
      seq0 += nbat;
      tmem_wait_st();
    }
    // ---- state for the next call: the last min(nh, P) frames go back to the global ring; the overlap-add tail ----
    if (have && ALGO != ALGO_GSS) {
      const int nsave = min(nh, p.P);
      const int a0 = (sig0 + nh - 1) % Dt;           // ring slot (tensor memory) of the launch's last frame
      const int g0 = (p.ring_slot0 + nh - 1) % D;    // its slot in the global ring
      const size_t slot_stride = (size_t)M * p.Lsel;
#pragma unroll 1
      for (int k2 = 0; k2 < kSsBlocks; k2++) {
        uint32_t r[24];
        tmem_ld16(tm + (uint32_t)(k2 * kSsSlots * 2), r);
        tmem_ld8(tm + (uint32_t)(k2 * kSsSlots * 2 + 16), r + 16);
        tmem_wait_ld();
        const int ss = sh.sel_slot[k2 * 32 + lane];
#pragma unroll
        for (int e = 0; e < kSsSlots; e++) {
          int age = a0 - e;                          // slot e holds frame (nh - 1 - age) of this launch
          if (age < 0) age += Dt;
          int g = g0 - age;
          if (g < 0) g += D;
          if (e < Dt && age < nsave && ss >= 0)
            hist_s[(size_t)g * slot_stride + ss] = make_float2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
        }
      }
    }
    if (is_inv) {
      __syncwarp();
      for (int i = lane; i < H; i += 32) p.tail[(size_t)s * H + i] = sh.tail[i];
      __syncwarp();
    }
    }   // streams of this CTA
    if (is_mic) {
      // end of the CTA's last stream: one more arrival per slot releases its solver, which leaves when its next batch number reaches the total
      if (m == 0 && lane == 0) *reinterpret_cast<volatile int*>(&sh.total_batches) = seq0;
      for (int w = 0; w < kSsSolvers; w++) {
        const int rounds = seq0 > w ? (seq0 - w + kSsSolvers - 1) / kSsSolvers : 0;   // batches that went through slot w
        if (rounds >= 1) ss_wait(&sh.empty[w], (uint32_t)(rounds - 1) & 1u);
        ss_publish(w);
      }
    }
  } else {
    // ================================================================== solver warps
    const int w = warp - kSsMicWarps;
#pragma unroll 1
    for (int round = 0;; round++) {
      const int n = w + round * kSsSolvers;
      named_bar_sync(kSsBarBase + w, kSsBarThreads);   // blocked in hardware until the eight microphone warps have filled slot w
      const int tot = *reinterpret_cast<volatile int*>(&sh.total_batches);
      if (tot >= 0 && n >= tot) break;
      const SsBatch& bt = sh.batch[w];
      const int ky = bt.ybuf;
      if constexpr (ALGO == ALGO_GSS) ss_solve_batch_gss(p, sh, bt, lane, p.stream_begin + (int)blockIdx.x);
      else ss_solve_batch<ALGO>(p, sh, bt, lane);
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        mbar_arrive(&sh.empty[w]);
        mbar_arrive(&sh.y_done[ky]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

size_t sel_stream_smem() { return sizeof(SsShared) + 128; }

// The on-chip ring holds logical bins 0..351 with P + 1 <= 11 slots; the band must end below bin 352 (launch values:
// 100-16 000 Hz -> bins 3..341) and the pseudo-bin must be outside it.
bool sel_stream_supported(const KernelParams& p, int algo, const uint8_t* inband_host) {
  if (getenv("BF_SEL_OLD")) return false;
  // gss was tried on this pipeline (W recursion by one solver lane per bin): 189 k audio-s/s against 248 k in sel_pairs_kernel, whose
  // 512 threads per SM keep far more of the latency-bound W loads in flight; it stays there
  if (!(p.H == 512 && p.M <= 8 && algo == ALGO_MVDR && p.P >= 1 && p.P + 1 <= kSsSlots)) return false;
  for (int l = kSsBins; l < kL1K; l++)
    if (inband_host[l]) return false;
  return true;
}

cudaError_t launch_sel_stream(int algo, const KernelParams& p, cudaStream_t st) {
  const int use_tma = ((reinterpret_cast<uintptr_t>(p.in) & 15) == 0 && (p.in_stream_stride & 3) == 0 && (p.in_mic_stride & 3) == 0) ? 1 : 0;
  const size_t smem = sel_stream_smem();
  void (*k)(KernelParams, int) = nullptr;
  switch (algo) {
    case ALGO_MVDR: k = sel_stream_kernel<ALGO_MVDR>; break;
    default: return cudaErrorNotSupported;
  }
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  k<<<p.n_streams < sms ? p.n_streams : sms, SsCfg<ALGO_MVDR>::kThreads, smem, st>>>(p, use_tma);   // one persistent CTA per SM
  return cudaGetLastError();
}

}   // namespace bf
