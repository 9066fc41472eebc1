// Fused kernel for the magnitude-gated nodes (mvdr / lcmv / gss), 1024-point frames, M <= 8, sm_100a.
//
//   reference path replaced (citations /root/reference/beamform/src/): util.h:217-253,289-314 (window, framing,
//   OLA) and apply_weights of mvdr.cpp:62-115, lcmv.cpp:88-140, gss.cpp:96-156.
//
// One CTA (8 warps) owns one stream and walks its frame pairs (t, t+1) in order; two CTAs share an SM (96 KB of
// shared memory and <= 128 registers per thread each) so one CTA's latency-bound phase overlaps the other's math:
//   A   warp w: hops t-1..t+1 of microphone w arrive in its (then idle) spectrum tile by TMA bulk copy, issued as
//       soon as the previous pair's solves are done; window -> packed 1024-point FFT (registers + one swizzled
//       shared-memory transpose) -> Z_w in the same tile
//   B1  thread per logical bin: even/odd separation of every microphone's packed spectrum into X_t, X_{t+1}, FP32
//       magnitude gate with a guard band, append both frames to the per-stream history ring (global, L2-resident:
//       ring depth P+2 so a pair's two appends never overwrite a frame its own solves still need), default outputs
//   B1b guarded bins are re-decided in FP64 (exact double DFT of that bin) -> bit-exact selected-bin set
//   B2  MVDR: thread per selected (bin, frame); LCMV: lane pair per selected bin (one lane per frame of the pair): the
//       P-1 history frames the two frames share are summed once (half per lane + one shfl.xor), each lane adds the
//       frame only its own history holds; Cholesky, weights (GSS: thread per bin); ring frames arrive through a
//       per-thread cp.async pipeline; the LAST warp first runs the inverse FFT + overlap-add of the PREVIOUS pair
//   B3  Hermitian assembly of G = Yh_t + i*Yh_{t+1} for the next inverse
// Spectra never leave the SM except for the history ring the algorithm itself keeps (mvdr.cpp:99-101).
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "async_copy.cuh"
#include "bf_device.h"
#include "fft_reg.cuh"
#include "warp_fft1024.cuh"
#include "phase_b_select.cuh"
#include "sel_solve.cuh"

namespace bf {

constexpr int kSelWarps = 8;
constexpr int kSelThreads = kSelWarps * 32;


struct SelShared {
  float2 y[2][kL1K];
  float2 g[1024];
  float sqrtE[2][8];
  unsigned short items[2 * kL1K];
  unsigned short recheck[2 * kL1K];
  int n_items, n_recheck;
  unsigned char flag[2][kL1K];
  float tail[512];
  int nonfinite[2][2];   // [pair parity][frame]: the frame's spectrum holds inf/NaN (cold start, SURVEY B-10)
  uint64_t bars[kSelWarps];
  short sel_slot[kL1K];  // copies of the per-bin tables (B1 would otherwise chain two global loads per bin)
  unsigned char inband[kL1K];
};

// both frames of one logical bin from the packed half-scaled spectrum Z = FFT(0.5*w*(x_t + i x_{t+1}))
__device__ __forceinline__ void unpack2(const float2* z, int l, float2& x0, float2& x1) {
  const int j = (l == kL1K - 1) ? 511 : l;   // pseudo-bin: conj of bin N/2-1 (SURVEY B-4)
  const float2 a = z[j], b = z[(1024 - j) & 1023];
  x0 = make_float2(a.x + b.x, a.y - b.y);    // Z[j] + conj(Z[N-j])
  x1 = make_float2(a.y + b.y, b.x - a.x);    // -i (Z[j] - conj(Z[N-j]))
  if (l == kL1K - 1) { x0.y = -x0.y; x1.y = -x1.y; }
}

// Per-thread software pipeline over the item's P+1 ring frames (P history frames oldest first, then the item's own
// frame): every thread owns kStageDepth frame slots of M float2 in the (idle) spectrum tiles and keeps that many
// frames in flight with cp.async (8-byte pieces, coalesced across the neighbouring bins of a warp: the ring is
// bin-fastest).  ALL 256 threads of the CTA can carry an item at once (staging whole histories capped a batch at
// 92); the depth is chosen per batch from the number of items: a quiet pair (few items) keeps the whole history in
// flight in one L2 round trip, a busy pair runs depth 3 (a frame is consumed two rank-1 updates after its request).
template <int kStageDepth>
struct RingPipe {
  const float2* ring_l;   // &ring[slot 0][mic 0][this bin]
  float2* my;             // this thread's staging slots
  size_t mic_stride, slot_stride;
  int M, D;
  // request schedule, as frame offsets from frame t-P (ring slot slot_base): n_sh shared history frames starting at sh_off,
  // then (n_req > n_sh) the thread's own extra history frame and its own frame x
  int slot_base, n_sh, sh_off, extra_off, x_off, n_req;
  int next_k, next_stage, read_stage;
  __device__ __forceinline__ void issue() {   // request frame next_k (if any) and close one cp.async group either way
    if (next_k < n_req) {
      const int off = next_k < n_sh ? sh_off + next_k : (next_k == n_sh ? extra_off : x_off);
      int ring = slot_base + off;
      if (ring >= D) ring -= D;
      const float2* src = ring_l + (size_t)ring * slot_stride;
      float2* dst = my + next_stage * M;
      for (int i = 0; i < M; i++)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst + i)), "l"(src + (size_t)i * mic_stride) : "memory");
      if (++next_stage == kStageDepth) next_stage = 0;
    }
    next_k++;
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  __device__ __forceinline__ void start() {
#pragma unroll
    for (int d = 0; d < kStageDepth; d++) issue();
  }
  // the oldest requested frame, once it has landed
  template <int MM>
  __device__ __forceinline__ void take(float2 (&h)[MM]) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kStageDepth - 1) : "memory");
    const float2* src = my + read_stage * M;
#pragma unroll
    for (int i = 0; i < MM; i++) h[i] = (i < M) ? src[i] : make_float2(0.f, 0.f);
    if (++read_stage == kStageDepth) read_stage = 0;
  }
};

// Covariance for the two frames (t, t+1) of one bin by a PAIR of adjacent lanes (lane parity f = frame).  The histories
// of the two frames share P-1 frames (t-P+1 .. t-1): each lane accumulates half of them, one butterfly step (shfl.xor 1)
// gives both the full shared sum, then a selected lane adds the one frame only its own history holds (t-P for frame t,
// t for frame t+1; mvdr.cpp:87, :239-243), factorises and keeps its own frame x.  All 32 lanes of the warp call this
// (inactive ones with an empty schedule) so that the shuffle is unconditional.  Returns with the Cholesky factor in A.
template <int MM, typename T, class Pipe>
__device__ __forceinline__ void pair_cov_chol(const KernelParams& p, HermLower<MM, T>& A, T (&invd)[MM], Pipe& pipe, bool sel, float2 (&x)[MM]) {
  typedef HermLower<MM, T> HL;
  const int M = p.M;
#pragma unroll
  for (int i = 0; i < MM; i++) A.dg[i] = T(0);
#pragma unroll
  for (int i = 0; i < MM * (MM - 1) / 2; i++) A.lo[i] = mk<T>(T(0), T(0));
  pipe.start();
#pragma unroll 1
  for (int k = 0; k < pipe.n_sh; k++) {
    float2 hf[MM];
    pipe.template take<MM>(hf);
    cov_rank1<MM, T>(A, hf);
    pipe.issue();   // into the slot just consumed
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < MM; i++) A.dg[i] += __shfl_xor_sync(0xffffffffu, A.dg[i], 1);
#pragma unroll
  for (int i = 0; i < MM * (MM - 1) / 2; i++) {
    A.lo[i].x += __shfl_xor_sync(0xffffffffu, A.lo[i].x, 1);
    A.lo[i].y += __shfl_xor_sync(0xffffffffu, A.lo[i].y, 1);
  }
  if (sel) {
    float2 hf[MM];
    pipe.template take<MM>(hf);
    cov_rank1<MM, T>(A, hf);
    pipe.issue();
    pipe.template take<MM>(x);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (!sel) return;
  chol_in_place<MM, T>(p, A, invd);
}

// One selected (bin, frame) per thread: the P history frames in ring order, then the item's own frame (measured 1 %
// faster than the lane-pair scheme for the FP32 MVDR solve, 4 % slower for the FP64 LCMV solve: each node takes its own).
template <int MM, typename T, class Pipe>
__device__ __forceinline__ void single_cov_chol(const KernelParams& p, HermLower<MM, T>& A, T (&invd)[MM], Pipe& pipe, float2 (&x)[MM]) {
  typedef HermLower<MM, T> HL;
  const int M = p.M;
#pragma unroll
  for (int i = 0; i < MM; i++) A.dg[i] = T(0);
#pragma unroll
  for (int i = 0; i < MM * (MM - 1) / 2; i++) A.lo[i] = mk<T>(T(0), T(0));
  pipe.start();
#pragma unroll 1
  for (int k = 0; k < pipe.n_sh; k++) {
    float2 hf[MM];
    pipe.template take<MM>(hf);
    cov_rank1<MM, T>(A, hf);
    pipe.issue();   // into the slot just consumed
  }
  pipe.template take<MM>(x);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  chol_in_place<MM, T>(p, A, invd);
}

// One MVDR item, out of line: ptxas allocates its registers on its own instead of across the whole frame-pair loop
// (the kernel is capped at 128 registers for two CTAs per SM; inlined, the solve's live ranges spill into the transforms).
template <int kDepth>
__device__ __noinline__ float2 mvdr_item_fn(const KernelParams& p, const float2* ring_l, float2* my, int slot, const float2* steer_l) {
  RingPipe<kDepth> pipe;
  pipe.ring_l = ring_l;
  pipe.my = my;
  pipe.mic_stride = (size_t)p.Lsel;
  pipe.slot_stride = (size_t)p.M * p.Lsel;
  pipe.M = p.M; pipe.D = p.ring_depth;
  pipe.slot_base = slot;
  pipe.n_sh = p.P; pipe.sh_off = 0; pipe.extra_off = p.P; pipe.x_off = p.P; pipe.n_req = p.P + 1;
  pipe.next_k = 0; pipe.next_stage = 0; pipe.read_stage = 0;
  HermLower<8, float> A;
  float invd[8];
  float2 x[8];
  single_cov_chol<8, float>(p, A, invd, pipe, x);
  return mvdr_finish<8, float>(p, A, invd, x, steer_l);
}

template <int ALGO>
__global__ void __launch_bounds__(kSelThreads, ALGO == ALGO_LCMV ? 1 : 2) sel_pairs_kernel(const __grid_constant__ KernelParams p, const int use_tma) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* tw = reinterpret_cast<float2*>(smem_raw);             // [32][32]
  float2* ztiles = tw + 1024;                                     // [8][1024] exchange tile, then Z linear
  SelShared& sc = *reinterpret_cast<SelShared*>(ztiles + kSelWarps * 1024);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = p.M, D = p.ring_depth;
  constexpr int H = 512;
  const int s = blockIdx.x + p.stream_begin;

  for (int i = tid; i < 1024; i += kSelThreads) {
    const int k1 = i >> 5, l = i & 31;
    float sn, cs;
    sincospif(-2.0f * (float)((k1 * l) & 1023) / 1024.0f, &sn, &cs);
    tw[i] = make_float2(cs, sn);
  }
  for (int i = tid; i < H; i += kSelThreads) sc.tail[i] = p.tail[(size_t)s * H + i];
  for (int i = tid; i < kL1K; i += kSelThreads) { sc.sel_slot[i] = (short)p.sel_slot[i]; sc.inband[i] = p.inband[i]; }
  if (lane == 0) mbar_init(&sc.bars[warp], 1);
  mbar_fence_init();
  __syncthreads();

  double sd, cd;
  sincospi((double)lane / 1024.0, &sd, &cd);
  const float s_l = (float)(0.5 * sd), c_l = (float)(0.5 * cd);                 // analysis window * 0.5
  const float s_o = (float)(sd * p.out_scale), c_o = (float)(cd * p.out_scale); // synthesis window * out_amp / N

  float2* myz = ztiles + (size_t)warp * 1024;
  float* mystage = reinterpret_cast<float*>(myz);   // hops t-1..t+1 land in the tile before it holds Z
  const float* in_s = p.in + (size_t)s * p.in_stream_stride + (size_t)warp * p.in_mic_stride;
  const int nh = p.hop_end - p.hop_begin;
  const int npairs = (nh + 1) >> 1;

  // hops t-1..t+1 of microphone `warp` -> staging (TMA), hop -1 = per-stream state (util.h:275-277)
  auto issue = [&](int t) {
    if (lane != 0 || warp >= M) return;
    const bool two = t + 1 < p.hop_end;
    const uint32_t nb = (two ? 2u : 1u) * H * 4u;
    mbar_expect_tx(&sc.bars[warp], nb + H * 4u);
    const float* prev = (t - 1 < 0) ? p.prev_hop + ((size_t)s * M + warp) * H : in_s + (size_t)(t - 1) * H;
    bulk_g2s(mystage, prev, H * 4u, &sc.bars[warp]);
    bulk_g2s(mystage + H, in_s + (size_t)t * H, nb, &sc.bars[warp]);
  };
  if (use_tma && npairs > 0) issue(p.hop_begin);

#ifdef BF_PHASE_TIMERS   // nvcc -DBF_PHASE_TIMERS + env BF_DEBUG=1: per-phase cycle totals of CTA 0 (profiling builds only)
  long long ph_clk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long ph_t = clock64();
  int ph_items = 0;
#define BF_PHASE(i) do { if (p.debug == 1) { const long long n_ = clock64(); ph_clk[i] += n_ - ph_t; ph_t = n_; } } while (0)
#else
#define BF_PHASE(i) do { } while (0)
#endif
  for (int ip = 0; ip <= npairs; ip++) {
    const int t = p.hop_begin + 2 * ip;
    const bool live = ip < npairs;
    const bool two = t + 1 < p.hop_end;
    const int nf = two ? 2 : 1;
    // ------------------------------------------------------------------ A: forward transforms
    if (live && warp < M) {
      float2 v[32];
      if (use_tma) {
        mbar_wait(&sc.bars[warp], ip & 1);
      } else {   // unaligned caller buffers: plain warp copy, no prefetch
        const float* prev = (t - 1 < 0) ? p.prev_hop + ((size_t)s * M + warp) * H : in_s + (size_t)(t - 1) * H;
        for (int i = lane; i < H; i += 32) {
          mystage[i] = prev[i];
          mystage[H + i] = in_s[(size_t)t * H + i];
          mystage[2 * H + i] = two ? in_s[(size_t)(t + 1) * H + i] : 0.f;
        }
        __syncwarp();
      }
      static_for<0, 16>([&](auto r) {
        const float a = mystage[32 * r + lane], bb = mystage[512 + 32 * r + lane];
        const float c = two ? mystage[1024 + 32 * r + lane] : 0.0f;
        const float w0 = win1024<r>(s_l, c_l);
        const float w1 = win1024<r + 16>(s_l, c_l);
        v[brev5(r)] = make_float2(a * w0, bb * w0);
        v[brev5(r + 16)] = make_float2(bb * w1, c * w1);
      });
      __syncwarp();   // staged samples consumed: the tile becomes the exchange buffer, then Z
      {   // sqrt of the windowed frame energies: scale of the FP32 FFT's absolute error (gate guard band)
        float e0 = 0.f, e1 = 0.f;
#pragma unroll
        for (int r = 0; r < 32; r++) { e0 = fmaf(v[r].x, v[r].x, e0); e1 = fmaf(v[r].y, v[r].y, e1); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { e0 += __shfl_xor_sync(0xffffffffu, e0, o); e1 += __shfl_xor_sync(0xffffffffu, e1, o); }
        if (lane == 0) { sc.sqrtE[0][warp] = 2.0f * sqrtf(e0); sc.sqrtE[1][warp] = 2.0f * sqrtf(e1); }
      }
      warp_fft1024_fwd(v, myz, tw, lane);
#pragma unroll
      for (int k2 = 0; k2 < 32; k2++) myz[k2 * 32 + lane] = v[k2];
    }
    BF_PHASE(0);
    if (tid == 0) { sc.n_items = 0; sc.n_recheck = 0; sc.nonfinite[ip & 1][0] = 0; sc.nonfinite[ip & 1][1] = 0; }
    __syncthreads();   // (a) Z complete; G of the previous pair complete
    BF_PHASE(1);
    const int fr0 = (p.ring_slot0 + (t - p.hop_begin)) % D;   // ring slot of frame t
    // ------------------------------------------------------------------ B1: gate, history append, defaults
    if (live) {
      for (int l = tid; l < kL1K; l += kSelThreads) {
        const bool inb = sc.inband[l] != 0 && !(ALGO == ALGO_MVDR && l == 0);
        float2 x0[8], x1[8];
        float st0 = 0.f, st1 = 0.f;
        if (!inb) {   // out of band: output 0 (mvdr.cpp:103), except mvdr's bin 0 which passes microphone 0 through
          float2 a = make_float2(0.f, 0.f), b = a;
          if (ALGO == ALGO_MVDR && l == 0) unpack2(ztiles, 0, a, b);
          sc.flag[0][l] = 0; sc.flag[1][l] = 0;
          sc.y[0][l] = a; sc.y[1][l] = b;
          continue;
        }
#pragma unroll
        for (int ch = 0; ch < 8; ch++) {
          if (ch < M) {
            unpack2(ztiles + ch * 1024, l, x0[ch], x1[ch]);
            st0 += sqrt_approx(fmaf(x0[ch].x, x0[ch].x, x0[ch].y * x0[ch].y));
            st1 += sqrt_approx(fmaf(x1[ch].x, x1[ch].x, x1[ch].y * x1[ch].y));
          } else {
            x0[ch] = x1[ch] = make_float2(0.f, 0.f);
          }
        }
        float2 y0 = make_float2(0.f, 0.f), y1 = y0;
        unsigned char f0 = 0, f1 = 0;
        {
          float es0 = 0.f, es1 = 0.f;
          for (int ch = 0; ch < M; ch++) { es0 += sc.sqrtE[0][ch]; es1 += sc.sqrtE[1][ch]; }
          const float g0 = 2.0e-5f * es0 + 1.0e-6f * p.thr_mag, g1 = 2.0e-5f * es1 + 1.0e-6f * p.thr_mag;
          if (fabsf(st0 - p.thr_mag) <= g0) sc.recheck[atomicAdd(&sc.n_recheck, 1)] = (unsigned short)(l * 2);
          else if (st0 > p.thr_mag) f0 = 1;
          if (two) {
            if (fabsf(st1 - p.thr_mag) <= g1) sc.recheck[atomicAdd(&sc.n_recheck, 1)] = (unsigned short)(l * 2 + 1);
            else if (st1 > p.thr_mag) f1 = 1;
          }
          y0 = make_float2(0.01f * x0[0].x, 0.01f * x0[0].y);   // mvdr.cpp:96 (overwritten when selected)
          y1 = make_float2(0.01f * x1[0].x, 0.01f * x1[0].y);
          if (ALGO != ALGO_GSS) {   // mvdr.cpp:99-101: every in-band bin appends every frame
            float2* ring_l = p.hist + (size_t)s * D * M * p.Lsel + sc.sel_slot[l];
            int fs = fr0;
#pragma unroll
            for (int f = 0; f < 2; f++) {
              if (f < nf) {
                float2* dst = ring_l + (size_t)fs * M * p.Lsel;
#pragma unroll
                for (int i = 0; i < 8; i++)
                  if (i < M) dst[(size_t)i * p.Lsel] = f ? x1[i] : x0[i];
              }
              if (++fs == D) fs = 0;
            }
          }
        }
        sc.flag[0][l] = f0; sc.flag[1][l] = two ? f1 : 0;
        sc.y[0][l] = y0; sc.y[1][l] = y1;
      }
    }
    BF_PHASE(2);
    __syncthreads();   // (b)
    // ------------------------------------------------------------------ B1b: FP64 re-decision of guarded bins
    if (live && sc.n_recheck > 0) {
      for (int q = warp; q < sc.n_recheck; q += kSelWarps) {
        const int l = sc.recheck[q] >> 1, f = sc.recheck[q] & 1;
        const bool sel = gate_fp64(p, s, t, l, f, lane);
        if (lane == 0 && sel) sc.flag[f][l] = 1;
      }
    }
    __syncthreads();
    // work list: one item per bin selected in either frame; runs of consecutive bins stay contiguous, so neighbouring
    // threads of B2 read neighbouring ring addresses.  lcmv: one item per bin, a lane pair per item (one lane per frame,
    // the shared part of the two histories summed once); mvdr: one item and one thread per selected (bin, frame);
    // gss: one item and one thread per bin (its recursion is sequential over the frames).
    if (live) {
      for (int f = 0; f < (ALGO == ALGO_MVDR ? nf : 1); f++)
        for (int base = warp * 32; base < kL1K; base += kSelThreads) {
          const int l = base + lane;
          bool on = false;
          if (l < kL1K) on = (ALGO == ALGO_MVDR) ? (sc.flag[f][l] != 0) : ((sc.flag[0][l] | sc.flag[1][l]) != 0);
          const unsigned m = __ballot_sync(0xffffffffu, on);
          int pos = 0;
          if (lane == 0 && m) pos = atomicAdd(&sc.n_items, __popc(m));
          pos = __shfl_sync(0xffffffffu, pos, 0);
          if (on) sc.items[pos + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(l * 2 + f);
        }
    }
    __syncthreads();   // (c) work list complete
    BF_PHASE(3);
#ifdef BF_PHASE_TIMERS
    ph_items += sc.n_items;
#endif
    // ------------------------------------------------------------------ inverse of the PREVIOUS pair (last warp)
    if (warp == kSelWarps - 1 && ip > 0) {
      const int tp = t - 2;
      const bool ptwo = tp + 1 < p.hop_end;
      float2 v[32];
      static_for<0, 32>([&](auto n1) { const float2 gg = sc.g[n1 * 32 + lane]; v[brev5(n1)] = make_float2(gg.y, gg.x); });
      __syncwarp();
      warp_fft1024_fwd(v, sc.g, tw, lane);   // IFFT(G) = swap(FFT(swap(G))); G itself is the exchange tile
      float* o0 = p.out + (size_t)s * p.out_stream_stride + (size_t)tp * H;
      // one inf/NaN bin makes the reference's whole inverse frame NaN; the two frames of a pair share one complex
      // transform here, so a poisoned frame was zeroed in B3 and is re-poisoned now without touching its partner
      const float bad0 = sc.nonfinite[(ip - 1) & 1][0] ? __int_as_float(0x7fc00000) : 0.f;
      const float bad1 = sc.nonfinite[(ip - 1) & 1][1] ? __int_as_float(0x7fc00000) : 0.f;
      static_for<0, 16>([&](auto m2) {
        const float w0 = win1024<m2>(s_o, c_o);
        const float w1 = win1024<m2 + 16>(s_o, c_o);
        const float y0a = v[m2].y * w0 + bad0, y0b = v[m2 + 16].y * w1 + bad0;   // frame t: first / second half
        const float y1a = v[m2].x * w0 + bad1, y1b = v[m2 + 16].x * w1 + bad1;   // frame t+1
        o0[32 * m2 + lane] = sc.tail[32 * m2 + lane] + y0a;        // util.h:301-302
        if (ptwo) o0[H + 32 * m2 + lane] = y0b + y1a;
        sc.tail[32 * m2 + lane] = ptwo ? y1b : y0b;
      });
    }
    // ------------------------------------------------------------------ B2: per-item solves
    if (live && ALGO == ALGO_GSS) {
      // a group of kGssGroup lanes per selected bin (rows of W split over the lanes), both frames in turn (W carries over)
      const int g = lane & (kGssGroup - 1);
      const unsigned gmask = ((1u << kGssGroup) - 1u) << (lane & ~(kGssGroup - 1));
      for (int q = tid / kGssGroup; q < sc.n_items; q += kSelThreads / kGssGroup) {
        const int l = sc.items[q] >> 1;
        const int slot = p.sel_slot[l];
        const float2* steer_l = p.steer + (size_t)l * p.C * M;
        float2* Wg = p.gss_w + (size_t)s * BF_GSS_ROWS * M * p.Lsel + slot;   // [B][BF_GSS_ROWS][M][Lsel]
        GssRows R;   // this lane's rows of W: loaded once, carried through both frames in registers, stored once
        gss_rows_load(p, Wg, (size_t)p.Lsel, g, R);
        for (int ff = 0; ff < nf; ff++) {
          if (!sc.flag[ff][l]) continue;
          float2 x[8];
#pragma unroll
          for (int ch = 0; ch < 8; ch++) {
            float2 a, b;
            if (ch < M) unpack2(ztiles + ch * 1024, l, a, b); else a = b = make_float2(0.f, 0.f);
            x[ch] = ff ? b : a;
          }
          const float2 y0 = gss_rows_step(p, R, x, steer_l, g, gmask);
          if (g == 0) sc.y[ff][l] = y0;
        }
        gss_rows_store(p, Wg, (size_t)p.Lsel, g, R);
      }
    }
    if (live && ALGO != ALGO_GSS) {
      // mvdr / lcmv.  The spectrum tiles are dead after B1 (the current frame is in the ring too), so they become the
      // per-thread staging slots of the ring pipeline (RingPipe).
      if (ALGO == ALGO_MVDR) {   // one thread per selected (bin, frame)
        const int n_items = sc.n_items;
        for (int base = 0; base < n_items; base += kSelThreads) {
          const int nb = min(n_items - base, kSelThreads);
          const int cap = (kSelWarps * 1024) / nb;           // float2 of staging per item of this batch
          const int q = base + tid;
          auto run = [&](auto depth_c) {
            constexpr int kDepth = decltype(depth_c)::value;
            const int pitch = (kDepth * M) | 1;              // odd pitch in float2: conflict-free 8-byte accesses
            if (q < n_items) {
              const int l = sc.items[q] >> 1, f = sc.items[q] & 1;
              int slot = (fr0 + f - p.P) % D;                // ring slot of frame (t+f) - P
              if (slot < 0) slot += D;
              sc.y[f][l] = mvdr_item_fn<kDepth>(p, p.hist + (size_t)s * D * M * p.Lsel + sc.sel_slot[l], ztiles + (size_t)tid * pitch, slot,
                                                p.steer + (size_t)l * p.C * M);
            }
          };
          if (cap > 11 * M) run(IC<11>{});                   // P + 1 <= 11 frames at once is the common launch value (P = 10)
          else if (cap > 6 * M) run(IC<6>{});
          else run(IC<3>{});
        }
      } else {   // lcmv: lane pair (2i, 2i+1) carries item i of the batch
      const int n_items = sc.n_items;
      constexpr int kPairs = kSelThreads / 2;
      const int n_sh_tot = p.P - 1, n_sh0 = (n_sh_tot + 1) / 2;   // shared history frames; lane 0 takes the first n_sh0
      for (int base = 0; base < n_items; base += kPairs) {
        const int nbp = min(n_items - base, kPairs);         // items (pairs) in this batch
        if (warp * 32 >= 2 * nbp) continue;                  // warp-uniform: no item in this warp
        const int cap = (kSelWarps * 1024) / (2 * nbp);      // float2 of staging per thread of this batch
        const int q = base + (tid >> 1), f = tid & 1;
        const bool on = q < n_items;
        auto run = [&](auto depth_c) {
          constexpr int kDepth = decltype(depth_c)::value;
          const int pitch = (kDepth * M) | 1;                // odd pitch in float2: conflict-free 8-byte accesses
          const int l = on ? (sc.items[q] >> 1) : 0;
          const bool sel = on && sc.flag[f][l] != 0;
          RingPipe<kDepth> pipe;
          pipe.ring_l = p.hist + (size_t)s * D * M * p.Lsel + (on ? sc.sel_slot[l] : 0);
          pipe.my = ztiles + (size_t)tid * pitch;
          pipe.mic_stride = (size_t)p.Lsel;
          pipe.slot_stride = (size_t)M * p.Lsel;
          pipe.M = M; pipe.D = D;
          int slot = (fr0 - p.P) % D;                        // ring slot of frame t - P
          if (slot < 0) slot += D;
          pipe.slot_base = slot;
          pipe.n_sh = on ? (f ? n_sh_tot - n_sh0 : n_sh0) : 0;
          pipe.sh_off = f ? 1 + n_sh0 : 1;                   // shared frames are t-P+1 .. t-1
          pipe.extra_off = f ? p.P : 0;                      // frame t+1 also has frame t, frame t also has frame t-P
          pipe.x_off = p.P + f;
          pipe.n_req = pipe.n_sh + (sel ? 2 : 0);
          pipe.next_k = 0; pipe.next_stage = 0; pipe.read_stage = 0;
          typedef typename std::conditional<ALGO == ALGO_MVDR, float, double>::type T;
          HermLower<8, T> A;
          T invd[8];
          float2 x[8];
          pair_cov_chol<8, T>(p, A, invd, pipe, sel, x);
          if (sel) {
            const float2* steer_l = p.steer + (size_t)l * p.C * M;
            sc.y[f][l] = (ALGO == ALGO_MVDR) ? mvdr_finish<8, T>(p, A, invd, x, steer_l) : lcmv_finish<8, T>(p, A, invd, x, steer_l);
          }
        };
        if (cap > 7 * M) run(IC<7>{});                       // at most ceil((P-1)/2) + 2 = 7 frames per thread at P = 10: all in flight
        else run(IC<3>{});
      }
      }
    }
    BF_PHASE(4);
    __syncthreads();   // (e) Y complete; G consumed by the inverse; Z no longer needed
    BF_PHASE(5);
    if (use_tma && ip + 1 < npairs) {   // the spectrum tiles are free: stage the next pair's hops into them
      fence_proxy_async();
      issue(t + 2);
    }
    // ------------------------------------------------------------------ B3: Hermitian assembly, diagnostics
    if (live && (ALGO == ALGO_MVDR || ALGO == ALGO_LCMV)) {
      bool b0 = false, b1 = false;
      for (int l = tid; l < kL1K; l += kSelThreads) {
        const float2 y0 = sc.y[0][l], y1 = sc.y[1][l];
        b0 |= !(isfinite(y0.x) && isfinite(y0.y));
        b1 |= two && !(isfinite(y1.x) && isfinite(y1.y));
      }
      if (b0) sc.nonfinite[ip & 1][0] = 1;
      if (b1) sc.nonfinite[ip & 1][1] = 1;
      __syncthreads();
    }
    if (live) {
      const bool z0 = sc.nonfinite[ip & 1][0] != 0, z1 = sc.nonfinite[ip & 1][1] != 0;
      for (int l = tid; l < kL1K; l += kSelThreads) {
        if (l <= 512) {
          float2 y0 = z0 ? make_float2(0.f, 0.f) : sc.y[0][l], y1 = (two && !z1) ? sc.y[1][l] : make_float2(0.f, 0.f);
          if (l == 511) {   // Hermitian part of the asymmetric pair (N/2-1, N/2+1): Yh = (Y[N/2-1] + conj(Y[N/2+1])) / 2
            const float2 p0 = z0 ? make_float2(0.f, 0.f) : sc.y[0][kL1K - 1], p1 = (two && !z1) ? sc.y[1][kL1K - 1] : make_float2(0.f, 0.f);
            y0 = make_float2(0.5f * (y0.x + p0.x), 0.5f * (y0.y - p0.y));
            y1 = make_float2(0.5f * (y1.x + p1.x), 0.5f * (y1.y - p1.y));
          }
          if (l == 0 || l == 512) { y0.y = 0.f; y1.y = 0.f; }   // Re(): self-conjugate bins
          sc.g[l] = make_float2(y0.x - y1.y, y0.y + y1.x);                          // Yh_t + i Yh_{t+1}
          if (l > 0 && l < 512) sc.g[1024 - l] = make_float2(y0.x + y1.y, y1.x - y0.y);   // conj(Yh_t) + i conj(Yh_{t+1})
        }
        if (p.capture) {
          for (int f = 0; f < nf; f++) {
            unsigned char* cap = p.capture + (size_t)s * p.capture_stream_stride + (size_t)(t + f) * p.N;
            const unsigned char fl = sc.flag[f][l];
            if (l <= 512) {
              cap[l] = fl;
              if (l > 0 && l < 511) cap[1024 - l] = fl;
            } else {
              cap[513] = fl;
            }
          }
        }
      }
    }
  }
  BF_PHASE(6);
#ifdef BF_PHASE_TIMERS
  if (p.debug == 1 && blockIdx.x == 0 && (tid == 0 || tid == 255))
    printf("sel_pairs phases (clk/pair, thread %d): A %lld | wait(a) %lld | B1 %lld | B1b+list %lld | own B2 %lld | wait(e) %lld | B3 %lld | items/pair %.1f\n",
           tid, ph_clk[0] / npairs, ph_clk[1] / npairs, ph_clk[2] / npairs, ph_clk[3] / npairs, ph_clk[4] / npairs, ph_clk[5] / npairs,
           ph_clk[6] / npairs, (double)ph_items / npairs);
#endif
#undef BF_PHASE
  __syncthreads();
  for (int i = tid; i < H; i += kSelThreads) p.tail[(size_t)s * H + i] = sc.tail[i];
}

size_t sel_pairs_smem() { return sizeof(float2) * (1024 + kSelWarps * 1024) + sizeof(SelShared) + 128; }

bool sel_pairs_supported(const KernelParams& p, int algo) {
  return p.H == 512 && p.M <= 8 && (algo == ALGO_MVDR || algo == ALGO_LCMV || algo == ALGO_GSS) && p.C <= kMaxC;
}

cudaError_t launch_sel_pairs(int algo, const KernelParams& p, cudaStream_t st) {
  const int use_tma = ((reinterpret_cast<uintptr_t>(p.in) & 15) == 0 && (p.in_stream_stride & 3) == 0 && (p.in_mic_stride & 3) == 0) ? 1 : 0;
  const size_t smem = sel_pairs_smem();
  void (*k)(KernelParams, int) = nullptr;
  switch (algo) {
    case ALGO_MVDR: k = sel_pairs_kernel<ALGO_MVDR>; break;
    case ALGO_LCMV: k = sel_pairs_kernel<ALGO_LCMV>; break;
    case ALGO_GSS: k = sel_pairs_kernel<ALGO_GSS>; break;
    default: return cudaErrorNotSupported;
  }
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<p.n_streams, kSelThreads, smem, st>>>(p, use_tma);
  return cudaGetLastError();
}

}   // namespace bf
