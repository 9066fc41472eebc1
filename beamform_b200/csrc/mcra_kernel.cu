// Stand-alone MCRA noise-reduction node for 1024-point frames, sm_100a: every warp is an independent worker.
//
//   reference path replaced (citations /root/reference/beamform/src/): util.h:217-253,289-314 (window, framing, OLA),
//   mcra.cpp:62-155 (|X|^2, 3-tap smoothing over neighbouring bins, recursive averaging, minima tracking over windows of
//   L frames, noise estimate, spectral subtraction with the phase of X) on the FIRST microphone only.
//
// The recursion runs along the frames of a stream, so a warp owns whole streams (stream = warp index + k * warps of the
// grid) and walks their frame pairs in order, the way das_pairs_kernel walks its range:
//   1. hops t-1..t+1 of microphone 0 land in the warp's shared-memory tile (TMA bulk copy, mbarrier),
//   2. window, pack z = 0.5 w (frame_t + i frame_{t+1}), warp-private 1024-point FFT (32 points per lane in registers),
//   3. unpack by warp shuffles (Z[N-j] sits in lane (32-lane)%32, register 31-k2), |X|^2 of both frames to two lines in
//      the (now idle) exchange tile so that the 3-tap smoothing reads its neighbours there,
//   4. the MCRA recursion per bin on the 4 state scalars the warp keeps in shared memory for the whole stream,
//      Y = max(0, |X| - sqrt(lambda)) out_amp e^{i arg X} -> G = Yh_t + i Yh_{t+1} assembled IN PLACE in the spectrum
//      registers (one shuffle hands G[N-l] to the partner lane),
//   5. inverse transform through the same code (IFFT(x) = swap(FFT(swap(x)))), synthesis window, overlap-add with the
//      tail in registers.
// No block-level barrier after start-up; the state (mpf_state slots 0-3, layout shared with the CTA-per-stream kernel
// frames_kernel_mcra) is read at the start of a launch and written back at its end.
#include <cstdlib>

#include "async_copy.cuh"
#include "bf_device.h"
#include "fft_reg.cuh"
#include "warp_fft1024.cuh"

namespace bf {

constexpr int kMcWarps = 8;   // 10 and 12 warps (204 / 168 registers) measured slower at the bench shape: whole streams are the unit of work, 2 368 streams = 2 per warp with 8
constexpr int kMcLine = 17 * 32;   // bins 0..543 in (row k2, lane) order; rows 0..15 and bin 512 are the half spectrum

struct McWarp {
  float2 tile[1024];          // staged hops -> FFT exchange tile -> |X_f[j]|^2 lines (2 x kMcLine floats) -> exchange tile of the inverse
  float st[4][kMcLine];       // S_prev, S_tmp, S_min, lambda
  uint64_t bar;
  uint64_t pad;
};


__device__ __forceinline__ float mc_sqrt(float x) {   // MUFU.SQRT; NaN and negative inputs behave as sqrtf
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One bin, both frames of the pair: mcra.cpp:77-135 on the state the warp keeps in shared memory, then the bin's two cells of
// G.  Inlined at its 17 call sites (compile-time register rows): as a real call it cost more than the instruction-cache misses
// it saved (profiles/r02_experiments.md).  bk: bit0/1 window reset in frame 0/1, bit2/3 first window, bit4 frame t+1 exists.
__device__ __forceinline__ void mcra_bin(const KernelParams& p, McWarp& my, int l, float2 x0, float2 x1, int bk, float inv_cl_0, float inv_cl_1,
                                         float2& g_lo, float2& g_hi) {
  constexpr int H = 512;
  const int nf = (bk & 16) ? 2 : 1;
  const float2 xf[2] = {x0, x1};
  float S_prev = my.st[0][l], S_tmp = my.st[1][l], S_min = my.st[2][l], lam = my.st[3][l];
  float2 yy[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
  for (int f = 0; f < 2; f++) {
    if (f >= nf) break;
    const bool reset_f = (bk >> f) & 1;
    const float inv_cl_f = f ? inv_cl_1 : inv_cl_0;
    const int fst_f = (bk >> (2 + f)) & 1;
    const float* ps = reinterpret_cast<const float*>(my.tile) + f * kMcLine;
    const float pq = ps[l];
    float Sf;
    if (l == 0) {
      Sf = sqrtf(pq);   // "passing on the DC component": abs, not squared (mcra.cpp:81)
    } else {
      // neighbours j-1, j, j+1 clipped to [1, N-1]; bin N/2+1 is the mirror of N/2-1 (row 16 of the line holds it)
      const float lo = (l - 1 >= 1) ? ps[l - 1] : 0.f;
      const float hi = ps[l + 1];
      Sf = 0.25f * lo + 0.5f * pq + 0.25f * hi;
    }
    const float S = p.mcra_alphaS * S_prev + (1.0f - p.mcra_alphaS) * Sf;
    if (reset_f) { S_min = fminf(S_tmp, S); S_tmp = S; }
    else { S_min = fminf(S_min, S); S_tmp = fminf(S_tmp, S); }
    {   // branch-free: the three conditions differ from lane to lane
      const bool upd = fst_f || S < S_min * p.mcra_delta || lam > pq;
      const bool avg = fst_f && inv_cl_f > p.mcra_alphaD;
      const float ca = avg ? inv_cl_f : p.mcra_alphaD2, cb = avg ? 1.0f - inv_cl_f : 1.0f - p.mcra_alphaD;
      lam = upd ? ca * lam + cb * pq : lam;
    }
    S_prev = S;
    if (l > 0) {
      // |X| = pq * rsqrt(pq) and sqrt(lambda) through the MUFU approximations (~2 ulp): the IEEE square root is a 10-instruction
      // sequence with a slow-path branch and was 13 % of the kernel's instructions
      const float r0 = rsqrtf(pq);
      const float mag_x = pq == 0.f ? 0.f : pq * r0;   // NaN stays NaN
      const float sl = mc_sqrt(lam);
      float mag = p.out_only_noise ? sl * p.out_amp : (mag_x - sl) * p.out_amp;
      if (!p.out_only_noise) mag = mag < 0.f ? 0.f : mag;   // (a NaN passes, as in mcra.cpp)
      const float2 unit = pq > 0.f ? make_float2(xf[f].x * r0, xf[f].y * r0) : make_float2(1.f, 0.f);   // e^{i arg X}
      yy[f] = make_float2(mag * unit.x, mag * unit.y);
    }
  }
  my.st[0][l] = S_prev; my.st[1][l] = S_tmp; my.st[2][l] = S_min; my.st[3][l] = lam;
  float2 y0 = yy[0], y1 = yy[1];
  if (l == 0 || l == H) { y0.y = 0.f; y1.y = 0.f; }
  g_lo = make_float2(y0.x - y1.y, y0.y + y1.x);   // G[l]     = Yh_t + i Yh_{t+1}
  g_hi = make_float2(y0.x + y1.y, y1.x - y0.y);   // G[N - l] = conj(Yh_t) + i conj(Yh_{t+1})   (0 < l < N/2)
}

__global__ void __launch_bounds__(kMcWarps * 32, 1) mcra_pairs_kernel(const __grid_constant__ KernelParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* tw = reinterpret_cast<float2*>(smem_raw);   // [32][32]
  McWarp* ws = reinterpret_cast<McWarp*>(tw + 1024);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int H = 512, L = 514;

  for (int i = tid; i < 1024; i += blockDim.x) {
    const int k1 = i >> 5, l = i & 31;
    float sn, cs;
    sincospif(-2.0f * (float)((k1 * l) & 1023) / 1024.0f, &sn, &cs);
    tw[i] = make_float2(cs, sn);
  }
  McWarp& my = ws[warp];
  if (lane == 0) mbar_init(&my.bar, 1);
  mbar_fence_init();
  __syncthreads();

  double sd, cd;
  sincospi((double)lane / 1024.0, &sd, &cd);
  const float s_l = (float)(0.5 * sd), c_l = (float)(0.5 * cd);                  // analysis window * 0.5
  const float s_o = (float)(sd * p.out_scale), c_o = (float)(cd * p.out_scale);  // synthesis window * out_amp / N
  const int nh = p.hop_end - p.hop_begin;
  const int npairs = (nh + 1) >> 1;
  unsigned job = 0;

  auto issue = [&](int s, int t, bool two) {
    if (lane != 0) return;
    const float* base = p.in + (size_t)s * p.in_stream_stride;
    float* dst = reinterpret_cast<float*>(my.tile);
    const uint32_t nb = (two ? 2u : 1u) * H * 4u;
    mbar_expect_tx(&my.bar, nb + H * 4u);
    const float* prev = (t - 1 < 0) ? p.prev_hop + (size_t)s * p.M * H : base + (size_t)(t - 1) * H;
    bulk_g2s(dst, prev, H * 4u, &my.bar);
    bulk_g2s(dst + H, base + (size_t)t * H, nb, &my.bar);
  };

#pragma unroll 1
  for (int sl = blockIdx.x * kMcWarps + warp; sl < p.n_streams; sl += gridDim.x * kMcWarps) {
    const int s = sl + p.stream_begin;
    fence_proxy_async();
    __syncwarp();
    issue(s, p.hop_begin, p.hop_begin + 1 < p.hop_end);
    float* stg = p.mpf_state + (size_t)s * 7 * L;
    for (int j = lane; j < kMcLine; j += 32) {
#pragma unroll
      for (int q = 0; q < 4; q++) my.st[q][j] = (j <= H) ? stg[q * L + j] : 0.f;
    }
    float tail[16];   // OLA tail of the stream: sample 32 m2 + lane of the last frame's second half
#pragma unroll
    for (int m2 = 0; m2 < 16; m2++) tail[m2] = p.tail[(size_t)s * H + 32 * m2 + lane];
    int cur_L = p.mcra_cur_L0, first_L = p.mcra_first0;
    __syncwarp();

#pragma unroll 1
    for (int ip = 0; ip < npairs; ip++) {
      const int t = p.hop_begin + 2 * ip;
      const bool two = t + 1 < p.hop_end;
      // the forward and the inverse transform go through ONE call site (a rolled 2-trip loop): one copy of the butterfly code
      float2 v[32];   // samples -> Z -> G (assembled in place) -> output samples
#pragma unroll 1
      for (int dir = 0; dir < 2; dir++) {
        if (dir == 0) {
          const float* stage = reinterpret_cast<const float*>(my.tile);
          mbar_wait(&my.bar, job & 1);
          job++;
          static_for<0, 16>([&](auto r) {
            const float a = stage[32 * r + lane], bb = stage[512 + 32 * r + lane];
            const float c = two ? stage[1024 + 32 * r + lane] : 0.0f;
            const float w0 = win1024<r>(s_l, c_l);        // 0.5 * w[32r + lane]
            const float w1 = win1024<r + 16>(s_l, c_l);   // 0.5 * w[32r + lane + 512]
            v[brev5(r)] = make_float2(a * w0, (two ? bb : 0.f) * w0);
            v[brev5(r + 16)] = make_float2(bb * w1, c * w1);
          });
        } else {
          // inverse through the forward code: v[row] = G[32 row + lane]; stage 1 wants row n1 in slot brev5(n1), parts swapped
          static_for<0, 32>([&](auto n1c) {
            constexpr int n1 = decltype(n1c)::value, r = brev5(n1);
            if constexpr (n1 < r) {
              const float2 a = v[n1], b = v[r];
              v[n1] = make_float2(b.y, b.x);
              v[r] = make_float2(a.y, a.x);
            } else if constexpr (n1 == r) {
              v[n1] = make_float2(v[n1].y, v[n1].x);
            }
          });
        }
        __syncwarp();   // staged samples / G consumed: the tile becomes the exchange buffer
        warp_fft1024_fwd(v, my.tile, tw, lane, [&]() {
          if (dir == 1 && ip + 1 < npairs) {   // the tile is free: the next pair's hops may land
            fence_proxy_async();
            issue(s, t + 2, t + 3 < p.hop_end);
          }
        });
        if (dir == 0) {
          // v[k2] = Z[32 k2 + lane].  X_t[j] = Z[j] + conj(Z[N-j]), X_{t+1}[j] = -i (Z[j] - conj(Z[N-j]))
          const int src_lane = (32 - lane) & 31;
          auto unpack = [&](auto k2c, float2& x0, float2& x1) {
            constexpr int k2 = decltype(k2c)::value;
            const float2 a = v[k2];
            float2 b;
            b.x = __shfl_sync(0xffffffffu, v[31 - k2].x, src_lane);
            b.y = __shfl_sync(0xffffffffu, v[31 - k2].y, src_lane);
            if (lane == 0) b = v[(32 - k2) & 31];
            x0 = make_float2(a.x + b.x, a.y - b.y);
            x1 = make_float2(a.y + b.y, b.x - a.x);
          };
          float* psq = reinterpret_cast<float*>(my.tile);   // the exchange step is over: the tile holds the two |X|^2 lines now
          static_for<0, 17>([&](auto k2c) {   // in_fft_square (mcra.cpp:73-76); row 16 = bins 512..543 (|X[512+i]| = |X[512-i]|)
            constexpr int k2 = decltype(k2c)::value;
            float2 x0, x1;
            unpack(k2c, x0, x1);
            psq[32 * k2 + lane] = fmaf(x0.x, x0.x, x0.y * x0.y);
            psq[kMcLine + 32 * k2 + lane] = fmaf(x1.x, x1.x, x1.y * x1.y);
          });
          __syncwarp();
          // window bookkeeping of the two frames (mcra.cpp:100-113), global per frame
          const bool reset_0 = cur_L > p.mcra_L;
          if (reset_0) { cur_L = 1; first_L = 0; } else { cur_L++; }
          const float inv_cl_0 = 1.0f / (float)cur_L;
          const int fst_0 = first_L;
          bool reset_1 = false;
          float inv_cl_1 = 1.f;
          int fst_1 = first_L;
          if (two) {
            reset_1 = cur_L > p.mcra_L;
            if (reset_1) { cur_L = 1; first_L = 0; } else { cur_L++; }
            inv_cl_1 = 1.0f / (float)cur_L; fst_1 = first_L;
          }
          const int bk = (reset_0 ? 1 : 0) | (reset_1 ? 2 : 0) | (fst_0 ? 4 : 0) | (fst_1 ? 8 : 0) | (two ? 16 : 0);
          // G is assembled IN PLACE in the spectrum registers: after its unpack, row k2 of a lane and row 31-k2 of its partner
          // lane (lane 0: its own row 32-k2) are dead, and those are the cells G[l] and G[N-l] of the bins of this step
          static_for<0, 17>([&](auto k2c) {
            constexpr int k2 = decltype(k2c)::value;
            float2 x0, x1;
            unpack(k2c, x0, x1);                 // all lanes: the shuffles are warp-wide
            float2 g_lo = make_float2(0.f, 0.f), g_hi = make_float2(0.f, 0.f);
            if (k2 < 16 || lane == 0)            // only the Nyquist bin of row 16 is an output bin
              mcra_bin(p, my, 32 * k2 + lane, x0, x1, bk, inv_cl_0, inv_cl_1, g_lo, g_hi);
            if constexpr (k2 < 16) {
              const float hx = __shfl_sync(0xffffffffu, g_hi.x, src_lane), hy = __shfl_sync(0xffffffffu, g_hi.y, src_lane);
              v[k2] = g_lo;
              if (lane != 0) v[31 - k2] = make_float2(hx, hy);
              else if (k2 >= 1) v[32 - k2] = g_hi;
            } else {
              if (lane == 0) v[16] = g_lo;
            }
          });
          __syncwarp();
        } else {
          // v = swap(IFFT(G)): frame t in .y, frame t+1 in .x; synthesis window + overlap-add (util.h:244-253,301-302)
          float* o0 = p.out + (size_t)s * p.out_stream_stride + (size_t)t * H;
          static_for<0, 16>([&](auto m2) {
            const float w0 = win1024<m2>(s_o, c_o);
            const float w1 = win1024<m2 + 16>(s_o, c_o);
            const float y0a = v[m2].y * w0, y0b = v[m2 + 16].y * w1;   // frame t: first / second half
            const float y1a = v[m2].x * w0, y1b = v[m2 + 16].x * w1;   // frame t+1
            o0[32 * m2 + lane] = tail[m2] + y0a;
            if (two) o0[H + 32 * m2 + lane] = y0b + y1a;
            tail[m2] = two ? y1b : y0b;
          });
        }
      }
    }
    __syncwarp();
    for (int j = lane; j <= H; j += 32) {
#pragma unroll
      for (int q = 0; q < 4; q++) stg[q * L + j] = my.st[q][j];
    }
#pragma unroll
    for (int m2 = 0; m2 < 16; m2++) p.tail[(size_t)s * H + 32 * m2 + lane] = tail[m2];
  }
}

// true when the bulk-copy alignment rules hold (16-byte aligned base, strides multiples of 4 floats)
bool mcra_pairs_supported(const KernelParams& p) {
  if (getenv("BF_MCRA_OLD")) return false;   // tests: the CTA-per-stream kernel instead (read per launch)
  return p.H == 512 && (reinterpret_cast<uintptr_t>(p.in) & 15) == 0 && (p.in_stream_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(p.prev_hop) & 15) == 0;
}

cudaError_t launch_mcra_pairs(const KernelParams& p, cudaStream_t st, int sm_count) {
  const size_t smem = 1024 * sizeof(float2) + kMcWarps * sizeof(McWarp);
  cudaError_t e = cudaFuncSetAttribute(mcra_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int ctas = (p.n_streams + kMcWarps - 1) / kMcWarps;
  if (ctas > sm_count) ctas = sm_count;
  if (ctas < 1) ctas = 1;
  mcra_pairs_kernel<<<ctas, kMcWarps * 32, smem, st>>>(p);
  return cudaGetLastError();
}

}   // namespace bf
