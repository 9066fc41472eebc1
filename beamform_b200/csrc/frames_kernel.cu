// Fused per-stream beamforming kernel for 1024-point frames (hop 512), sm_100a.
//
// One CTA owns one stream and walks its hops in order, two frames (t, t+1) per iteration:
//   warps 1..W-1 : FORWARD   window -> packed complex FFT of every microphone -> Z[ch] in shared memory
//   all threads  : PHASE B   the per-bin beamformer of the selected node (das/mvdr/lcmv/gss/phase/phasempf)
//                            -> G = Yh_t + i*Yh_{t+1} in shared memory (Hermitian-ised half spectra)
//   warp 0       : INVERSE   packed inverse FFT -> synthesis window -> 50% overlap-add in registers -> out
// Warp 0 runs the inverse of pair p while warps 1.. run the forward of pair p+1 (two barriers per pair).
// Spectra never leave the SM; input samples are read once from HBM (the second touch of each hop hits
// L1/L2) and every output sample is written once.
//
// Reference path replaced (citations /root/reference/beamform/src/): util.h:217-253,289-314 (framing,
// window, OLA), das.cpp:47-70 and the apply_weights of every other node (see each phase_b_* below).
#include "bf_device.h"
#include "warp_fft1024.cuh"
#include "phase_b_select.cuh"
#include "phase_b_phase.cuh"

namespace bf {

constexpr int N1K = 1024;
constexpr int H1K = 512;

struct PairCtx {
  int t;        // first frame of the pair (hop index inside this call's arrays)
  bool two;     // frame t+1 exists
};

__device__ __forceinline__ const float* hop_ptr(const KernelParams& p, int s, int ch, int h) {
  if (h < 0) return p.prev_hop + ((size_t)s * p.M + ch) * p.H;   // the hop before this call (zeros initially)
  return p.in + (size_t)s * p.in_stream_stride + (size_t)ch * p.in_mic_stride + (size_t)h * p.H;
}

// FORWARD for one microphone: z[n] = 0.5*w[n]*(frame_t[n] + i*frame_{t+1}[n]), Z = FFT_1024(z) -> zbuf (linear [1024]).
// frame_t = [hop t-1 | hop t] (util.h:217-242: ring buffer holds previous + new hop).  The 0.5 makes the
// later even/odd separation X_t = Z[j] + conj(Z[N-j]) exact without a scale.
template <bool ENERGY>
__device__ __forceinline__ void forward_mic(const KernelParams& p, int s, int ch, PairCtx pc, float2* zbuf, const float2* tw,
                                            int lane, float s_l, float c_l, float* sqrtE0, float* sqrtE1) {
  const float* ha = hop_ptr(p, s, ch, pc.t - 1);
  const float* hb = hop_ptr(p, s, ch, pc.t);
  const float* hc = pc.two ? hop_ptr(p, s, ch, pc.t + 1) : hb;
  float a[16], b[16], c[16];
#pragma unroll
  for (int r = 0; r < 16; r++) {
    a[r] = __ldg(ha + 32 * r + lane);
    b[r] = __ldg(hb + 32 * r + lane);
    c[r] = pc.two ? __ldg(hc + 32 * r + lane) : 0.0f;
  }
  float2 v[32];
  static_for<0, 16>([&](auto r) {
    const float w0 = win1024<r>(s_l, c_l);        // w[32r + lane]
    const float w1 = win1024<r + 16>(s_l, c_l);   // w[32r + lane + 512]
    v[brev5(r)] = make_float2(a[r] * w0, b[r] * w0);
    v[brev5(r + 16)] = make_float2(b[r] * w1, c[r] * w1);
  });
  if (ENERGY) {   // sqrt of the windowed frame energies: scale of the FP32 FFT's absolute error (gate guard band)
    float e0 = 0.f, e1 = 0.f;
#pragma unroll
    for (int r = 0; r < 32; r++) { e0 = fmaf(v[r].x, v[r].x, e0); e1 = fmaf(v[r].y, v[r].y, e1); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { e0 += __shfl_xor_sync(0xffffffffu, e0, o); e1 += __shfl_xor_sync(0xffffffffu, e1, o); }
    if (lane == 0) { sqrtE0[ch] = 2.0f * sqrtf(e0); sqrtE1[ch] = 2.0f * sqrtf(e1); }
  }
  warp_fft1024<-1>(v, zbuf, tw, lane);
#pragma unroll
  for (int k2 = 0; k2 < 32; k2++) zbuf[k2 * 32 + lane] = v[k2];
}

// INVERSE + synthesis window + overlap-add (util.h:244-253, 301-302).  g = G[1024] in shared memory,
// y_t = Re(IFFT(G)), y_{t+1} = Im(IFFT(G)).  tail[] holds out_buff[0][j+H] of the previous frame for
// j = lane + 32*m2; each lane owns the same sample columns in every iteration, so the OLA never leaves
// registers.  out hop t = tail + y_t[:H]; out hop t+1 = y_t[H:] + y_{t+1}[:H].
template <bool SMOOTH>
__device__ __forceinline__ void inverse_pair(const KernelParams& p, int s, PairCtx pc, const float2* g, float2* tile,
                                             const float2* tw, int lane, float s_o, float c_o, float (&tail)[16], float* ola) {
  float2 v[32];
  static_for<0, 32>([&](auto j1) { v[brev5(j1)] = g[j1 * 32 + lane]; });
  __syncwarp();
  warp_fft1024<1>(v, tile, tw, lane);
  float* o0 = p.out + (size_t)s * p.out_stream_stride + (size_t)pc.t * p.H;
  float* o1 = o0 + p.H;
  const int S1 = SMOOTH ? p.smooth_size - 1 : 0;
  static_for<0, 16>([&](auto m2) {
    const float w0 = win1024<m2>(s_o, c_o);
    const float w1 = win1024<m2 + 16>(s_o, c_o);
    const float y0a = v[m2].x * w0, y0b = v[m2 + 16].x * w1;     // frame t: first / second half
    const float y1a = v[m2].y * w0, y1b = v[m2 + 16].y * w1;     // frame t+1
    const float r0 = tail[m2] + y0a;
    if (SMOOTH) ola[S1 + 32 * m2 + lane] = r0; else o0[32 * m2 + lane] = r0;
    if (pc.two) {
      const float r1 = y0b + y1a;
      if (SMOOTH) ola[S1 + 512 + 32 * m2 + lane] = r1; else o1[32 * m2 + lane] = r1;
      tail[m2] = y1b;
    } else {
      tail[m2] = y0b;
    }
  });
  if (SMOOTH) {
    // phasempf.cpp:78-83,122-130,331-334: every output sample becomes the mean of the last smooth_size
    // OLA samples (double accumulation, zero-initialised history carried in ola[0 .. smooth_size-2])
    __syncwarp();
    const int cnt = pc.two ? 1024 : 512, S = p.smooth_size;
    const double inv = 1.0 / (double)S;
    for (int n = lane; n < cnt; n += 32) {
      double acc = 0.0;
      for (int k = 0; k < S; k++) acc += (double)ola[n + k];
      o0[n] = (float)(acc * inv);
    }
    __syncwarp();
    float keep[2];
    for (int i = lane, q = 0; i < S1; i += 32, q++) keep[q] = ola[cnt + i];
    __syncwarp();
    for (int i = lane, q = 0; i < S1; i += 32, q++) ola[i] = keep[q];
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// PHASE B: das.cpp:60-63.  Y[j] = (1/M) sum_i conj(w_ij) X_i[j] is linear with frame-independent
// weights, so it commutes with the frame packing: G[j] = sum_i ceff_i[j] * Z_i[j] over all N bins,
// ceff = Hermitian part of conj(w)/M (host, double), which also folds the one asymmetric bin pair.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void phase_b_das(const KernelParams& p, const float2* zall, float2* g, int tid, int nthreads) {
  for (int j = tid; j < N1K; j += nthreads) {
    float2 acc = make_float2(0.f, 0.f);
    for (int ch = 0; ch < p.M; ch++) {
      float2 z = zall[ch * kXTile + j];
      float2 w = __ldg(p.das_ceff + (size_t)ch * N1K + j);
      acc.x = fmaf(z.x, w.x, acc.x); acc.x = fmaf(-z.y, w.y, acc.x);
      acc.y = fmaf(z.x, w.y, acc.y); acc.y = fmaf(z.y, w.x, acc.y);
    }
    g[j] = make_float2(2.0f * acc.x, 2.0f * acc.y);   // Z carries the 0.5 of the packing
  }
}

template <int ALGO>
__global__ void __launch_bounds__(288) frames_kernel_1024(const KernelParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tw = reinterpret_cast<float2*>(smem_raw);       // [32][32] W_1024^{lane*k1}
  float2* gbuf = tw + 1024;                                // G, then warp 0's exchange tile
  float2* zall = gbuf + kXTile;                            // [M] tiles: exchange tile, then Z linear
  SelScratch& sel = *reinterpret_cast<SelScratch*>(zall + (size_t)p.M * kXTile);
  PhaseScratch& phs = *reinterpret_cast<PhaseScratch*>(zall + (size_t)p.M * kXTile);
  constexpr bool kSel = (ALGO == ALGO_MVDR || ALGO == ALGO_LCMV || ALGO == ALGO_GSS);
  constexpr bool kPha = (ALGO == ALGO_PHASE || ALGO == ALGO_PHASEMPF);
  constexpr bool kSmooth = (ALGO == ALGO_PHASEMPF);
  float* sqrtE0 = kSel ? sel.sqrtE[0] : phs.sqrtE[0];
  float* sqrtE1 = kSel ? sel.sqrtE[1] : phs.sqrtE[1];
  int cur_L = p.mcra_cur_L0, first_L = p.mcra_first0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int s = blockIdx.x + p.stream_begin;

  for (int i = tid; i < 1024; i += blockDim.x) {
    int k1 = i >> 5, l = i & 31;
    float sn, cs;
    sincospif(-2.0f * (float)((k1 * l) & 1023) / 1024.0f, &sn, &cs);
    tw[i] = make_float2(cs, sn);
  }
  double sd, cd;
  sincospi((double)lane / N1K, &sd, &cd);
  const float s_l = (float)(0.5 * sd), c_l = (float)(0.5 * cd);                 // analysis window * 0.5
  const float s_o = (float)(sd * p.out_scale), c_o = (float)(cd * p.out_scale); // synthesis window * out_amp / N
  float tail[16];
  if (warp == 0) {
#pragma unroll
    for (int m2 = 0; m2 < 16; m2++) tail[m2] = p.tail[(size_t)s * p.H + 32 * m2 + lane];
    if (kSmooth)
      for (int i = lane; i < p.smooth_size - 1; i += 32) phs.ola[i] = p.smooth_hist[(size_t)s * 64 + i];
  }
  if (ALGO == ALGO_PHASEMPF)
    for (int i = tid; i < 7 * kL1K; i += blockDim.x) (&phs.state[0][0])[i] = p.mpf_state[(size_t)s * 7 * kL1K + i];
  __syncthreads();

  const int npairs = (p.hop_end - p.hop_begin + 1) / 2;
  for (int ip = 0; ip <= npairs; ip++) {
    PairCtx pc;
    pc.t = p.hop_begin + 2 * ip;
    pc.two = pc.t + 1 < p.hop_end;
    if (warp == 0) {
      if (ip > 0) {
        PairCtx pv;
        pv.t = pc.t - 2;
        pv.two = pv.t + 1 < p.hop_end;
        inverse_pair<kSmooth>(p, s, pv, gbuf, gbuf, tw, lane, s_o, c_o, tail, phs.ola);
      }
    } else if (ip < npairs) {
      for (int ch = warp - 1; ch < p.M; ch += nwarps - 1)
        forward_mic<kSel || kPha>(p, s, ch, pc, zall + ch * kXTile, tw, lane, s_l, c_l, sqrtE0, sqrtE1);
    }
    if (ip == npairs) break;
    __syncthreads();   // Z complete; previous G consumed
    if (ALGO == ALGO_DAS) phase_b_das(p, zall, gbuf, tid, blockDim.x);
    if (kSel) phase_b_select<ALGO>(p, s, pc.t, pc.two, zall, gbuf, sel, tid, blockDim.x);
    if (kPha) phase_b_phase<ALGO>(p, s, pc.t, pc.two, zall, gbuf, phs, cur_L, first_L, tid, blockDim.x);
    __syncthreads();   // G complete; Z consumed
  }
  if (warp == 0) {
#pragma unroll
    for (int m2 = 0; m2 < 16; m2++) p.tail[(size_t)s * p.H + 32 * m2 + lane] = tail[m2];
    if (kSmooth)
      for (int i = lane; i < p.smooth_size - 1; i += 32) p.smooth_hist[(size_t)s * 64 + i] = phs.ola[i];
  }
  if (ALGO == ALGO_PHASEMPF) {
    __syncthreads();
    for (int i = tid; i < 7 * kL1K; i += blockDim.x) p.mpf_state[(size_t)s * 7 * kL1K + i] = (&phs.state[0][0])[i];
  }
}

// Saves the last hop of the call as "previous hop" state for the next call (util.h ring buffer).
__global__ void save_prev_hop_kernel(const KernelParams p, int last_hop) {
  const int s = blockIdx.x + p.stream_begin, ch = blockIdx.y;   // streams on x: gridDim.y is limited to 65 535
  const float* src = p.in + (size_t)s * p.in_stream_stride + (size_t)ch * p.in_mic_stride + (size_t)last_hop * p.H;
  float* dst = p.prev_hop + ((size_t)s * p.M + ch) * p.H;
  for (int i = threadIdx.x; i < p.H; i += blockDim.x) dst[i] = src[i];
}

// gss.cpp:90-93: every update_weights resets sep_matrix[j] = weights[j]^H (adaptation is discarded)
__global__ void gss_reset_kernel(const KernelParams p) {
  // state layout [B][BF_GSS_ROWS][M][Lsel] (bin fastest); only the first C rows are live
  const size_t per = (size_t)p.C * p.M * p.Lsel;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)p.n_streams * per; i += (size_t)gridDim.x * blockDim.x) {
    const size_t sidx = i / per, r = i % per;
    const int slot = (int)(r % p.Lsel);
    const int cm = (int)(r / p.Lsel), c = cm / p.M, m = cm % p.M;
    const int l = p.sel_list[slot];
    const float2 a = p.steer[((size_t)l * p.C + c) * p.M + m];
    p.gss_w[sidx * BF_GSS_ROWS * p.M * p.Lsel + (size_t)cm * p.Lsel + slot] = make_float2(a.x, -a.y);
  }
}
cudaError_t launch_gss_reset(const KernelParams& p, cudaStream_t st) {
  gss_reset_kernel<<<296, 256, 0, st>>>(p);
  return cudaGetLastError();
}

size_t frames_kernel_smem(int M) { return sizeof(float2) * (1024 + (size_t)kXTile * (1 + M)) + (sizeof(SelScratch) > sizeof(PhaseScratch) ? sizeof(SelScratch) : sizeof(PhaseScratch)); }

cudaError_t launch_frames_kernel_1024(int algo, const KernelParams& p, cudaStream_t st) {
  const int fwd = p.M < 8 ? p.M : 8;
  const int threads = 32 * (1 + fwd);
  const size_t smem = frames_kernel_smem(p.M);
  void (*k)(KernelParams) = nullptr;
  switch (algo) {
    case ALGO_DAS: k = frames_kernel_1024<ALGO_DAS>; break;
    case ALGO_GSS: k = frames_kernel_1024<ALGO_GSS>; break;
    case ALGO_PHASE: k = frames_kernel_1024<ALGO_PHASE>; break;
    case ALGO_PHASEMPF: k = frames_kernel_1024<ALGO_PHASEMPF>; break;
    default: return cudaErrorNotSupported;
  }
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k<<<p.n_streams, threads, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_save_prev_hop(const KernelParams& p, int last_hop, cudaStream_t st) {
  dim3 grid(p.n_streams, p.M);
  save_prev_hop_kernel<<<grid, 128, 0, st>>>(p, last_hop);
  return cudaGetLastError();
}

}   // namespace bf
