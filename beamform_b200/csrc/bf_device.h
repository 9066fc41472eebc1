// Host<->device parameter block shared by capi.cu and the kernels.  Plain data only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define BF_MAX_MICS_DEV 64
#define BF_GSS_ROWS 16   /* rows of the gss separation-matrix state: look direction + up to 15 interferers (beamform_config.yaml:43-57) */

namespace bf {

enum { ALGO_DAS = 0, ALGO_MVDR = 1, ALGO_LCMV = 2, ALGO_GSS = 3, ALGO_PHASE = 4, ALGO_PHASEMPF = 5, ALGO_MCRA = 6, ALGO_REF = 7, ALGO_GSC = 8 };

// "Logical bins": the reference loops over all N FFT bins (das.cpp:60); inputs are real so bin N-j is
// the conjugate of bin j and every per-bin rule is conjugate-equivariant EXCEPT at the pair
// (N/2-1, N/2+1), whose frequencies differ (util.h:198, SURVEY B-1..B-4).  The kernels therefore
// process l = 0..N/2 plus one pseudo-bin l = N/2+1 (input = conj of bin N/2-1, own steering, own state).
// All per-bin tables are indexed by l in [0, N/2+2).
struct KernelParams {
  // ---- I/O (floats; strides in floats) ----
  const float* in;
  long long in_stream_stride, in_mic_stride;
  float* out;
  long long out_stream_stride;
  int n_streams, M, H, N;
  int stream_begin;         // first stream of this launch (grid = n_streams CTAs, stream = blockIdx.x + stream_begin)
  int hop_begin, hop_end;   // this launch processes hops [hop_begin, hop_end) of the arrays above
  int frame_index0;         // global frame counter of hop 0 of this call (history / MCRA bookkeeping)
  // ---- per-stream state carried between launches / calls ----
  float* prev_hop;          // [B][M][H]  the hop before hop 0 of this call (zeros at start: util.h:275-277)
  float* tail;              // [B][H]     second half of the last synthesised frame (out_buff[0], util.h:301-302)
  float* tail_out;          // [B][H]     where a launch leaves the new tails: == tail for kernels whose CTA owns a whole stream; a second
                            //            buffer for das_pairs_kernel, whose warps cut streams at arbitrary pairs (reader and writer of a
                            //            stream's tail may be different warps: the host swaps the two buffers after the segment)
  // ---- tables ----
  const float2* steer;      // [L][C][M]  weights[j](i,k): steering vectors, look direction k=0, interferers k>=1
  const float2* das_ceff;   // [M][N]     DAS only: Hermitian-ised effective weights (see capi.cu: build_das_ceff)
  const uint8_t* inband;    // [L]        freq_min <= |freqs[j]| <= freq_max (mvdr.cpp:84)
  int C;                    // K+1 columns of the steering matrix
  int sel_chunk;            // general gated kernel: microphones transformed per pass (< M: spectra spill to sel_ws)
  float2* sel_ws;           // [B][M][N] spectra workspace of the general gated kernel when they do not fit shared memory, else null
  int das_chunk;            // das, frame-size-generic kernel: microphones transformed per pass through shared memory (<= M)
  float out_scale;          // out_amp / N
  // ---- diagnostics ----
  uint8_t* capture;         // [B][n_hops_call][N] or null
  long long capture_stream_stride;
  // ---- algorithm parameters (float copies of bf_config) ----
  float thr_mag;            // freq_mag_threshold * M * N   (gate compares sum_i |X_i| directly)
  double thr_mag_d;         // exact double threshold for the FP64 recheck
  int P;                    // past_windows
  float2* hist;             // [B][D][M][Lsel]  mvdr/lcmv history ring (bin fastest), D = ring_depth = P + 2, slot = frame % D
  int debug;                // profiling experiments only (env BF_DEBUG): 0 in production
  int ring_depth, ring_slot0;   // ring_slot0 = slot of hop 0 of this launch
  const int* sel_slot;      // [L] -> slot in the in-band compact list or -1
  const int* sel_list;      // [Lsel] -> logical bin
  int Lsel;
  float mu, lambda_mu;      // gss: mu, (1 - lambda*mu)
  float2* gss_w;            // [B][BF_GSS_ROWS][M][Lsel] (bin fastest)
  float gss_dj2_scale;      // 2 * (1/(K+1)) in integer arithmetic (gss.cpp:133) -> 2 for K=0 else 0
  float min_phase_rad, mag_mult, thr_phase_mag;   // phase
  float min_mag;            // phasempf
  const float* win_f;       // [N] sqrt-hann (float), frame-size-generic kernel
  const float2* twid_f;     // [N] e^{-2 pi i k/N} (float), frame-size-generic kernel
  const double* win_d;      // [N] sqrt-hann in double, for FP64 rechecks
  const double2* twid_d;    // [N] e^{-2 pi i k/N} in double, for FP64 rechecks
  const double2* steer_d;   // [L][M] look-direction steering in double (phase family rechecks)
  double mag_threshold_d, min_phase_rad_d;
  // phasempf state, [B][7][L] floats: S_prev, S_tmp, S_min, lambda_noise, Z, rev0, rev1
  float* mpf_state;
  float mcra_alphaS, mcra_alphaD, mcra_alphaD2, mcra_delta;
  int mcra_L, mcra_cur_L0, mcra_first0;   // counters at hop_begin (advance deterministically per frame)
  float mpf_alphaS, mpf_eta, mpf_gamma, mpf_rev_gain, out_amp, noise_floor;
  int out_only_noise, out_only_mcra;
  int smooth_size;
  float* smooth_hist;       // [B][smooth_size-1] last OLA samples before hop 0 of this call
  // ---- gsc (gsc.cpp:93-197) ----
  float* gsc_aligned;       // [B][M][hops*H] workspace: per-microphone aligned signals of this launch (do_overlap_bymic output)
  long long gsc_aligned_stream_stride;   // floats; microphone stride = hops*H of this launch
  float* gsc_tail;          // [B][M][H] per-microphone overlap-add tails (out_buff_mic, util.h:336-344)
  float* gsc_state;         // [B][(2(M-1)+1)][F]: blocking-matrix delay lines, NLMS filters, last outputs (rings: see gsc_head)
  int* gsc_head;            // [B] ring position of logical tap 0
  int gsc_F, gsc_use_vad;   // filter_size, use_vad
  double gsc_vad_threshold, gsc_mu0, gsc_mu_max;
};

}   // namespace bf
