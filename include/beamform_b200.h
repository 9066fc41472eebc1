/*
 * beamform_b200 — C ABI of the B200-native frequency-domain beamforming hot path.
 *
 * Drop-in boundary for balkce/beamform (citations relative to /root/reference/beamform/src/):
 * every entry point replaces one piece of the per-node interface the reference builds out of
 * globals + ROS/JACK callbacks.  Plain C types only; no CUDA or torch types in any signature
 * (a CUDA stream crosses as void*).  All functions return bf_status (0 = ok), never throw and
 * never abort — mirroring rosjack_create's 0/1 convention (rosjack/rosjack.cpp:109,149,221,278).
 * There is no CPU fallback: without a usable sm_100 device bf_create fails with BF_ERR_NO_DEVICE.
 */
#ifndef BEAMFORM_B200_H
#define BEAMFORM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BF_MAX_MICS 64
#define BF_MAX_INTERF 16

/* One node executable of the reference each (das.cpp, mvdr.cpp, lcmv.cpp, gss.cpp, phase.cpp, phasempf.cpp). */
typedef enum bf_algo {
  BF_ALGO_DAS = 0,
  BF_ALGO_MVDR = 1,
  BF_ALGO_LCMV = 2,
  BF_ALGO_GSS = 3,
  BF_ALGO_PHASE = 4,
  BF_ALGO_PHASEMPF = 5,
  BF_ALGO_MCRA = 6,   /* stand-alone MCRA noise-reduction node (mcra.cpp:62-155): first microphone only; launch keys
                         alphaS, alphaD, alphaD2, delta, L, out_amp, out_only_noise (launch/mcra.launch) */
  BF_ALGO_REF = 7,    /* rosjack_ref (jack_ref.cpp:19-60): window^2 overlap-add of the first microphone, the delay-matched
                         reference signal of the evaluation scripts */
  BF_ALGO_GSC = 8     /* generalized sidelobe canceller (gsc.cpp:54-197): per-microphone frequency-domain alignment, then a
                         time-domain NLMS with a power-normalised step; launch keys use_vad, vad_threshold, mu0, mu_max,
                         filter_size (launch/gsc.launch) */
} bf_algo;

typedef enum bf_status {
  BF_OK = 0,
  BF_ERR_INVALID = 1,     /* bad argument / unsupported frame size or mic count */
  BF_ERR_NO_DEVICE = 2,   /* no CUDA device of compute capability 10.x: there is no CPU path */
  BF_ERR_CUDA = 3,        /* a CUDA runtime call failed; see bf_last_error() */
  BF_ERR_ALLOC = 4,
  BF_ERR_IO = 5           /* config file could not be read */
} bf_status;

/*
 * Everything the reference reads from the ROS parameter server and from JACK, under the same
 * names.  bf_config_init() fills the getParam fall-backs of the chosen node
 * (mvdr.cpp:146-187, lcmv.cpp:170-219, gss.cpp:177-240, phase.cpp:166-191, phasempf.cpp:357-473,
 * gsc.cpp:206-258, mcra.cpp:177-226);
 * bf_config_load_yaml() reads beamform_config.yaml (util.h:52-113); bf_config_set() takes one
 * launch-file <rosparam> key (launch/xxx.launch).
 */
typedef struct bf_config {
  int32_t algo;                    /* bf_algo */
  double sample_rate;              /* rosjack_sample_rate (rosjack.cpp:134) */
  uint32_t hop;                    /* rosjack_window_size = JACK period; fft_win = 2*hop (util.h:261) */
  int32_t n_mics;                  /* number of micN entries found (util.h:82-92) */
  double mic_x[BF_MAX_MICS];       /* RAW yaml coordinates; mic0 re-referencing happens inside (util.h:116-119) */
  double mic_y[BF_MAX_MICS];
  double initial_angle;            /* util.h:68 */
  int32_t n_angle_interf;          /* angle_interf1..; ingestion stops at the first |a| > 180 (util.h:101-112) */
  double angle_interf[BF_MAX_INTERF];
  uint32_t past_windows;           /* mvdr/lcmv */
  double freq_mag_threshold, freq_max, freq_min, out_amp, interf_angle_threshold;
  double mu, lambda;               /* gss */
  double min_phase, mag_mult, mag_threshold;   /* phase */
  double min_mag;                  /* phasempf */
  int32_t smooth_size;
  double MCRA_alphaS, MCRA_alphaD, MCRA_alphaD2, MCRA_delta;
  int32_t MCRA_L;
  double MPF_alphaS, MPF_eta, MPF_rev_gamma, MPF_rev_delta;
  double noise_floor;
  int32_t out_only_noise, out_only_mcra;
  /* offline-driver knobs (no reference counterpart) */
  int32_t dropped_hops_on_restructure; /* hops lost while READY=false after an interference add/remove
                                          (lcmv.cpp:271-276 sleeps 30 ms); default 0 */
  int32_t device;                      /* CUDA device ordinal */
  /* gsc (gsc.cpp:206-258); appended so that older callers' layouts stay valid */
  int32_t use_vad;
  double vad_threshold, mu0, mu_max;
  int32_t filter_size;                 /* NLMS taps per blocking-matrix channel: multiple of 32, <= 256 */
} bf_config;

/* The two control topics, scheduled offline: applied atomically BEFORE hop `hop_index` is processed. */
typedef struct bf_event {
  uint32_t hop_index;
  int32_t kind;      /* 0: /theta  std_msgs/Float32 (das.cpp:94-99); 1: /theta_interference InterfTheta (lcmv.cpp:258-309) */
  uint32_t id;       /* InterfTheta.id (uint16 on the wire) */
  float value;       /* angle in degrees (float32 on the wire) */
} bf_event;

typedef struct bf_handle bf_handle;

/* --- configuration ------------------------------------------------------------------------- */
int bf_config_init(bf_config* cfg, int algo);
int bf_config_load_yaml(bf_config* cfg, const char* path);            /* beamform_config.yaml */
int bf_config_set(bf_config* cfg, const char* key, const char* value); /* one rosparam key */
int bf_config_load_launch(bf_config* cfg, const char* path);           /* launch/<node>.launch: the inline <rosparam> block / <param> tags */

/* --- lifecycle: replaces handle_params + <algo>_handle_params + prepare_overlap_and_add +
 *     fftw_plan_dft_1d + update_weights(true) in every node's main() (das.cpp:102-140) -------- */
int bf_create(bf_handle** out, const bf_config* cfg, uint32_t n_streams);
void bf_destroy(bf_handle* h);

/* --- control: theta_roscallback (das.cpp:94-99) / interf_theta_roscallback (lcmv.cpp:258-309,
 *     gss.cpp:288-339).  Thread-safe: the setters queue the message, the processing thread applies it at the
 *     next hop boundary; the getters are read-only and report the state as of the last processed hop. ------ */
int bf_set_theta(bf_handle* h, float angle_deg);
int bf_set_interference(bf_handle* h, uint16_t id, float angle_deg);
int bf_get_theta(bf_handle* h, double* angle_deg);
int bf_get_interferences(bf_handle* h, double* angles, uint32_t cap, uint32_t* n);

/* --- the process callback: jack_callback(nframes, arg) of every node (das.cpp:72-92), i.e.
 *     input_from_rosjack -> do_overlap(apply_weights) -> output_to_rosjack.  `in[m]` are the M
 *     JACK port buffers (host, nframes floats each), `out` is the output port buffer (host).
 *     One hop in, one hop out, one hop of latency; requires n_streams == 1. --------------------- */
int bf_process_hop(bf_handle* h, const float* const* in, float* out, uint32_t nframes);

/* --- the per-frame operator seam: `void (*weight_func)(jack_ringbuffer_t **in, rosjack_data *out)` handed to
 *     do_overlap (util.h:289-314), i.e. apply_weights of every node (das.cpp:47-70, mvdr.cpp:62-115, ...).
 *     `in_frames[m]` = the fft_win samples of microphone m the ring buffer holds ([previous hop | new hop]),
 *     `out_frame` = fft_win windowed output samples (util.h:244-253); the caller overlap-adds (and, for phasempf,
 *     smooths: phasempf.cpp:331-334).  Advances the node's per-frame state (histories, W, MCRA) by one frame.
 *     n_streams == 1; not for gsc / rosjack_ref; do not mix with bf_process_* on one handle. ------------------- */
int bf_apply_weights(bf_handle* h, const float* const* in_frames, float* out_frame, uint32_t fft_win);

/* --- offline batched driver: the same callback applied to n_hops consecutive hops of n_streams
 *     independent streams.  Host variant copies H2D/D2H itself; the device variant takes device
 *     pointers and a CUDA stream (cudaStream_t as void*).  Layout: in[s*in_stream_stride +
 *     m*in_mic_stride + t*hop + j], out[s*out_stream_stride + t*hop + j] (strides in floats). ---- */
int bf_process_batch(bf_handle* h, const float* in_host, size_t in_stream_stride, size_t in_mic_stride, float* out_host,
                     size_t out_stream_stride, uint32_t n_hops, const bf_event* events, uint32_t n_events);
int bf_process_batch_device(bf_handle* h, const float* in_dev, size_t in_stream_stride, size_t in_mic_stride,
                            float* out_dev, size_t out_stream_stride, uint32_t n_hops, const bf_event* events,
                            uint32_t n_events, void* cuda_stream);

/* --- diagnostics for the bit-exact gates: when set, every processed frame writes one byte per
 *     FFT bin j in [0, fft_win): bit0 = magnitude gate passed inside the band (mvdr.cpp:84-85),
 *     bit1 = phase mask kept the bin (phase.cpp:114, phasempf.cpp:234).
 *     Device buffer, layout [n_streams][n_hops][fft_win]; NULL disables. ------------------------ */
int bf_set_capture(bf_handle* h, uint8_t* dev_flags);

/* --- steered-response sweep (config C5; the reference DAS response das.cpp:41,61-62 evaluated for
 *     n_dirs look directions): maps[s][t][d] = sum_j |(1/M) w_d(:,j)^H X(:,j)|^2, device pointers. - */
int bf_srp_batch_device(bf_handle* h, const float* in_dev, size_t in_stream_stride, size_t in_mic_stride,
                        const float* thetas_deg_host, uint32_t n_dirs, float* maps_dev, uint32_t n_hops, void* cuda_stream);

/* --- measurement: when enabled, CUDA events bracket every launch of the fused frames kernel on the
 *     launching stream; bf_get_profile() synchronises them and returns the summed device time (ms) and
 *     the number of launches since the last call. -------------------------------------------------- */
int bf_set_profiling(bf_handle* h, int enabled);
int bf_get_profile(bf_handle* h, double* kernel_ms, uint64_t* kernel_launches);

/* --- introspection ------------------------------------------------------------------------- */
uint32_t bf_fft_win(const bf_handle* h);
uint64_t bf_kernel_launches(const bf_handle* h);   /* kernels launched so far by this handle */
const char* bf_last_error(void);                   /* thread-local message for the last non-zero status */
const char* bf_version(void);

#ifdef __cplusplus
}
#endif
#endif
