// TEST INFRASTRUCTURE — see the header of bf_oracle.hpp.  Flat C ABI over bfo::Beamformer so
// tests/ and bench.py (cpu_baseline / --impl reference legs only) can drive the CPU restatement
// through ctypes.  Nothing in beamform_b200/ may link or load this library.
#include "bf_oracle.hpp"

extern "C" {

#define BFO_MAX_MICS 64
#define BFO_MAX_INTERF 16

typedef struct bfo_config {
  int32_t algo;
  double sample_rate;
  uint32_t hop;
  int32_t n_mics;
  double mic_x[BFO_MAX_MICS], mic_y[BFO_MAX_MICS];
  double initial_angle;
  int32_t n_angle_interf;
  double angle_interf[BFO_MAX_INTERF];
  uint32_t past_windows;
  double freq_mag_threshold, freq_max, freq_min, out_amp, interf_angle_threshold;
  double mu, lambda;
  double min_phase, mag_mult, mag_threshold;
  double min_mag;
  int32_t smooth_size;
  double MCRA_alphaS, MCRA_alphaD, MCRA_alphaD2, MCRA_delta;
  int32_t MCRA_L;
  double MPF_alphaS, MPF_eta, MPF_rev_gamma, MPF_rev_delta;
  double noise_floor;
  int32_t out_only_noise, out_only_mcra;
  int32_t use_vad;
  double vad_threshold, mu0, mu_max;
  int32_t filter_size;
} bfo_config;

typedef struct bfo_event {
  uint32_t hop_index;   // applied atomically before this hop is processed (B-13)
  int32_t kind;         // 0: /theta (std_msgs/Float32), 1: /theta_interference (InterfTheta.msg)
  uint32_t id;          // uint16 on the wire
  float value;          // float32 on the wire
} bfo_event;

static bfo::Config to_cfg(const bfo_config* c) {
  bfo::Config k;
  k.algo = c->algo; k.sample_rate = c->sample_rate; k.hop = c->hop;
  for (int i = 0; i < c->n_mics; i++) { k.mic_x.push_back(c->mic_x[i]); k.mic_y.push_back(c->mic_y[i]); }
  k.initial_angle = c->initial_angle;
  for (int i = 0; i < c->n_angle_interf; i++) k.angle_interf.push_back(c->angle_interf[i]);
  k.past_windows = c->past_windows; k.freq_mag_threshold = c->freq_mag_threshold; k.freq_max = c->freq_max;
  k.freq_min = c->freq_min; k.out_amp = c->out_amp; k.interf_angle_threshold = c->interf_angle_threshold;
  k.mu = c->mu; k.lambda = c->lambda; k.min_phase = c->min_phase; k.mag_mult = c->mag_mult;
  k.mag_threshold = c->mag_threshold; k.min_mag = c->min_mag;
  k.smooth_size = c->smooth_size < 1 ? 20 : c->smooth_size;   // phasempf.cpp:377-381
  k.MCRA_alphaS = c->MCRA_alphaS; k.MCRA_alphaD = c->MCRA_alphaD; k.MCRA_alphaD2 = c->MCRA_alphaD2;
  k.MCRA_delta = c->MCRA_delta; k.MCRA_L = c->MCRA_L; k.MPF_alphaS = c->MPF_alphaS; k.MPF_eta = c->MPF_eta;
  k.MPF_rev_gamma = c->MPF_rev_gamma; k.MPF_rev_delta = c->MPF_rev_delta; k.noise_floor = c->noise_floor;
  k.out_only_noise = c->out_only_noise != 0; k.out_only_mcra = c->out_only_mcra != 0;
  k.use_vad = c->use_vad != 0; k.vad_threshold = c->vad_threshold; k.mu0 = c->mu0; k.mu_max = c->mu_max; k.filter_size = c->filter_size;
  return k;
}

void* bfo_create(const bfo_config* c) { return new bfo::Beamformer(to_cfg(c)); }
void bfo_destroy(void* h) { delete (bfo::Beamformer*)h; }
uint32_t bfo_fft_win(void* h) { return ((bfo::Beamformer*)h)->fft_win; }

void bfo_set_theta(void* h, float deg) { ((bfo::Beamformer*)h)->theta_roscallback(deg); }
int bfo_set_interference(void* h, uint16_t id, float deg) { return ((bfo::Beamformer*)h)->interf_theta_roscallback(id, deg); }
int bfo_get_interferences(void* h, double* out, int cap) {
  bfo::Beamformer* b = (bfo::Beamformer*)h;
  int n = (int)b->interference_angles.size();
  for (int i = 0; i < n && i < cap; i++) out[i] = b->interference_angles[i];
  return n;
}
void bfo_get_freqs(void* h, double* out) { bfo::Beamformer* b = (bfo::Beamformer*)h; for (unsigned j = 0; j < b->fft_win; j++) out[j] = b->freqs[j]; }
void bfo_get_delays(void* h, double* out) { bfo::Beamformer* b = (bfo::Beamformer*)h; for (int i = 0; i < b->number_of_microphones; i++) out[i] = b->delays[i]; }
void bfo_get_window(void* h, double* out) { bfo::Beamformer* b = (bfo::Beamformer*)h; for (unsigned j = 0; j < b->fft_win; j++) out[j] = b->hann_win[j]; }
// weights[j](i,k) as interleaved re,im doubles, layout [N][M][K+1]; returns K+1
int bfo_get_weights(void* h, double* out) {
  bfo::Beamformer* b = (bfo::Beamformer*)h;
  int C = b->weights[0].c, M = b->number_of_microphones;
  if (out)
    for (unsigned j = 0; j < b->fft_win; j++)
      for (int i = 0; i < M; i++)
        for (int k = 0; k < C; k++) { out[2 * ((j * M + i) * C + k)] = b->weights[j](i, k).real(); out[2 * ((j * M + i) * C + k) + 1] = b->weights[j](i, k).imag(); }
  return C;
}

// One hop through jack_callback (das.cpp:72-92): in[m] -> hop floats, out -> hop floats.
int bfo_process_hop(void* h, const float* const* in, float* out, uint32_t nframes) {
  return ((bfo::Beamformer*)h)->jack_callback(in, out, nframes);
}

// Offline driver: n_hops hops of planar input in[m*mic_stride + t*hop + j] -> out[t*hop + j].
// sel (optional) receives last_selected per hop ([n_hops][N]); mask likewise for last_mask.
// dropped_hops: hops lost while READY=false after a list restructure (lcmv.cpp:271-276); 0 by default.
int bfo_process(void* h, const float* in, size_t mic_stride, float* out, uint32_t n_hops, const bfo_event* ev, int n_ev,
                int dropped_hops, uint8_t* sel, uint8_t* mask) {
  bfo::Beamformer* b = (bfo::Beamformer*)h;
  const unsigned H = b->hop, N = b->fft_win;
  std::vector<const float*> ptr(b->number_of_microphones);
  int e = 0, drop = 0;
  for (uint32_t t = 0; t < n_hops; t++) {
    while (e < n_ev && ev[e].hop_index <= t) {
      if (ev[e].kind == 0) b->theta_roscallback(ev[e].value);
      else if (b->uses_interf() && b->interf_theta_roscallback((uint16_t)ev[e].id, ev[e].value)) drop = dropped_hops;
      e++;
    }
    for (int m = 0; m < b->number_of_microphones; m++) ptr[m] = in + (size_t)m * mic_stride + (size_t)t * H;
    if (drop > 0) { b->READY = false; drop--; } else b->READY = true;
    std::fill(b->last_selected.begin(), b->last_selected.end(), 0);
    std::fill(b->last_mask.begin(), b->last_mask.end(), 0);
    b->jack_callback(ptr.data(), out + (size_t)t * H, H);
    if (sel) std::memcpy(sel + (size_t)t * N, b->last_selected.data(), N);
    if (mask) std::memcpy(mask + (size_t)t * N, b->last_mask.data(), N);
  }
  b->READY = true;
  return 0;
}

// Steered-response sweep (config C5): maps[t*D + d], frame t = [hop t-1 | hop t].
int bfo_srp(void* h, const float* in, size_t mic_stride, uint32_t n_hops, const double* thetas, int D, double* maps) {
  bfo::Beamformer* b = (bfo::Beamformer*)h;
  const unsigned H = b->hop;
  for (uint32_t t = 0; t < n_hops; t++) {
    for (int m = 0; m < b->number_of_microphones; m++)
      std::memcpy(&b->in_buff[m][H], in + (size_t)m * mic_stride + (size_t)t * H, sizeof(float) * H);
    b->srp_frame(thetas, D, maps + (size_t)t * D);
    for (int m = 0; m < b->number_of_microphones; m++) std::memmove(&b->in_buff[m][0], &b->in_buff[m][H], sizeof(float) * H);
  }
  return 0;
}

}   // extern "C"
