// C ABI + host runtime of beamform_b200 (see include/beamform_b200.h).
//
// Host side of the drop-in boundary: configuration ingest (util.h:52-134 + <algo>_handle_params),
// geometry -> delays -> steering tables in double (util.h:136-199, das.cpp:27-45 and its five copies),
// the /theta and /theta_interference state machine (lcmv.cpp:258-309), event-segmented launches of the
// fused CUDA kernels, and per-stream state.  There is no CPU compute path: every sample goes through the
// sm_100a kernels in frames_kernel.cu.  Citations are relative to /root/reference/beamform/src/.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/beamform_b200.h"
#include "bf_device.h"

namespace bf {
cudaError_t launch_frames_kernel_1024(int algo, const KernelParams& p, cudaStream_t st);
cudaError_t launch_save_prev_hop(const KernelParams& p, int last_hop, cudaStream_t st);
cudaError_t launch_gss_reset(const KernelParams& p, cudaStream_t st);
size_t frames_kernel_smem(int M);
bool das_pairs_supported(const KernelParams& p);
cudaError_t launch_das_pairs(const KernelParams& p, cudaStream_t st, int sm_count);
bool sel_pairs_supported(const KernelParams& p, int algo);
cudaError_t launch_sel_pairs(int algo, const KernelParams& p, cudaStream_t st);
bool sel_stream_supported(const KernelParams& p, int algo, const uint8_t* inband_host);
cudaError_t launch_sel_stream(int algo, const KernelParams& p, cudaStream_t st);
cudaError_t launch_frames_kernel_n(int algo, const KernelParams& p, cudaStream_t st);
size_t frames_kernel_n_smem(int N, int M, int algo);
int frames_kernel_n_das_chunk(int N, int M);
int frames_kernel_sel_chunk(int N, int M);
cudaError_t launch_frames_kernel_sel(int algo, const KernelParams& p, cudaStream_t st);
size_t frames_kernel_sel_smem(int N, int M);
cudaError_t launch_frames_kernel_mcra(const KernelParams& p, cudaStream_t st);
bool mcra_pairs_supported(const KernelParams& p);
cudaError_t launch_mcra_pairs(const KernelParams& p, cudaStream_t st, int sm_count);
bool phase_n_supported(const KernelParams& p, int algo);
cudaError_t launch_phase_n(int algo, const KernelParams& p, cudaStream_t st);
cudaError_t launch_ref_kernel(const KernelParams& p, cudaStream_t st);
cudaError_t launch_gsc(const KernelParams& p, cudaStream_t st);
size_t gsc_align_smem(int N, int M);
cudaError_t launch_srp(const KernelParams& p, unsigned char* xi, const double* tau, const double* freqs_l, float* maps, int D, int n_hops,
                       cudaStream_t st);
size_t srp_workspace_bytes(long long F);
}   // namespace bf

typedef std::complex<double> cd;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(x)                                                                                  \
  do {                                                                                               \
    cudaError_t e_ = (x);                                                                            \
    if (e_ != cudaSuccess) return fail(BF_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// util.h:23-27
static const double kPi = 3.141592653589793238462643383279502884;
static const double kVSound = 343;
static const double kRad2Deg = 180.0 / kPi;
static const double kDeg2Rad = kPi / 180.0;

static const size_t kMaxInterferers = 15;   // beamform_config.yaml:43-57 ships 15 slots; the kernels carry look direction + 15 columns

struct PendingEvent {
  int kind;
  uint16_t id;
  float value;
};

struct bf_handle {
  bf_config cfg;
  uint32_t B = 0, M = 0, H = 0, N = 0, L = 0;
  int dev = 0;
  cudaStream_t own_stream = nullptr;
  // ---- reference globals (util.h:31-36) ----
  double angle = 0;
  std::vector<double> mic_dist, mic_angle;
  std::vector<double> interference_angles;
  std::vector<double> freqs;
  // weights[j](i,k): [N][M][C]; row-0 semantics of update_weights(ini) are kept (SURVEY B-8)
  std::vector<cd> weights;
  int C = 1;
  // ---- device ----
  float* d_prev_hop = nullptr;
  float* d_tail = nullptr;
  float* d_tail2 = nullptr;   // das: second tail buffer (das_pairs_kernel reads one and writes the other, swapped per segment)
  int sm_count = 148;
  float2* d_steer = nullptr;
  float2* d_das_ceff = nullptr;
  float *d_gsc_aligned = nullptr, *d_gsc_tail = nullptr, *d_gsc_state = nullptr;   // gsc: workspace, per-mic OLA tails, delay lines + filters
  int* d_gsc_head = nullptr;
  size_t gsc_aligned_cap = 0;
  uint8_t* d_inband = nullptr;
  std::vector<uint8_t> inband_host;   // copy of the device table (kernel selection)
  uint8_t* d_capture = nullptr;
  size_t steer_cap = 0;
  int* d_sel_slot = nullptr;
  int* d_sel_list = nullptr;
  int Lsel = 0;
  float2* d_hist = nullptr;     // mvdr/lcmv: [B][Lsel][P+2][M]
  float2* d_gss_w = nullptr;    // gss: [B][Lsel][C][M]
  float2* d_sel_ws = nullptr;   // general gated kernel: [B][M][N] spectra workspace when they do not fit shared memory
  // steered-response sweep workspace
  unsigned char* d_srp_xs = nullptr;   // operand images [514][frame tiles][bf16 hi | lo tile] (srp_kernel.cu)
  size_t srp_xs_cap = 0;
  double* d_srp_tau = nullptr;
  size_t srp_tau_cap = 0;
  double* d_srp_freqs = nullptr;
  float* d_win_f = nullptr;
  float2* d_twid_f = nullptr;
  double* d_win_d = nullptr;
  double2* d_twid_d = nullptr;
  double2* d_steer_d = nullptr;   // phase family: [L][M] look-direction steering in double
  float* d_mpf_state = nullptr;   // phasempf: [B][7][L]
  float* d_smooth_hist = nullptr; // phasempf: [B][64]
  int mcra_cur_L = 0, mcra_first = 1;   // phasempf.cpp:47-48
  bool gss_reset_pending = true;
  // host-batch staging + copy/compute overlap
  float* d_io_in = nullptr;
  float* d_io_out = nullptr;
  size_t io_in_cap = 0, io_out_cap = 0;
  cudaStream_t st_h2d = nullptr, st_d2h = nullptr;
  cudaEvent_t ev_in[8] = {nullptr}, ev_k[8] = {nullptr};
  // hop-at-a-time staging
  float* h_stage_in = nullptr;
  float* h_stage_out = nullptr;
  float* d_stage_in = nullptr;
  float* d_stage_out = nullptr;
  // ---- bookkeeping ----
  uint64_t frames_done = 0;
  uint64_t launches = 0;
  int drop_left = 0;
  bool tables_dirty = true;
  bool profiling = false;
  bool raw_frame_mode = false;   // bf_apply_weights: the smoother of phasempf stays with the caller (phasempf.cpp:331-334)
  std::vector<std::pair<cudaEvent_t, cudaEvent_t> > prof_events;
  std::mutex mtx;         // guards `pending` (setters may run on another thread, like the ROS callbacks)
  std::mutex state_mtx;   // guards angle / interference_angles against the read-only getters
  std::vector<PendingEvent> pending;
};

// ------------------------------------------------------------------------------------------------
// configuration
// ------------------------------------------------------------------------------------------------
extern "C" int bf_config_init(bf_config* c, int algo) {
  if (!c || algo < 0 || algo > 8) return fail(BF_ERR_INVALID, "bf_config_init: bad arguments");
  memset(c, 0, sizeof(*c));
  c->algo = algo;
  c->sample_rate = 48000;   // rosjack_config.yaml:9 (JACK decides at run time)
  c->hop = 512;
  c->initial_angle = 0.0;   // util.h:71
  // getParam fall-backs: mvdr.cpp:155-184, lcmv.cpp:179-216, gss.cpp:186-237
  c->past_windows = 10;
  c->freq_mag_threshold = 1.5;
  c->freq_max = 4000;
  c->freq_min = 400;
  c->out_amp = (algo == BF_ALGO_PHASEMPF || algo == BF_ALGO_MCRA) ? 2.0 : 4.5;   // phasempf.cpp:451, mcra.cpp:215
  c->interf_angle_threshold = 5.0;
  c->mu = 0.01;      // gss.cpp fall-back
  c->lambda = 0.0;
  // phase.cpp:170-189
  c->min_phase = 10.0;
  c->mag_mult = 0.1;
  c->mag_threshold = 0.05;
  // phasempf.cpp:361-470 (fall-backs, not the global initialisers; SURVEY B-11)
  c->min_mag = 10.0;
  c->smooth_size = 20;
  c->MCRA_alphaS = 0.95; c->MCRA_alphaD = 0.95; c->MCRA_alphaD2 = 0.97; c->MCRA_delta = 0.001;
  c->MCRA_L = 0;     // "MCRA_L = 0.01" assigned to an int (phasempf.cpp:416)
  c->MPF_alphaS = 0.3; c->MPF_eta = 0.3; c->MPF_rev_gamma = 0.3; c->MPF_rev_delta = 1.0;
  c->noise_floor = 0.001;
  c->out_only_noise = (algo == BF_ALGO_MCRA) ? 1 : 0;   // mcra.cpp:222: the fall-back is `true`
  c->out_only_mcra = 0;
  c->dropped_hops_on_restructure = 0;
  c->device = 0;
  // gsc.cpp:206-258
  c->use_vad = 0; c->vad_threshold = 0.1; c->mu0 = 0.0005; c->mu_max = 0.01; c->filter_size = 128;
  return BF_OK;
}

static std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
static bool parse_bool(const std::string& v) { return v == "true" || v == "True" || v == "1"; }

extern "C" int bf_config_set(bf_config* c, const char* key_c, const char* val_c) {
  if (!c || !key_c || !val_c) return fail(BF_ERR_INVALID, "bf_config_set: null argument");
  const std::string key = trim(key_c), val = trim(val_c);
  const double v = atof(val.c_str());
#define KEYD(name) if (key == #name) { c->name = v; return BF_OK; }
  KEYD(freq_mag_threshold) KEYD(freq_max) KEYD(freq_min) KEYD(out_amp) KEYD(interf_angle_threshold) KEYD(mu) KEYD(lambda)
  KEYD(min_phase) KEYD(mag_mult) KEYD(mag_threshold) KEYD(min_mag) KEYD(MCRA_alphaS) KEYD(MCRA_alphaD) KEYD(MCRA_alphaD2)
  KEYD(MCRA_delta) KEYD(MPF_alphaS) KEYD(MPF_eta) KEYD(MPF_rev_gamma) KEYD(MPF_rev_delta) KEYD(noise_floor)
  KEYD(initial_angle) KEYD(sample_rate) KEYD(vad_threshold) KEYD(mu0) KEYD(mu_max)
#undef KEYD
  if (key == "use_vad") { c->use_vad = parse_bool(val); return BF_OK; }
  if (key == "filter_size") { c->filter_size = (int)v; return BF_OK; }
  if (key == "past_windows") { c->past_windows = (uint32_t)(int)v; return BF_OK; }   // mvdr.cpp:152 (int) cast
  if (key == "smooth_size") { c->smooth_size = (int)v < 1 ? 20 : (int)v; return BF_OK; }   // phasempf.cpp:377-381
  if (key == "MCRA_L") { c->MCRA_L = (int)v; return BF_OK; }
  // the stand-alone mcra node names the same quantities without the prefix (mcra.cpp:181-224)
  if (key == "alphaS") { c->MCRA_alphaS = v; return BF_OK; }
  if (key == "alphaD") { c->MCRA_alphaD = v; return BF_OK; }
  if (key == "alphaD2") { c->MCRA_alphaD2 = v; return BF_OK; }
  if (key == "delta") { c->MCRA_delta = v; return BF_OK; }
  if (key == "L") { c->MCRA_L = (int)v; return BF_OK; }
  if (key == "hop" || key == "period") { c->hop = (uint32_t)v; return BF_OK; }
  if (key == "out_only_noise") { c->out_only_noise = parse_bool(val); return BF_OK; }
  if (key == "out_only_mcra") { c->out_only_mcra = parse_bool(val); return BF_OK; }
  if (key == "dropped_hops_on_restructure") { c->dropped_hops_on_restructure = (int)v; return BF_OK; }
  if (key == "device") { c->device = (int)v; return BF_OK; }
  // keys a node never reads are ignored, exactly like an unused ROS parameter (SURVEY B-11:
  // phase.launch sets min_mag / smooth_size, which phase.cpp does not read)
  return BF_OK;
}

// beamform_config.yaml subset: scalars and one-line flow maps "micN: {id: .., x: .., y: ..[, z: ..]}".
extern "C" int bf_config_load_yaml(bf_config* c, const char* path) {
  if (!c || !path) return fail(BF_ERR_INVALID, "bf_config_load_yaml: null argument");
  std::ifstream f(path);
  if (!f) return fail(BF_ERR_IO, std::string("cannot open ") + path);
  std::map<int, std::pair<double, double> > mics;
  std::map<int, double> interf;
  std::string line;
  while (std::getline(f, line)) {
    size_t hash = line.find('#');
    if (hash != std::string::npos) line = line.substr(0, hash);
    size_t colon = line.find(':');
    if (colon == std::string::npos) continue;
    std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
    if (key.empty()) continue;
    if (key.compare(0, 3, "mic") == 0 && key.size() > 3 && isdigit((unsigned char)key[3]) && !val.empty() && val[0] == '{') {
      int idx = atoi(key.c_str() + 3);
      double x = 0, y = 0;
      std::string body = val.substr(1, val.find('}') == std::string::npos ? std::string::npos : val.find('}') - 1);
      std::stringstream ss(body);
      std::string item;
      while (std::getline(ss, item, ',')) {
        size_t c2 = item.find(':');
        if (c2 == std::string::npos) continue;
        std::string k = trim(item.substr(0, c2));
        double vv = atof(trim(item.substr(c2 + 1)).c_str());
        if (k == "x") x = vv;
        if (k == "y") y = vv;   // z is ignored by the reference as well (SURVEY B-6)
      }
      mics[idx] = std::make_pair(x, y);
    } else if (key.compare(0, 12, "angle_interf") == 0) {
      interf[atoi(key.c_str() + 12)] = atof(val.c_str());
    } else if (key == "initial_angle") {
      c->initial_angle = atof(val.c_str());
    } else {
      bf_config_set(c, key.c_str(), val.c_str());
    }
  }
  // util.h:82-92: micN are read until the first missing index
  c->n_mics = 0;
  for (int i = 0; i < BF_MAX_MICS; i++) {
    auto it = mics.find(i);
    if (it == mics.end()) break;
    c->mic_x[i] = it->second.first;
    c->mic_y[i] = it->second.second;
    c->n_mics = i + 1;
  }
  // util.h:94-113: angle_interfK are read until the first missing one or the first |a| > 180;
  // the cut is re-applied in bf_create so raw arrays behave the same.
  c->n_angle_interf = 0;
  for (int k = 1; k <= BF_MAX_INTERF; k++) {
    auto it = interf.find(k);
    if (it == interf.end()) break;
    c->angle_interf[k - 1] = it->second;
    c->n_angle_interf = k;
  }
  return BF_OK;
}

// A reference launch file (launch/<node>.launch): the node's operating values live in the inline <rosparam> block
// ("key: value" lines, launch/mvdr.launch:5-11) or in <param name=".." value=".."/> tags; <rosparam command="load" .../>
// lines point at the yaml files, which bf_config_load_yaml reads.
extern "C" int bf_config_load_launch(bf_config* c, const char* path) {
  if (!c || !path) return fail(BF_ERR_INVALID, "bf_config_load_launch: null argument");
  std::ifstream f(path);
  if (!f) return fail(BF_ERR_IO, std::string("cannot open ") + path);
  std::stringstream buf;
  buf << f.rdbuf();
  const std::string txt = buf.str();
  size_t pos = 0;
  int found = 0;
  while ((pos = txt.find("<rosparam", pos)) != std::string::npos) {
    const size_t gt = txt.find('>', pos);
    if (gt == std::string::npos) break;
    const bool self_closing = gt > 0 && txt[gt - 1] == '/';
    pos = gt + 1;
    if (self_closing) continue;   // command="load" file=...
    const size_t end = txt.find("</rosparam>", pos);
    if (end == std::string::npos) return fail(BF_ERR_INVALID, std::string(path) + ": unterminated <rosparam> block");
    std::stringstream block(txt.substr(pos, end - pos));
    std::string line;
    while (std::getline(block, line)) {
      const size_t hash = line.find('#');
      if (hash != std::string::npos) line = line.substr(0, hash);
      const size_t colon = line.find(':');
      if (colon == std::string::npos) continue;
      const std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
      if (key.empty() || val.empty()) continue;
      bf_config_set(c, key.c_str(), val.c_str());
      found++;
    }
    pos = end;
  }
  pos = 0;
  while ((pos = txt.find("<param", pos)) != std::string::npos) {
    const size_t gt = txt.find('>', pos);
    if (gt == std::string::npos) break;
    const std::string tag = txt.substr(pos, gt - pos);
    auto attr = [&](const char* name) -> std::string {
      const size_t a = tag.find(std::string(name) + "=\"");
      if (a == std::string::npos) return std::string();
      const size_t b = a + strlen(name) + 2, e = tag.find('"', b);
      return e == std::string::npos ? std::string() : tag.substr(b, e - b);
    };
    const std::string key = attr("name"), val = attr("value");
    if (!key.empty() && !val.empty()) { bf_config_set(c, key.c_str(), val.c_str()); found++; }
    pos = gt + 1;
  }
  (void)found;   // a launch file without inline parameters is legal (launch/das.launch): the getParam fall-backs stay
  return BF_OK;
}

// ------------------------------------------------------------------------------------------------
// geometry, frequency vector, steering (all double, as the reference)
// ------------------------------------------------------------------------------------------------
static void calculate_delays(const bf_handle* h, double look, double* delay) {   // util.h:136-188
  for (uint32_t i = 0; i < h->M; i++) {
    if (i == 0) { delay[i] = 0.0; continue; }
    double a = h->mic_angle[i] - look;
    if (a > 180) a -= 360;
    else if (a < -180) a += 360;
    delay[i] = h->mic_dist[i] * std::cos(a * kDeg2Rad) / (-kVSound);
  }
}

static void calculate_frequency_vector(bf_handle* h) {   // util.h:190-199
  const uint32_t N = h->N;
  h->freqs.assign(N, 0.0);   // [N/2] is never written by the reference: pinned to 0.0 (SURVEY B-2)
  for (uint32_t i = 0; i < N / 2 - 1; ++i) {
    h->freqs[i + 1] = ((double)(i + 1) / (double)N) * h->cfg.sample_rate;
    h->freqs[N - 1 - i] = -((double)(i + 1) / (double)N) * h->cfg.sample_rate;
  }
  h->freqs[N / 2 - 1] = h->cfg.sample_rate / 2;   // util.h:198 (SURVEY B-1)
}

static inline cd& W(bf_handle* h, uint32_t j, uint32_t i, int k) { return h->weights[((size_t)j * h->M + i) * h->C + k]; }

// lcmv.cpp:221-256: buffers are re-created zero-filled whenever the interferer count changes
static void allocate_interf_buffers(bf_handle* h) {
  h->C = (int)h->interference_angles.size() + 1;
  h->weights.assign((size_t)h->N * h->M * h->C, cd(0, 0));
}

// das.cpp:27-45 / lcmv.cpp:44-86: row 0 is only written when ini (SURVEY B-8)
static void update_weights(bf_handle* h, bool ini) {
  std::vector<double> delay(h->M);
  for (int k = 0; k < h->C; k++) {
    calculate_delays(h, k == 0 ? h->angle : h->interference_angles[k - 1], delay.data());
    for (uint32_t i = 0; i < h->M; i++) {
      if (i == 0) {
        if (ini)
          for (uint32_t j = 0; j < h->N; j++) W(h, j, 0, k) = 1.0;
      } else {
        for (uint32_t j = 0; j < h->N; j++) W(h, j, i, k) = std::exp(-cd(0, 1) * (double)2 * kPi * h->freqs[j] * delay[i]);
      }
    }
  }
  h->tables_dirty = true;
  h->gss_reset_pending = true;   // gss.cpp:90-93
}

static int upload_tables(bf_handle* h, cudaStream_t st) {
  if (!h->tables_dirty) return BF_OK;
  const uint32_t N = h->N, M = h->M, L = h->L;
  const size_t nsteer = (size_t)L * h->C * M;
  if (nsteer > h->steer_cap) {
    if (h->d_steer) cudaFree(h->d_steer);
    CUDA_TRY(cudaMalloc(&h->d_steer, sizeof(float2) * nsteer));
    h->steer_cap = nsteer;
  }
  std::vector<float2> steer(nsteer);
  for (uint32_t l = 0; l < L; l++)
    for (int k = 0; k < h->C; k++)
      for (uint32_t i = 0; i < M; i++) {
        cd w = W(h, l, i, k)   /* logical bin N/2+1 is FFT bin N/2+1 itself */;
        steer[((size_t)l * h->C + k) * M + i] = make_float2((float)w.real(), (float)w.imag());
      }
  CUDA_TRY(cudaMemcpyAsync(h->d_steer, steer.data(), sizeof(float2) * nsteer, cudaMemcpyHostToDevice, st));
  if (h->cfg.algo == BF_ALGO_GSC) {
    // gsc.cpp:62-65: x_fft[j] *= conj(w_ij), no 1/M; Re() of the inverse keeps the Hermitian part (see the DAS table below)
    std::vector<float2> ceff((size_t)M * N);
    for (uint32_t i = 0; i < M; i++)
      for (uint32_t j = 0; j < N; j++) {
        cd a = std::conj(W(h, j, i, 0)), b = W(h, (N - j) % N, i, 0);
        cd c = (a + b) / 2.0;
        ceff[(size_t)i * N + j] = make_float2((float)c.real(), (float)c.imag());
      }
    CUDA_TRY(cudaMemcpyAsync(h->d_das_ceff, ceff.data(), sizeof(float2) * ceff.size(), cudaMemcpyHostToDevice, st));
  }
  if (h->cfg.algo == BF_ALGO_DAS) {
    // Y[j] = (1/M) sum_i conj(w_ij) X_i[j] (das.cpp:60-63), out = Re(IFFT(Y)) (util.h:249).  Re() keeps the
    // Hermitian part Yh[j] = (Y[j] + conj(Y[N-j]))/2 = ceff_i[j] X_i[j] with
    // ceff_i[j] = (conj(w_ij) + w_i,N-j)/(2M): exactly Hermitian, so two frames can share one complex inverse.
    std::vector<float2> ceff((size_t)M * N);
    for (uint32_t i = 0; i < M; i++)
      for (uint32_t j = 0; j < N; j++) {
        cd a = std::conj(W(h, j, i, 0)), b = W(h, (N - j) % N, i, 0);
        cd c = (a + b) / (2.0 * M);
        ceff[(size_t)i * N + j] = make_float2((float)c.real(), (float)c.imag());
      }
    CUDA_TRY(cudaMemcpyAsync(h->d_das_ceff, ceff.data(), sizeof(float2) * ceff.size(), cudaMemcpyHostToDevice, st));
  }
  if (h->d_steer_d) {
    std::vector<double2> sd((size_t)L * M);
    for (uint32_t l = 0; l < L; l++)
      for (uint32_t i = 0; i < M; i++) { cd w = W(h, l, i, 0); sd[(size_t)l * M + i] = make_double2(w.real(), w.imag()); }
    CUDA_TRY(cudaMemcpyAsync(h->d_steer_d, sd.data(), sizeof(double2) * sd.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  std::vector<uint8_t> inband(L);
  for (uint32_t l = 0; l < L; l++) {
    double f = std::fabs(h->freqs[l]);   // mvdr.cpp:78,84
    inband[l] = (f >= h->cfg.freq_min && f <= h->cfg.freq_max) ? 1 : 0;
  }
  CUDA_TRY(cudaMemcpyAsync(h->d_inband, inband.data(), L, cudaMemcpyHostToDevice, st));
  h->inband_host = inband;
  CUDA_TRY(cudaStreamSynchronize(st));   // host vectors go out of scope
  h->tables_dirty = false;
  return BF_OK;
}

// ------------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------------
extern "C" int bf_create(bf_handle** out, const bf_config* cfg, uint32_t n_streams) {
  if (!out || !cfg) return fail(BF_ERR_INVALID, "bf_create: null argument");
  *out = nullptr;
  if (cfg->n_mics < 1 || cfg->n_mics > BF_MAX_MICS) return fail(BF_ERR_INVALID, "bf_create: n_mics must be in [1, 64]");
  if (n_streams < 1) return fail(BF_ERR_INVALID, "bf_create: n_streams must be >= 1");
  if (cfg->hop != 256 && cfg->hop != 512 && cfg->hop != 1024 && cfg->hop != 2048)
    return fail(BF_ERR_INVALID, "bf_create: hop (JACK period) must be 256, 512, 1024 or 2048 (512- to 4096-point frames)");
  {
    const bool sel = cfg->algo == BF_ALGO_MVDR || cfg->algo == BF_ALGO_LCMV || cfg->algo == BF_ALGO_GSS;
    // 1024-point frames with <= 8 microphones run on the register-resident kernels; everything else keeps the
    // spectra of one frame pair in shared memory
    if (sel && (cfg->hop != 512 || cfg->n_mics > 8) && !(cfg->algo == BF_ALGO_GSS && cfg->hop == 512)) {
      if (cfg->n_mics > 16) return fail(BF_ERR_INVALID, "bf_create: mvdr/lcmv (and gss at this frame size) support at most 16 microphones");
      if (bf::frames_kernel_sel_smem(2 * (int)cfg->hop, 1) > 232448)
        return fail(BF_ERR_INVALID, "bf_create: frame size too large for the shared memory");
    }
    if (!sel && cfg->algo != BF_ALGO_DAS && cfg->hop != 512 && bf::frames_kernel_n_smem(2 * (int)cfg->hop, cfg->n_mics, cfg->algo) > 232448)
      return fail(BF_ERR_INVALID, "bf_create: too many microphones for this frame size (spectra must fit 227 KB of shared memory)");
  }
  if (cfg->algo < 0 || cfg->algo > 8) return fail(BF_ERR_INVALID, "bf_create: unknown algo");
  if (cfg->hop == 512 && (cfg->algo == BF_ALGO_PHASE || cfg->algo == BF_ALGO_PHASEMPF || (cfg->algo == BF_ALGO_GSS && cfg->n_mics > 8)) &&
      bf::frames_kernel_smem(cfg->n_mics) > 232448)
    return fail(BF_ERR_INVALID, "bf_create: too many microphones for this node at 1024-point frames (spectra must fit 227 KB of shared memory)");
  if (cfg->algo == BF_ALGO_GSC) {
    if (cfg->filter_size < 32 || cfg->filter_size > 256 || cfg->filter_size % 32) return fail(BF_ERR_INVALID, "bf_create: gsc filter_size must be a multiple of 32 in [32, 256]");
    if (cfg->n_mics > 16) return fail(BF_ERR_INVALID, "bf_create: gsc supports at most 16 microphones");
    if (bf::gsc_align_smem(2 * (int)cfg->hop, cfg->n_mics) > 232448)
      return fail(BF_ERR_INVALID, "bf_create: too many microphones for this frame size (spectra must fit 227 KB of shared memory)");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || cfg->device >= ndev)
    return fail(BF_ERR_NO_DEVICE, "bf_create: no CUDA device (beamform_b200 has no CPU path)");
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail(BF_ERR_NO_DEVICE, "bf_create: kernels are built for sm_100a only");
  CUDA_TRY(cudaSetDevice(cfg->device));

  bf_handle* h = new bf_handle();
  h->sm_count = prop.multiProcessorCount;
  h->cfg = *cfg;
  h->dev = cfg->device;
  h->B = n_streams; h->M = cfg->n_mics; h->H = cfg->hop; h->N = 2 * cfg->hop; h->L = h->N / 2 + 2;
  // util.h:82-92,116-119: polar coordinates from the RAW positions; the later mic0 re-referencing of x,y
  // never feeds back into dist/angle (SURVEY B-6), so only the raw values matter.
  for (uint32_t i = 0; i < h->M; i++) {
    h->mic_dist.push_back(std::sqrt(cfg->mic_x[i] * cfg->mic_x[i] + cfg->mic_y[i] * cfg->mic_y[i]));
    h->mic_angle.push_back(std::atan2(cfg->mic_y[i], cfg->mic_x[i]) * kRad2Deg);
  }
  h->angle = cfg->initial_angle;
  if (cfg->algo == BF_ALGO_LCMV || cfg->algo == BF_ALGO_GSS) {
    for (int k = 0; k < cfg->n_angle_interf && k < BF_MAX_INTERF; k++) {   // util.h:101-112
      if (std::fabs(cfg->angle_interf[k]) > 180) break;
      if (h->interference_angles.size() == kMaxInterferers) {
        delete h;
        return fail(BF_ERR_INVALID, "bf_create: at most 15 interferers (beamform_config.yaml ships angle_interf1..15)");
      }
      h->interference_angles.push_back(cfg->angle_interf[k]);
    }
  }
  calculate_frequency_vector(h);
  allocate_interf_buffers(h);
  update_weights(h, true);

  cudaError_t e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete h; return fail(BF_ERR_CUDA, cudaGetErrorString(e)); }
  cudaStreamCreateWithFlags(&h->st_h2d, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&h->st_d2h, cudaStreamNonBlocking);
  for (int i = 0; i < 8; i++) {
    cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming);
  }
  const size_t prev_n = (size_t)h->B * h->M * h->H, tail_n = (size_t)h->B * h->H;
  bool ok = cudaMalloc(&h->d_prev_hop, sizeof(float) * prev_n) == cudaSuccess &&
            cudaMalloc(&h->d_tail, sizeof(float) * tail_n) == cudaSuccess &&
            (cfg->algo != BF_ALGO_DAS || cudaMalloc(&h->d_tail2, sizeof(float) * tail_n) == cudaSuccess) &&
            cudaMalloc(&h->d_inband, h->L) == cudaSuccess &&
            cudaMalloc(&h->d_das_ceff, sizeof(float2) * h->M * h->N) == cudaSuccess &&
            cudaMalloc(&h->d_stage_in, sizeof(float) * h->M * h->H) == cudaSuccess &&
            cudaMalloc(&h->d_stage_out, sizeof(float) * h->H) == cudaSuccess &&
            cudaMallocHost(&h->h_stage_in, sizeof(float) * h->M * h->H) == cudaSuccess &&
            cudaMallocHost(&h->h_stage_out, sizeof(float) * h->H) == cudaSuccess;
  if (!ok) { bf_destroy(h); return fail(BF_ERR_ALLOC, "bf_create: device allocation failed"); }
  {
    // in-band logical bins (fixed by freq_min/freq_max at create) -> compact state slots
    std::vector<int> slot(h->L, -1), list;
    for (uint32_t l = 0; l < h->L; l++) {
      double f = std::fabs(h->freqs[l]);
      if (f >= cfg->freq_min && f <= cfg->freq_max) { slot[l] = (int)list.size(); list.push_back((int)l); }
    }
    h->Lsel = (int)list.size();
    const bool sel_algo = cfg->algo == BF_ALGO_MVDR || cfg->algo == BF_ALGO_LCMV || cfg->algo == BF_ALGO_GSS;
    const bool pha_algo = cfg->algo == BF_ALGO_PHASE || cfg->algo == BF_ALGO_PHASEMPF || cfg->algo == BF_ALGO_MCRA || cfg->algo == BF_ALGO_REF;
    if (pha_algo) {
      if (cfg->algo == BF_ALGO_PHASEMPF && (cfg->smooth_size < 1 || cfg->smooth_size > 64)) {
        bf_destroy(h);
        return fail(BF_ERR_INVALID, "bf_create: smooth_size must be in [1, 64]");
      }
      const size_t ns = (size_t)h->B * 7 * h->L;
      bool ok3 = cudaMalloc(&h->d_steer_d, sizeof(double2) * h->L * h->M) == cudaSuccess &&
                 cudaMalloc(&h->d_mpf_state, sizeof(float) * ns) == cudaSuccess &&
                 cudaMalloc(&h->d_smooth_hist, sizeof(float) * h->B * 64) == cudaSuccess;
      if (!ok3) { bf_destroy(h); return fail(BF_ERR_ALLOC, "bf_create: device allocation failed (phase state)"); }
      cudaMemset(h->d_mpf_state, 0, sizeof(float) * ns);              // phasempf.cpp:535-545
      cudaMemset(h->d_smooth_hist, 0, sizeof(float) * h->B * 64);     // phasempf.cpp:510 calloc
    }
    if (sel_algo || pha_algo) {
      if (cfg->past_windows < 1) { bf_destroy(h); return fail(BF_ERR_INVALID, "bf_create: past_windows must be >= 1"); }
      std::vector<double> win(h->N);
      std::vector<double2> twd(h->N);
      for (uint32_t n = 0; n < h->N; n++) {
        win[n] = std::sqrt(0.5 - 0.5 * std::cos(2 * kPi * n / (h->N)));   // util.h:201-211
        double ang = -2.0 * kPi * (double)n / (double)h->N;
        twd[n] = make_double2(std::cos(ang), std::sin(ang));
      }
      const size_t nl = std::max(1, h->Lsel);
      bool ok2 = cudaMalloc(&h->d_sel_slot, sizeof(int) * h->L) == cudaSuccess && cudaMalloc(&h->d_sel_list, sizeof(int) * nl) == cudaSuccess &&
                 cudaMalloc(&h->d_win_d, sizeof(double) * h->N) == cudaSuccess && cudaMalloc(&h->d_twid_d, sizeof(double2) * h->N) == cudaSuccess;
      if (ok2 && (cfg->algo == BF_ALGO_MVDR || cfg->algo == BF_ALGO_LCMV)) {
        const size_t nh = (size_t)h->B * nl * (cfg->past_windows + 2) * h->M;
        ok2 = cudaMalloc(&h->d_hist, sizeof(float2) * nh) == cudaSuccess;
        if (ok2) cudaMemset(h->d_hist, 0, sizeof(float2) * nh);   // past_ffts.setZero() (mvdr.cpp:229-233)
      }
      if (ok2 && cfg->algo == BF_ALGO_GSS) ok2 = cudaMalloc(&h->d_gss_w, sizeof(float2) * (size_t)h->B * nl * BF_GSS_ROWS * h->M) == cudaSuccess;
      if (!ok2) { bf_destroy(h); return fail(BF_ERR_ALLOC, "bf_create: device allocation failed (state)"); }
      cudaMemcpy(h->d_sel_slot, slot.data(), sizeof(int) * h->L, cudaMemcpyHostToDevice);
      if (h->Lsel) cudaMemcpy(h->d_sel_list, list.data(), sizeof(int) * h->Lsel, cudaMemcpyHostToDevice);
      cudaMemcpy(h->d_win_d, win.data(), sizeof(double) * h->N, cudaMemcpyHostToDevice);
      cudaMemcpy(h->d_twid_d, twd.data(), sizeof(double2) * h->N, cudaMemcpyHostToDevice);
    }
  }
  {
    // float window / twiddle tables of the frame-size-generic kernel (util.h:201-211)
    std::vector<float> wf(h->N);
    std::vector<float2> tf(h->N);
    for (uint32_t n = 0; n < h->N; n++) {
      wf[n] = (float)std::sqrt(0.5 - 0.5 * std::cos(2 * kPi * n / (h->N)));
      const double ang = -2.0 * kPi * (double)n / (double)h->N;
      tf[n] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    if (cudaMalloc(&h->d_win_f, sizeof(float) * h->N) != cudaSuccess || cudaMalloc(&h->d_twid_f, sizeof(float2) * h->N) != cudaSuccess) {
      bf_destroy(h);
      return fail(BF_ERR_ALLOC, "bf_create: device allocation failed (tables)");
    }
    cudaMemcpy(h->d_win_f, wf.data(), sizeof(float) * h->N, cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_twid_f, tf.data(), sizeof(float2) * h->N, cudaMemcpyHostToDevice);
  }
  if (cfg->algo == BF_ALGO_GSC) {
    const size_t nt = (size_t)h->B * h->M * h->H, nst = (size_t)h->B * (2 * (h->M - 1) + 1) * cfg->filter_size;
    if (cudaMalloc(&h->d_gsc_tail, sizeof(float) * nt) != cudaSuccess || cudaMalloc(&h->d_gsc_state, sizeof(float) * nst) != cudaSuccess ||
        cudaMalloc(&h->d_gsc_head, sizeof(int) * h->B) != cudaSuccess) {
      bf_destroy(h);
      return fail(BF_ERR_ALLOC, "bf_create: device allocation failed (gsc state)");
    }
    cudaMemset(h->d_gsc_tail, 0, sizeof(float) * nt);      // util.h:341-342 calloc
    cudaMemset(h->d_gsc_state, 0, sizeof(float) * nst);    // gsc.cpp:286-289 calloc
    cudaMemset(h->d_gsc_head, 0, sizeof(int) * h->B);
  }
  cudaMemset(h->d_prev_hop, 0, sizeof(float) * prev_n);   // util.h:275-277: one hop of zeros pre-loaded
  cudaMemset(h->d_tail, 0, sizeof(float) * tail_n);       // util.h:285: calloc'ed out_buff
  int rc = upload_tables(h, h->own_stream);
  if (rc != BF_OK) { bf_destroy(h); return rc; }
  *out = h;
  return BF_OK;
}

extern "C" void bf_destroy(bf_handle* h) {
  if (!h) return;
  cudaSetDevice(h->dev);
  if (h->own_stream) { cudaStreamSynchronize(h->own_stream); cudaStreamDestroy(h->own_stream); }
  cudaFree(h->d_prev_hop); cudaFree(h->d_tail); cudaFree(h->d_tail2); cudaFree(h->d_steer); cudaFree(h->d_das_ceff); cudaFree(h->d_inband);
  cudaFree(h->d_stage_in); cudaFree(h->d_stage_out); cudaFree(h->d_io_in); cudaFree(h->d_io_out);
  if (h->st_h2d) cudaStreamDestroy(h->st_h2d);
  if (h->st_d2h) cudaStreamDestroy(h->st_d2h);
  for (int i = 0; i < 8; i++) { if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]); if (h->ev_k[i]) cudaEventDestroy(h->ev_k[i]); }
  cudaFree(h->d_steer_d); cudaFree(h->d_mpf_state); cudaFree(h->d_smooth_hist);
  cudaFree(h->d_sel_ws); cudaFree(h->d_sel_slot); cudaFree(h->d_sel_list); cudaFree(h->d_hist); cudaFree(h->d_gss_w); cudaFree(h->d_win_d); cudaFree(h->d_twid_d); cudaFree(h->d_win_f); cudaFree(h->d_twid_f);
  cudaFree(h->d_srp_xs); cudaFree(h->d_srp_tau); cudaFree(h->d_srp_freqs);
  cudaFree(h->d_gsc_aligned); cudaFree(h->d_gsc_tail); cudaFree(h->d_gsc_state); cudaFree(h->d_gsc_head);
  if (h->h_stage_in) cudaFreeHost(h->h_stage_in);
  if (h->h_stage_out) cudaFreeHost(h->h_stage_out);
  delete h;
}

// ------------------------------------------------------------------------------------------------
// control topics
// ------------------------------------------------------------------------------------------------
static void apply_theta(bf_handle* h, float data) {   // das.cpp:94-99
  h->angle = data;
  update_weights(h, false);
}

// lcmv.cpp:258-309 == gss.cpp:288-339.  Returns 1 when the list was restructured.
static int apply_interference(bf_handle* h, uint16_t id, float msg_angle) {
  std::vector<double>& ia = h->interference_angles;
  int restructured = 0;
  if (id >= 1 && id <= ia.size()) {
    ia[id - 1] = msg_angle;
    for (int i = 0; i < (int)ia.size(); i++) {
      if (i != (id - 1) && std::fabs(ia[i] - msg_angle) < h->cfg.interf_angle_threshold) {
        ia.erase(ia.begin() + id - 1);
        allocate_interf_buffers(h);
        restructured = 1;
        break;
      }
    }
    update_weights(h, false);
  } else if (id > ia.size()) {
    int i;
    for (i = 0; i < (int)ia.size(); i++)
      if (std::fabs(ia[i] - msg_angle) < h->cfg.interf_angle_threshold) break;
    if (i == (int)ia.size() && ia.size() < kMaxInterferers) {   // the list is full: the message is dropped (the reference would grow without bound)
      ia.push_back(msg_angle);
      allocate_interf_buffers(h);
      restructured = 1;
      update_weights(h, false);
    }
  }   // id == 0: "Invalid interference id" (lcmv.cpp:306-308): no change
  return restructured;
}

static void apply_event(bf_handle* h, int kind, uint32_t id, float value) {
  if (kind == 0) {
    apply_theta(h, value);
  } else if (h->cfg.algo == BF_ALGO_LCMV || h->cfg.algo == BF_ALGO_GSS) {   // only lcmv/gss subscribe (lcmv.cpp:320)
    if (apply_interference(h, (uint16_t)id, value)) h->drop_left = h->cfg.dropped_hops_on_restructure;
  }
}

static void apply_event_locked(bf_handle* h, int kind, uint32_t id, float value) {
  std::lock_guard<std::mutex> lk(h->state_mtx);
  apply_event(h, kind, id, value);
}

extern "C" int bf_set_theta(bf_handle* h, float angle_deg) {
  if (!h) return fail(BF_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lk(h->mtx);
  h->pending.push_back(PendingEvent{0, 0, angle_deg});
  return BF_OK;
}
extern "C" int bf_set_interference(bf_handle* h, uint16_t id, float angle_deg) {
  if (!h) return fail(BF_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lk(h->mtx);
  h->pending.push_back(PendingEvent{1, id, angle_deg});
  return BF_OK;
}
static void drain_pending(bf_handle* h) {
  std::vector<PendingEvent> ev;
  {
    std::lock_guard<std::mutex> lk(h->mtx);
    ev.swap(h->pending);
  }
  if (ev.empty()) return;
  std::lock_guard<std::mutex> lk(h->state_mtx);
  for (const PendingEvent& e : ev) apply_event(h, e.kind, e.id, e.value);
}
extern "C" int bf_get_theta(bf_handle* h, double* a) {
  if (!h || !a) return fail(BF_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(h->state_mtx);   // read-only: queued updates are applied by the processing thread at a hop boundary
  *a = h->angle;
  return BF_OK;
}
extern "C" int bf_get_interferences(bf_handle* h, double* angles, uint32_t cap, uint32_t* n) {
  if (!h || !n) return fail(BF_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(h->state_mtx);
  *n = (uint32_t)h->interference_angles.size();
  for (uint32_t i = 0; i < *n && i < cap && angles; i++) angles[i] = h->interference_angles[i];
  return BF_OK;
}

// ------------------------------------------------------------------------------------------------
// processing
// ------------------------------------------------------------------------------------------------
__global__ void zero_hops_kernel(float* out, long long stride, int n) {
  float* o = out + (size_t)blockIdx.x * stride;
  for (int i = threadIdx.x; i < n; i += blockDim.x) o[i] = 0.0f;
}

static int run_segment(bf_handle* h, const float* in, size_t ss, size_t ms, float* out, size_t os, uint32_t h0, uint32_t h1,
                       uint32_t call_hops, cudaStream_t st, uint32_t s0 = 0, uint32_t ns = 0) {
  if (h1 <= h0) return BF_OK;
  int rc = upload_tables(h, st);
  if (rc != BF_OK) return rc;
  bf::KernelParams p;
  memset(&p, 0, sizeof(p));
  p.in = in + (size_t)h0 * h->H;
  p.in_stream_stride = (long long)ss; p.in_mic_stride = (long long)ms;
  p.out = out + (size_t)h0 * h->H;
  p.out_stream_stride = (long long)os;
  p.n_streams = ns ? ns : h->B; p.stream_begin = (int)s0; p.M = h->M; p.H = h->H; p.N = h->N;
  p.hop_begin = 0; p.hop_end = (int)(h1 - h0);
  p.frame_index0 = (int)(h->frames_done & 0x7fffffff);
  p.prev_hop = h->d_prev_hop; p.tail = h->d_tail; p.tail_out = h->d_tail;
  p.steer = h->d_steer; p.das_ceff = h->d_das_ceff; p.inband = h->d_inband; p.C = h->C;
  p.das_chunk = bf::frames_kernel_n_das_chunk((int)h->N, (int)h->M);
  p.sel_chunk = (int)h->M; p.sel_ws = nullptr;
  // (mcra applies out_amp to the magnitudes itself, like phasempf)
  const bool amp = h->cfg.algo == BF_ALGO_MVDR || h->cfg.algo == BF_ALGO_LCMV || h->cfg.algo == BF_ALGO_GSS;
  p.out_scale = (float)((amp ? h->cfg.out_amp : 1.0) / (double)h->N);
  if (h->d_capture) {
    p.capture = h->d_capture + (size_t)h0 * h->N;
    p.capture_stream_stride = (long long)call_hops * h->N;
  }
  p.thr_mag = (float)(h->cfg.freq_mag_threshold * (double)h->M * (double)h->N);
  p.thr_mag_d = h->cfg.freq_mag_threshold;
  p.P = (int)h->cfg.past_windows;
  { const char* dbg = getenv("BF_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }   // read per launch: tests switch it
  p.ring_depth = (int)h->cfg.past_windows + 2;
  p.ring_slot0 = (int)(h->frames_done % (uint64_t)p.ring_depth);
  p.hist = h->d_hist; p.sel_slot = h->d_sel_slot; p.sel_list = h->d_sel_list; p.Lsel = h->Lsel;
  p.mu = (float)h->cfg.mu;
  p.lambda_mu = (float)(1 - h->cfg.lambda * h->cfg.mu);
  p.gss_w = h->d_gss_w;
  p.gss_dj2_scale = (float)(2 * (1 / (size_t)h->C));   // gss.cpp:133: integer arithmetic (SURVEY B-7)
  p.win_f = h->d_win_f; p.twid_f = h->d_twid_f;
  p.win_d = h->d_win_d; p.twid_d = h->d_twid_d; p.steer_d = h->d_steer_d;
  p.mag_threshold_d = h->cfg.mag_threshold;
  p.min_phase_rad_d = h->cfg.min_phase * M_PI / 180;   // phase.cpp:175
  p.min_phase_rad = (float)p.min_phase_rad_d;
  p.thr_phase_mag = (float)(h->cfg.mag_threshold * (double)h->M * (double)h->N);
  p.mag_mult = (float)h->cfg.mag_mult;
  p.min_mag = (float)h->cfg.min_mag;
  p.mpf_state = h->d_mpf_state; p.smooth_hist = h->d_smooth_hist; p.smooth_size = h->raw_frame_mode ? 1 : h->cfg.smooth_size;
  p.mcra_alphaS = (float)h->cfg.MCRA_alphaS; p.mcra_alphaD = (float)h->cfg.MCRA_alphaD;
  p.mcra_alphaD2 = (float)h->cfg.MCRA_alphaD2; p.mcra_delta = (float)h->cfg.MCRA_delta;
  p.mcra_L = h->cfg.MCRA_L; p.mcra_cur_L0 = h->mcra_cur_L; p.mcra_first0 = h->mcra_first;
  p.mpf_alphaS = (float)h->cfg.MPF_alphaS; p.mpf_eta = (float)h->cfg.MPF_eta; p.mpf_gamma = (float)h->cfg.MPF_rev_gamma;
  p.mpf_rev_gain = (float)(1 - h->cfg.MPF_rev_gamma / h->cfg.MPF_rev_delta);   // phasempf.cpp:265-266
  p.out_amp = (float)h->cfg.out_amp; p.noise_floor = (float)h->cfg.noise_floor;
  p.out_only_noise = h->cfg.out_only_noise; p.out_only_mcra = h->cfg.out_only_mcra;
  if (h->cfg.algo == BF_ALGO_GSS && h->gss_reset_pending) {
    bf::KernelParams pr = p;
    pr.n_streams = h->B; pr.stream_begin = 0;
    CUDA_TRY(bf::launch_gss_reset(pr, st));
    h->launches++;
    h->gss_reset_pending = false;
  }
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (h->profiling) {
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    CUDA_TRY(cudaEventRecord(ev0, st));
  }
  bool swap_tails = false;
  static const bool force_generic = getenv("BF_GENERIC") != nullptr;   // debug: cross-check the generic kernel at N = 1024
  const bool gen_algo = h->cfg.algo == BF_ALGO_DAS || h->cfg.algo == BF_ALGO_PHASE || h->cfg.algo == BF_ALGO_PHASEMPF;
  const bool phase_f64 = getenv("BF_PHASE_F64") != nullptr;   // tests: the double-spectra / 1024-point CTA kernels instead of phase_n_kernel (read per launch)
  static const bool force_sel_generic = getenv("BF_SEL_GENERIC") != nullptr;   // debug: cross-check the general gated kernel
  const bool sel_algo = h->cfg.algo == BF_ALGO_MVDR || h->cfg.algo == BF_ALGO_LCMV || h->cfg.algo == BF_ALGO_GSS;
  if (h->cfg.algo == BF_ALGO_GSC) {
    const size_t need = (size_t)p.n_streams * h->M * (size_t)(h1 - h0) * h->H;
    if (need > h->gsc_aligned_cap) {
      CUDA_TRY(cudaStreamSynchronize(st));
      if (h->d_gsc_aligned) cudaFree(h->d_gsc_aligned);
      h->d_gsc_aligned = nullptr; h->gsc_aligned_cap = 0;
      if (cudaMalloc(&h->d_gsc_aligned, sizeof(float) * need) != cudaSuccess) return fail(BF_ERR_ALLOC, "gsc: aligned-signal workspace");
      h->gsc_aligned_cap = need;
    }
    p.gsc_aligned = h->d_gsc_aligned;
    p.gsc_aligned_stream_stride = (long long)h->M * (long long)(h1 - h0) * h->H;
    p.gsc_tail = h->d_gsc_tail; p.gsc_state = h->d_gsc_state; p.gsc_head = h->d_gsc_head;
    p.gsc_F = h->cfg.filter_size; p.gsc_use_vad = h->cfg.use_vad;
    p.gsc_vad_threshold = h->cfg.vad_threshold; p.gsc_mu0 = h->cfg.mu0; p.gsc_mu_max = h->cfg.mu_max;
    CUDA_TRY(bf::launch_gsc(p, st));
    h->launches++;
  } else if (h->cfg.algo == BF_ALGO_MCRA) {
    if (bf::mcra_pairs_supported(p)) CUDA_TRY(bf::launch_mcra_pairs(p, st, h->sm_count));   // 1024-point frames: a warp per stream
    else CUDA_TRY(bf::launch_frames_kernel_mcra(p, st));
  }
  else if (h->cfg.algo == BF_ALGO_REF) CUDA_TRY(bf::launch_ref_kernel(p, st));
  else if (sel_algo && (h->N != 1024 || force_sel_generic || (h->M > 8 && h->cfg.algo != BF_ALGO_GSS) || h->C > 8)) {   // > 7 interferers: general kernel
    p.sel_chunk = bf::frames_kernel_sel_chunk((int)h->N, (int)h->M);
    if (p.sel_chunk < (int)h->M) {
      if (!h->d_sel_ws && cudaMalloc(&h->d_sel_ws, sizeof(float2) * (size_t)h->B * h->M * h->N) != cudaSuccess)
        return fail(BF_ERR_ALLOC, "spectra workspace of the general gated kernel");
      p.sel_ws = h->d_sel_ws;
    }
    CUDA_TRY(bf::launch_frames_kernel_sel(h->cfg.algo, p, st));
  }
  else if (!phase_f64 && bf::phase_n_supported(p, h->cfg.algo)) CUDA_TRY(bf::launch_phase_n(h->cfg.algo, p, st));   // FP32 spectra, 3 CTAs per SM
  else if (h->N != 1024 || (force_generic && gen_algo)) CUDA_TRY(bf::launch_frames_kernel_n(h->cfg.algo, p, st));
  else if (h->cfg.algo == BF_ALGO_DAS && bf::das_pairs_supported(p)) {
    // a stream's first and last pair may belong to different warps: the new tails go to the second buffer (no
    // read-after-write inside the launch) and the buffers change roles once every stream chunk of the segment has run
    p.tail_out = h->d_tail2;
    CUDA_TRY(bf::launch_das_pairs(p, st, h->sm_count));
    swap_tails = true;
  }
  else if (bf::sel_stream_supported(p, h->cfg.algo, h->inband_host.data())) CUDA_TRY(bf::launch_sel_stream(h->cfg.algo, p, st));
  else if (bf::sel_pairs_supported(p, h->cfg.algo)) CUDA_TRY(bf::launch_sel_pairs(h->cfg.algo, p, st));
  else CUDA_TRY(bf::launch_frames_kernel_1024(h->cfg.algo, p, st));
  if (h->profiling) {
    CUDA_TRY(cudaEventRecord(ev1, st));
    h->prof_events.push_back(std::make_pair(ev0, ev1));
  }
  CUDA_TRY(bf::launch_save_prev_hop(p, (int)(h1 - h0) - 1, st));
  h->launches += 2;
  if (s0 + p.n_streams >= h->B) {   // the last stream chunk closes the segment
    if (swap_tails) std::swap(h->d_tail, h->d_tail2);
    h->frames_done += h1 - h0;
    if (h->cfg.algo == BF_ALGO_PHASEMPF || h->cfg.algo == BF_ALGO_MCRA)
      for (uint32_t t = h0; t < h1; t++) {   // phasempf.cpp:162-176, mcra.cpp:100-113: window counters advance once per frame
        if (h->mcra_cur_L > h->cfg.MCRA_L) { h->mcra_cur_L = 1; h->mcra_first = 0; } else h->mcra_cur_L++;
      }
  }
  return BF_OK;
}

extern "C" int bf_process_batch_device(bf_handle* h, const float* in, size_t ss, size_t ms, float* out, size_t os, uint32_t n_hops,
                                       const bf_event* ev, uint32_t n_ev, void* cuda_stream) {
  if (!h) return fail(BF_ERR_INVALID, "bf_process_batch_device: null handle");
  if (n_hops == 0) return BF_OK;   // an empty batch is a no-op (numpy hands out null data pointers for empty arrays)
  if (!in || !out) return fail(BF_ERR_INVALID, "bf_process_batch_device: null argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  drain_pending(h);
  uint32_t e = 0, t = 0;
  while (t < n_hops) {
    while (e < n_ev && ev[e].hop_index <= t) { apply_event_locked(h, ev[e].kind, ev[e].id, ev[e].value); e++; }
    if (h->drop_left > 0) {
      // READY=false: the callback emits zeros and does not feed the ring buffers (lcmv.cpp:148-155)
      uint32_t nd = std::min<uint32_t>((uint32_t)h->drop_left, n_hops - t);
      zero_hops_kernel<<<h->B, 256, 0, st>>>(out + (size_t)t * h->H, (long long)os, (int)(nd * h->H));
      h->launches++;
      h->drop_left -= (int)nd;
      t += nd;
      continue;
    }
    uint32_t t1 = (e < n_ev) ? std::min<uint32_t>(ev[e].hop_index, n_hops) : n_hops;
    if (t1 <= t) t1 = t + 1;
    int rc = run_segment(h, in, ss, ms, out, os, t, t1, n_hops, st);
    if (rc != BF_OK) return rc;
    t = t1;
  }
  while (e < n_ev) { apply_event_locked(h, ev[e].kind, ev[e].id, ev[e].value); e++; }   // events at/after the end
  return BF_OK;
}

static int ensure_io_staging(bf_handle* h, size_t in_floats, size_t out_floats) {
  if (in_floats > h->io_in_cap) {
    if (h->d_io_in) cudaFree(h->d_io_in);
    h->d_io_in = nullptr; h->io_in_cap = 0;
    if (cudaMalloc(&h->d_io_in, sizeof(float) * in_floats) != cudaSuccess) return fail(BF_ERR_ALLOC, "bf_process_batch: input staging alloc");
    h->io_in_cap = in_floats;
  }
  if (out_floats > h->io_out_cap) {
    if (h->d_io_out) cudaFree(h->d_io_out);
    h->d_io_out = nullptr; h->io_out_cap = 0;
    if (cudaMalloc(&h->d_io_out, sizeof(float) * out_floats) != cudaSuccess) return fail(BF_ERR_ALLOC, "bf_process_batch: output staging alloc");
    h->io_out_cap = out_floats;
  }
  return BF_OK;
}

// Host-buffer entry (what a rosjack-side caller or the offline driver uses): H2D, kernels, D2H.
// Without scheduled events the batch is cut into stream chunks that flow through three CUDA streams
// (copy-in / compute / copy-out) so PCIe transfers overlap the kernels.
extern "C" int bf_process_batch(bf_handle* h, const float* in_host, size_t ss, size_t ms, float* out_host, size_t os,
                                uint32_t n_hops, const bf_event* ev, uint32_t n_ev) {
  if (!h) return fail(BF_ERR_INVALID, "bf_process_batch: null handle");
  if (n_hops == 0) return BF_OK;   // an empty batch is a no-op
  if (!in_host || !out_host) return fail(BF_ERR_INVALID, "bf_process_batch: null argument");
  CUDA_TRY(cudaSetDevice(h->dev));
  const size_t L = (size_t)n_hops * h->H;
  int rc = ensure_io_staging(h, (size_t)h->B * h->M * L, (size_t)h->B * L);
  if (rc != BF_OK) return rc;
  float *d_in = h->d_io_in, *d_out = h->d_io_out;
  drain_pending(h);
  const bool dense = (ms == L && ss == (size_t)h->M * L);
  if (n_ev == 0 && h->drop_left == 0 && h->B >= 16) {
    rc = upload_tables(h, h->own_stream);
    if (rc != BF_OK) return rc;
    // gsc: the NLMS is one warp per stream and sequential in time, so a launch costs the same for 100 or 4 000 streams:
    // splitting the batch would multiply its time, not overlap it
    const uint32_t nchunk = (h->cfg.algo == BF_ALGO_GSC) ? 1 : 8, per = (h->B + nchunk - 1) / nchunk;
    for (uint32_t c = 0, s0 = 0; s0 < h->B; c++, s0 += per) {
      const uint32_t ns = std::min(per, h->B - s0);
      if (dense) {
        CUDA_TRY(cudaMemcpyAsync(d_in + (size_t)s0 * h->M * L, in_host + (size_t)s0 * ss, sizeof(float) * ns * h->M * L, cudaMemcpyHostToDevice, h->st_h2d));
      } else {
        for (uint32_t s = s0; s < s0 + ns; s++)
          CUDA_TRY(cudaMemcpy2DAsync(d_in + (size_t)s * h->M * L, sizeof(float) * L, in_host + s * ss, sizeof(float) * ms, sizeof(float) * L, h->M,
                                     cudaMemcpyHostToDevice, h->st_h2d));
      }
      CUDA_TRY(cudaEventRecord(h->ev_in[c], h->st_h2d));
      CUDA_TRY(cudaStreamWaitEvent(h->own_stream, h->ev_in[c], 0));
      rc = run_segment(h, d_in, (size_t)h->M * L, L, d_out, L, 0, n_hops, n_hops, h->own_stream, s0, ns);
      if (rc != BF_OK) return rc;
      CUDA_TRY(cudaEventRecord(h->ev_k[c], h->own_stream));
      CUDA_TRY(cudaStreamWaitEvent(h->st_d2h, h->ev_k[c], 0));
      CUDA_TRY(cudaMemcpy2DAsync(out_host + (size_t)s0 * os, sizeof(float) * os, d_out + (size_t)s0 * L, sizeof(float) * L, sizeof(float) * L, ns,
                                 cudaMemcpyDeviceToHost, h->st_d2h));
    }
    CUDA_TRY(cudaStreamSynchronize(h->st_d2h));
    CUDA_TRY(cudaStreamSynchronize(h->own_stream));
    return BF_OK;
  }
  cudaStream_t st = h->own_stream;
  for (uint32_t s = 0; s < h->B; s++)
    CUDA_TRY(cudaMemcpy2DAsync(d_in + (size_t)s * h->M * L, sizeof(float) * L, in_host + s * ss, sizeof(float) * ms, sizeof(float) * L, h->M,
                               cudaMemcpyHostToDevice, st));
  rc = bf_process_batch_device(h, d_in, (size_t)h->M * L, L, d_out, L, n_hops, ev, n_ev, st);
  if (rc != BF_OK) return rc;
  CUDA_TRY(cudaMemcpy2DAsync(out_host, sizeof(float) * os, d_out, sizeof(float) * L, sizeof(float) * L, h->B, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return BF_OK;
}

extern "C" int bf_process_hop(bf_handle* h, const float* const* in, float* out, uint32_t nframes) {
  if (!h || !in || !out) return fail(BF_ERR_INVALID, "bf_process_hop: null argument");
  if (h->B != 1) return fail(BF_ERR_INVALID, "bf_process_hop: handle must have n_streams == 1");
  if (nframes != h->H) return fail(BF_ERR_INVALID, "bf_process_hop: nframes != hop (JACK period changed?)");
  CUDA_TRY(cudaSetDevice(h->dev));
  cudaStream_t st = h->own_stream;
  for (uint32_t m = 0; m < h->M; m++) memcpy(h->h_stage_in + (size_t)m * h->H, in[m], sizeof(float) * h->H);
  CUDA_TRY(cudaMemcpyAsync(h->d_stage_in, h->h_stage_in, sizeof(float) * h->M * h->H, cudaMemcpyHostToDevice, st));
  int rc = bf_process_batch_device(h, h->d_stage_in, (size_t)h->M * h->H, h->H, h->d_stage_out, h->H, 1, nullptr, 0, st);
  if (rc != BF_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(h->h_stage_out, h->d_stage_out, sizeof(float) * h->H, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  memcpy(out, h->h_stage_out, sizeof(float) * h->H);
  return BF_OK;
}

// The per-frame operator seam: void (*weight_func)(jack_ringbuffer_t **in, rosjack_data *out) of util.h:289 — one frame
// of fft_win samples per microphone in, fft_win windowed output samples out (util.h:244-253), before the overlap-add.
// The fused kernels overlap-add internally, so the frame [a | b] is run as "previous hop a, hop b" from an empty
// overlap-add tail: the first half of the synthesised frame is the hop's output, the second half is the new tail.
extern "C" int bf_apply_weights(bf_handle* h, const float* const* in_frames, float* out_frame, uint32_t fft_win) {
  if (!h || !in_frames || !out_frame) return fail(BF_ERR_INVALID, "bf_apply_weights: null argument");
  if (h->B != 1) return fail(BF_ERR_INVALID, "bf_apply_weights: handle must have n_streams == 1");
  if (fft_win != h->N) return fail(BF_ERR_INVALID, "bf_apply_weights: frame length != fft_win (2 * JACK period)");
  if (h->cfg.algo == BF_ALGO_GSC || h->cfg.algo == BF_ALGO_REF)
    return fail(BF_ERR_INVALID, "bf_apply_weights: gsc and rosjack_ref have no per-frame operator of this shape (do_overlap_bymic / window only)");
  CUDA_TRY(cudaSetDevice(h->dev));
  cudaStream_t st = h->own_stream;
  const uint32_t H = h->H;
  for (uint32_t m = 0; m < h->M; m++) {
    CUDA_TRY(cudaMemcpyAsync(h->d_prev_hop + (size_t)m * H, in_frames[m], sizeof(float) * H, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->d_stage_in + (size_t)m * H, in_frames[m] + H, sizeof(float) * H, cudaMemcpyHostToDevice, st));
  }
  CUDA_TRY(cudaMemsetAsync(h->d_tail, 0, sizeof(float) * H, st));
  h->raw_frame_mode = true;
  int rc = bf_process_batch_device(h, h->d_stage_in, (size_t)h->M * H, H, h->d_stage_out, H, 1, nullptr, 0, st);
  h->raw_frame_mode = false;
  if (rc != BF_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(out_frame, h->d_stage_out, sizeof(float) * H, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out_frame + H, h->d_tail, sizeof(float) * H, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return BF_OK;
}

extern "C" int bf_set_capture(bf_handle* h, uint8_t* dev_flags) {
  if (!h) return fail(BF_ERR_INVALID, "null handle");
  h->d_capture = dev_flags;
  return BF_OK;
}

// Steered-response sweep (config C5).  Stateless with respect to the beamformer: frame 0 uses the handle's
// "previous hop" state (zeros before the first process call) and nothing is advanced.
extern "C" int bf_srp_batch_device(bf_handle* h, const float* in_dev, size_t ss, size_t ms, const float* thetas_deg_host, uint32_t n_dirs,
                                   float* maps_dev, uint32_t n_hops, void* cuda_stream) {
  if (!h || !in_dev || !thetas_deg_host || !maps_dev) return fail(BF_ERR_INVALID, "bf_srp_batch_device: null argument");
  if (h->N != 1024) return fail(BF_ERR_INVALID, "bf_srp_batch_device: built for 1024-point frames (hop 512)");
  if (n_dirs < 1 || n_hops < 1) return fail(BF_ERR_INVALID, "bf_srp_batch_device: empty sweep");
  CUDA_TRY(cudaSetDevice(h->dev));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const size_t F = (size_t)h->B * n_hops;
  const size_t need_xs = bf::srp_workspace_bytes((long long)F);
  if (need_xs > h->srp_xs_cap) {
    if (h->d_srp_xs) cudaFree(h->d_srp_xs);
    h->d_srp_xs = nullptr; h->srp_xs_cap = 0;
    if (cudaMalloc(&h->d_srp_xs, need_xs) != cudaSuccess) return fail(BF_ERR_ALLOC, "bf_srp_batch_device: spectra workspace");
    h->srp_xs_cap = need_xs;
    // rows of frames beyond F, microphones beyond M and the layout padding are never written: they must read as zero
    CUDA_TRY(cudaMemsetAsync(h->d_srp_xs, 0, need_xs, st));
  }
  const size_t need_tau = (size_t)n_dirs * h->M;
  if (need_tau > h->srp_tau_cap) {
    if (h->d_srp_tau) cudaFree(h->d_srp_tau);
    h->d_srp_tau = nullptr; h->srp_tau_cap = 0;
    if (cudaMalloc(&h->d_srp_tau, sizeof(double) * need_tau) != cudaSuccess) return fail(BF_ERR_ALLOC, "bf_srp_batch_device: delay table");
    h->srp_tau_cap = need_tau;
  }
  if (!h->d_srp_freqs && cudaMalloc(&h->d_srp_freqs, sizeof(double) * h->L) != cudaSuccess) return fail(BF_ERR_ALLOC, "bf_srp_batch_device: freq table");
  std::vector<double> tau(need_tau), fl(h->L);
  for (uint32_t d = 0; d < n_dirs; d++) calculate_delays(h, (double)thetas_deg_host[d], tau.data() + (size_t)d * h->M);   // util.h:136-161
  for (uint32_t l = 0; l < h->L; l++) fl[l] = h->freqs[l];   // logical bin N/2+1 is FFT bin N/2+1 itself
  CUDA_TRY(cudaMemcpyAsync(h->d_srp_tau, tau.data(), sizeof(double) * need_tau, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(h->d_srp_freqs, fl.data(), sizeof(double) * h->L, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));   // host tables go out of scope
  bf::KernelParams p;
  memset(&p, 0, sizeof(p));
  p.in = in_dev; p.in_stream_stride = (long long)ss; p.in_mic_stride = (long long)ms;
  p.n_streams = h->B; p.M = h->M; p.H = h->H; p.N = h->N;
  p.prev_hop = h->d_prev_hop;
  CUDA_TRY(bf::launch_srp(p, h->d_srp_xs, h->d_srp_tau, h->d_srp_freqs, maps_dev, (int)n_dirs, (int)n_hops, st));
  h->launches += 2;
  return BF_OK;
}

extern "C" int bf_set_profiling(bf_handle* h, int enabled) {
  if (!h) return fail(BF_ERR_INVALID, "null handle");
  h->profiling = enabled != 0;
  return BF_OK;
}
extern "C" int bf_get_profile(bf_handle* h, double* kernel_ms, uint64_t* n) {
  if (!h || !kernel_ms || !n) return fail(BF_ERR_INVALID, "null argument");
  double tot = 0;
  for (auto& e : h->prof_events) {
    CUDA_TRY(cudaEventSynchronize(e.second));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e.first, e.second));
    tot += ms;
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  *kernel_ms = tot;
  *n = h->prof_events.size();
  h->prof_events.clear();
  return BF_OK;
}

extern "C" uint32_t bf_fft_win(const bf_handle* h) { return h ? h->N : 0; }
extern "C" uint64_t bf_kernel_launches(const bf_handle* h) { return h ? h->launches : 0; }
extern "C" const char* bf_last_error(void) { return g_err.c_str(); }
extern "C" const char* bf_version(void) { return "beamform_b200 0.1 (sm_100a)"; }
