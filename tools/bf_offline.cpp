// bf_offline — offline file driver of beamform_b200: the stand-in for the JACK/ROS transport (rosjack.cpp) when beamforming
// recorded or synthetic audio.  Host side in C++, everything through the C ABI of include/beamform_b200.h.
//
//   bf_offline --algo mvdr --config beamform_config.yaml [--launch mvdr.launch] --in mics.wav --out beam.wav
//              [--hop 512] [--theta DEG] [--set key=value ...] [--theta-at HOP:DEG ...] [--interf-at HOP:ID:DEG ...]
//              [--events FILE] [--raw M:SR] [--pcm16]
//
//   --config   the reference's beamform_config.yaml, unchanged (geometry, initial_angle, angle_interfK; util.h:52-134)
//   --launch   a reference launch file: the keys of its inline <rosparam> block are applied (launch/mvdr.launch:5-11)
//   --set      one rosparam key (overrides the launch file), e.g. --set freq_mag_threshold=0.001
//   --in       WAV, one channel per microphone (PCM 16/24/32 or IEEE float 32), or with --raw M:SR a raw float32 file laid
//              out [M][L] (planar); --hop is the JACK period (frame = 2 x hop); a ragged tail is dropped (JACK delivers whole periods)
//   --out      mono WAV at the same rate: float32, or 16-bit PCM with --pcm16 (rosjack's own WAV sink is 16-bit, rosjack.cpp:196-198);
//              a name ending in .f32 writes raw float32
//   --theta-at / --interf-at / --events   the two control topics, scheduled: applied before hop HOP.  Event file lines:
//              "<hop> theta <deg>" | "<hop> interf <id> <deg>"
// Exit status 0 on success; 1 with a message on stderr otherwise (no CPU fallback: without a B200 bf_create fails).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../include/beamform_b200.h"

static int die(const std::string& msg) {
  fprintf(stderr, "bf_offline: %s\n", msg.c_str());
  return 1;
}

static int algo_of(const std::string& a) {
  static const char* names[] = {"das", "mvdr", "lcmv", "gss", "phase", "phasempf", "mcra", "ref", "gsc"};
  for (int i = 0; i < 9; i++)
    if (a == names[i]) return i;
  return -1;
}

// ---- minimal RIFF/WAVE reader: fmt (PCM 1, IEEE float 3, extensible 0xFFFE) + data; interleaved -> planar float32 ----
static bool read_wav(const std::string& path, std::vector<float>& planar, uint32_t& channels, uint32_t& rate, size_t& frames, std::string& err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { err = "cannot open " + path; return false; }
  std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (buf.size() < 12 || memcmp(buf.data(), "RIFF", 4) || memcmp(buf.data() + 8, "WAVE", 4)) { err = path + ": not a RIFF/WAVE file"; return false; }
  uint16_t fmt = 0, bits = 0, ch = 0;
  uint32_t sr = 0;
  const uint8_t* data = nullptr;
  size_t data_len = 0;
  for (size_t pos = 12; pos + 8 <= buf.size();) {
    uint32_t len;
    memcpy(&len, buf.data() + pos + 4, 4);
    const uint8_t* body = buf.data() + pos + 8;
    if (pos + 8 + len > buf.size()) len = (uint32_t)(buf.size() - pos - 8);   // truncated last chunk: take what is there
    if (!memcmp(buf.data() + pos, "fmt ", 4) && len >= 16) {
      memcpy(&fmt, body, 2); memcpy(&ch, body + 2, 2); memcpy(&sr, body + 4, 4); memcpy(&bits, body + 14, 2);
      if (fmt == 0xFFFE && len >= 26) memcpy(&fmt, body + 24, 2);   // WAVE_FORMAT_EXTENSIBLE: sub-format GUID starts with the tag
    } else if (!memcmp(buf.data() + pos, "data", 4)) {
      data = body; data_len = len;
    }
    pos += 8 + (size_t)len + (len & 1);
  }
  if (!data || !ch || !sr) { err = path + ": missing fmt/data chunk"; return false; }
  const size_t bps = bits / 8;
  if (!((fmt == 1 && (bits == 16 || bits == 24 || bits == 32)) || (fmt == 3 && bits == 32))) { err = path + ": unsupported sample format"; return false; }
  frames = data_len / (bps * ch);
  channels = ch; rate = sr;
  planar.resize((size_t)ch * frames);
  for (size_t n = 0; n < frames; n++)
    for (uint32_t c = 0; c < ch; c++) {
      const uint8_t* s = data + (n * ch + c) * bps;
      float v;
      if (fmt == 3) { memcpy(&v, s, 4); }
      else if (bits == 16) { int16_t q; memcpy(&q, s, 2); v = (float)q / 32768.0f; }
      else if (bits == 24) { int32_t q = (int32_t)((uint32_t)s[0] << 8 | (uint32_t)s[1] << 16 | (uint32_t)s[2] << 24); v = (float)(q >> 8) / 8388608.0f; }
      else { int32_t q; memcpy(&q, s, 4); v = (float)((double)q / 2147483648.0); }
      planar[(size_t)c * frames + n] = v;
    }
  return true;
}

static bool write_wav(const std::string& path, const float* x, size_t n, uint32_t rate, bool pcm16, std::string& err) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { err = "cannot write " + path; return false; }
  const uint16_t fmt = pcm16 ? 1 : 3, ch = 1, bits = pcm16 ? 16 : 32, align = bits / 8;
  const uint32_t data_len = (uint32_t)(n * align), riff_len = 36 + data_len, fmt_len = 16, byte_rate = rate * align;
  fwrite("RIFF", 1, 4, f); fwrite(&riff_len, 4, 1, f); fwrite("WAVEfmt ", 1, 8, f); fwrite(&fmt_len, 4, 1, f);
  fwrite(&fmt, 2, 1, f); fwrite(&ch, 2, 1, f); fwrite(&rate, 4, 1, f); fwrite(&byte_rate, 4, 1, f); fwrite(&align, 2, 1, f); fwrite(&bits, 2, 1, f);
  fwrite("data", 1, 4, f); fwrite(&data_len, 4, 1, f);
  if (pcm16) {
    std::vector<int16_t> q(n);
    for (size_t i = 0; i < n; i++) q[i] = (int16_t)std::lrint(std::max(-1.0f, std::min(1.0f, x[i])) * 32767.0f);
    fwrite(q.data(), 2, n, f);
  } else {
    fwrite(x, 4, n, f);
  }
  fclose(f);
  return true;
}

int main(int argc, char** argv) {
  std::string algo, config, launch, in_path, out_path, events_path;
  std::vector<std::pair<std::string, std::string> > sets;
  std::vector<bf_event> events;
  uint32_t hop = 512, raw_m = 0, raw_sr = 0;
  bool have_theta = false, pcm16 = false;
  double theta = 0;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto need = [&](const char* what) -> const char* {
      if (i + 1 >= argc) { die(std::string(what) + " needs a value"); exit(1); }
      return argv[++i];
    };
    if (a == "--algo") algo = need("--algo");
    else if (a == "--config") config = need("--config");
    else if (a == "--launch") launch = need("--launch");
    else if (a == "--in") in_path = need("--in");
    else if (a == "--out") out_path = need("--out");
    else if (a == "--events") events_path = need("--events");
    else if (a == "--hop") hop = (uint32_t)atoi(need("--hop"));
    else if (a == "--theta") { theta = atof(need("--theta")); have_theta = true; }
    else if (a == "--pcm16") pcm16 = true;
    else if (a == "--raw") { if (sscanf(need("--raw"), "%u:%u", &raw_m, &raw_sr) != 2) return die("--raw wants M:SR"); }
    else if (a == "--set") {
      const std::string kv = need("--set");
      const size_t eq = kv.find('=');
      if (eq == std::string::npos) return die("--set wants key=value");
      sets.push_back(std::make_pair(kv.substr(0, eq), kv.substr(eq + 1)));
    } else if (a == "--theta-at") {
      unsigned h; float d;
      if (sscanf(need("--theta-at"), "%u:%f", &h, &d) != 2) return die("--theta-at wants HOP:DEG");
      events.push_back(bf_event{h, 0, 0, d});
    } else if (a == "--interf-at") {
      unsigned h, id; float d;
      if (sscanf(need("--interf-at"), "%u:%u:%f", &h, &id, &d) != 3) return die("--interf-at wants HOP:ID:DEG");
      events.push_back(bf_event{h, 1, id, d});
    } else if (a == "--help" || a == "-h") {
      printf("usage: bf_offline --algo A --config beamform_config.yaml [--launch F.launch] --in IN.wav --out OUT.wav [--hop 512] [--theta DEG]\n"
             "                  [--set key=value ...] [--theta-at HOP:DEG ...] [--interf-at HOP:ID:DEG ...] [--events FILE] [--raw M:SR] [--pcm16]\n");
      return 0;
    } else return die("unknown argument " + a);
  }
  const int algo_id = algo_of(algo);
  if (algo_id < 0) return die("--algo must be one of das mvdr lcmv gss phase phasempf mcra ref gsc");
  if (config.empty() || in_path.empty() || out_path.empty()) return die("--config, --in and --out are required (see --help)");
  if (!events_path.empty()) {
    std::ifstream ef(events_path);
    if (!ef) return die("cannot open " + events_path);
    std::string line;
    while (std::getline(ef, line)) {
      std::stringstream ss(line);
      unsigned h; std::string kind;
      if (!(ss >> h >> kind)) continue;
      bf_event e{h, 0, 0, 0.f};
      if (kind == "theta") { if (!(ss >> e.value)) return die("bad event line: " + line); }
      else if (kind == "interf") { e.kind = 1; if (!(ss >> e.id >> e.value)) return die("bad event line: " + line); }
      else return die("bad event line: " + line);
      events.push_back(e);
    }
  }
  std::stable_sort(events.begin(), events.end(), [](const bf_event& a, const bf_event& b) { return a.hop_index < b.hop_index; });

  // ---- input ----
  std::vector<float> planar;
  uint32_t channels = 0, rate = 0;
  size_t frames = 0;
  std::string err;
  if (raw_m) {
    std::ifstream f(in_path, std::ios::binary | std::ios::ate);
    if (!f) return die("cannot open " + in_path);
    const size_t bytes = (size_t)f.tellg();
    f.seekg(0);
    channels = raw_m; rate = raw_sr; frames = bytes / 4 / raw_m;
    planar.resize((size_t)raw_m * frames);
    f.read(reinterpret_cast<char*>(planar.data()), (std::streamsize)(planar.size() * 4));
  } else if (!read_wav(in_path, planar, channels, rate, frames, err)) {
    return die(err);
  }

  // ---- configuration: getParam fall-backs -> yaml -> launch file -> --set (the order a roslaunch would produce) ----
  bf_config cfg;
  if (bf_config_init(&cfg, algo_id) != BF_OK) return die(bf_last_error());
  cfg.hop = hop;
  cfg.sample_rate = rate;
  if (bf_config_load_yaml(&cfg, config.c_str()) != BF_OK) return die(bf_last_error());
  if (!launch.empty() && bf_config_load_launch(&cfg, launch.c_str()) != BF_OK) return die(bf_last_error());
  for (auto& kv : sets)
    if (bf_config_set(&cfg, kv.first.c_str(), kv.second.c_str()) != BF_OK) return die(bf_last_error());
  cfg.hop = hop;
  cfg.sample_rate = rate;   // JACK decides these at run time (rosjack.cpp:131-134), not the parameter files
  if (have_theta) cfg.initial_angle = theta;
  if (algo_id == BF_ALGO_REF && cfg.n_mics > 1) cfg.n_mics = 1;   // rosjack_ref opens one input (jack_ref.cpp:68)
  if ((uint32_t)cfg.n_mics > channels) {
    char m[128];
    snprintf(m, sizeof(m), "the file has %u channels, the configuration %d microphones", channels, cfg.n_mics);
    return die(m);
  }
  const uint32_t n_hops = (uint32_t)(frames / hop);
  const size_t L = (size_t)n_hops * hop;
  std::vector<float> out(L);
  bf_handle* h = nullptr;
  if (bf_create(&h, &cfg, 1) != BF_OK) return die(bf_last_error());
  // planar input [channels][frames]: microphone m starts at m*frames, the stream stride is irrelevant for one stream
  if (bf_process_batch(h, planar.data(), (size_t)channels * frames, frames, out.data(), L, n_hops, events.empty() ? nullptr : events.data(),
                       (uint32_t)events.size()) != BF_OK) {
    const std::string m = bf_last_error();
    bf_destroy(h);
    return die(m);
  }
  double interf[BF_MAX_INTERF];
  uint32_t n_interf = 0;
  bf_get_interferences(h, interf, BF_MAX_INTERF, &n_interf);
  bf_destroy(h);

  if (out_path.size() > 4 && out_path.substr(out_path.size() - 4) == ".f32") {
    FILE* f = fopen(out_path.c_str(), "wb");
    if (!f) return die("cannot write " + out_path);
    fwrite(out.data(), 4, L, f);
    fclose(f);
  } else if (!write_wav(out_path, out.data(), L, rate, pcm16, err)) {
    return die(err);
  }
  size_t loud = 0;   // output_to_rosjack warns at |sample| >= 1 and never clips (rosjack.cpp:372-374)
  for (float v : out) loud += std::fabs(v) >= 1.0f;
  printf("%s: %d microphones, %u hops of %u at %u Hz -> %s", algo.c_str(), cfg.n_mics, n_hops, hop, rate, out_path.c_str());
  if (loud) printf(" (%zu samples at or above full scale)", loud);
  if (n_interf) {
    printf("; interferers:");
    for (uint32_t k = 0; k < n_interf; k++) printf(" %g", interf[k]);
  }
  printf("\n");
  return 0;
}
