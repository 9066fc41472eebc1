// TEST INFRASTRUCTURE: offline transport behind the reference's rosjack.h declarations and the ROS façade.
// Included AFTER the node source in the same translation unit (rosjack.h defines its globals in the
// header, so everything lives in one TU).  Environment:
//   BFREF_PARAMS  text file: "key value" per line; "micN x y" for microphones
//   BFREF_IN      raw float32 [M][L];  BFREF_OUT raw float32 [L]
//   BFREF_EVENTS  optional text file: "<hop> theta <deg>" | "<hop> interf <id> <deg>", applied before that hop
//   BFREF_HOP     JACK period (default 512);  BFREF_SR sample rate (default 48000)
#include <fstream>
static int (*bfshim_callback)(jack_nframes_t, void*) = nullptr;
static std::vector<float> bfshim_in, bfshim_out;
static size_t bfshim_L = 0, bfshim_pos = 0;
static int bfshim_M = 0;

void bfshim::load_params() {
  const char* path = getenv("BFREF_PARAMS");
  if (!path) return;
  std::ifstream f(path);
  std::string line;
  while (std::getline(f, line)) {
    std::stringstream ss(line);
    std::string key;
    if (!(ss >> key)) continue;
    if (key.compare(0, 3, "mic") == 0 && key.size() > 3 && isdigit((unsigned char)key[3])) {
      double x = 0, y = 0;
      ss >> x >> y;
      std::map<std::string, double> m;
      m["id"] = atoi(key.c_str() + 3) + 1; m["x"] = x; m["y"] = y;
      ParamServer::get().maps[key] = m;
    } else {
      std::string val;
      ss >> val;
      ParamServer::get().scalars[key] = val;
    }
  }
}

int rosjack_create(int, ros::NodeHandle*, const char*, const char*, int input_number, int (*callback_function)(jack_nframes_t, void*)) {
  bfshim_callback = callback_function;
  bfshim_M = input_number;
  jack_num_inputs = input_number;
  rosjack_window_size = getenv("BFREF_HOP") ? (unsigned)atoi(getenv("BFREF_HOP")) : 512u;
  rosjack_sample_rate = getenv("BFREF_SR") ? (unsigned)atoi(getenv("BFREF_SR")) : 48000u;
  output_type = ROSJACK_OUT_JACK;
  return 0;
}
rosjack_data** input_from_rosjack(int) {
  static std::vector<rosjack_data*> ptrs;
  ptrs.resize(bfshim_M);
  for (int m = 0; m < bfshim_M; m++) ptrs[m] = bfshim_in.data() + (size_t)m * bfshim_L + bfshim_pos;
  return ptrs.data();
}
void output_to_rosjack(rosjack_data* data, int n, int) { bfshim_out.insert(bfshim_out.end(), data, data + n); }
void output_to_rosjack(rosjack_data* data, int n) { output_to_rosjack(data, n, 0); }

void bfshim::run_offline() {
  const char* in_path = getenv("BFREF_IN");
  const char* out_path = getenv("BFREF_OUT");
  if (!in_path || !out_path || !bfshim_callback) { fprintf(stderr, "bfshim: BFREF_IN/BFREF_OUT not set\n"); exit(2); }
  FILE* f = fopen(in_path, "rb");
  if (!f) { fprintf(stderr, "bfshim: cannot open %s\n", in_path); exit(2); }
  fseek(f, 0, SEEK_END);
  const size_t total = (size_t)ftell(f) / sizeof(float);
  fseek(f, 0, SEEK_SET);
  bfshim_in.resize(total);
  if (fread(bfshim_in.data(), sizeof(float), total, f) != total) { fprintf(stderr, "bfshim: short read\n"); exit(2); }
  fclose(f);
  bfshim_L = total / (size_t)bfshim_M;
  struct Ev { unsigned hop; int kind; unsigned id; float val; };
  std::vector<Ev> evs;
  if (const char* evp = getenv("BFREF_EVENTS")) {
    std::ifstream ef(evp);
    std::string line;
    while (std::getline(ef, line)) {
      std::stringstream ss(line);
      Ev e; std::string kind;
      if (!(ss >> e.hop >> kind)) continue;
      if (kind == "theta") { e.kind = 0; e.id = 0; ss >> e.val; } else { e.kind = 1; ss >> e.id >> e.val; }
      evs.push_back(e);
    }
  }
  const unsigned H = rosjack_window_size;
  const unsigned T = (unsigned)(bfshim_L / H);
  size_t e = 0;
  for (unsigned t = 0; t < T; t++) {
    while (e < evs.size() && evs[e].hop <= t) {
      if (evs[e].kind == 0) { if (Topics::get().theta) Topics::get().theta(evs[e].val); }
      else if (Topics::get().interf) Topics::get().interf((unsigned short)evs[e].id, evs[e].val);
      e++;
    }
    bfshim_pos = (size_t)t * H;
    bfshim_callback(H, nullptr);
  }
  while (e < evs.size()) {
    if (evs[e].kind == 0) { if (Topics::get().theta) Topics::get().theta(evs[e].val); }
    else if (Topics::get().interf) Topics::get().interf((unsigned short)evs[e].id, evs[e].val);
    e++;
  }
  FILE* o = fopen(out_path, "wb");
  fwrite(bfshim_out.data(), sizeof(float), bfshim_out.size(), o);
  fclose(o);
  if (const char* ip = getenv("BFREF_INTERF_OUT")) {   // final interference list, for the bit-exact check
    FILE* q = fopen(ip, "w");
    for (double a : interference_angles) fprintf(q, "%.17g\n", a);
    fclose(q);
  }
}
