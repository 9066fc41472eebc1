// Common body of the drop-in nodes (examples/<node>_b200.cpp): what a maintainer of balkce/beamform adds to move a node's
// DSP to the GPU.  The reference's rosjack.h / util.h are used UNCHANGED for the ROS parameter
// handling (handle_params, util.h:52-134), the JACK client (rosjack_create) and the transport; only the body of the JACK
// callback goes through the C ABI of include/beamform_b200.h.  Each node source defines
//   BF_NODE_ALGO   one of BF_ALGO_*
//   BF_NODE_KEYS   the rosparam keys its <algo>_handle_params reads (e.g. mvdr.cpp:146-187), as {"name", 'd'|'i'|'b'}
//   BF_NODE_INTERF 1 when the node subscribes to theta_interference (lcmv.cpp:320, gss.cpp:350)
// before including this file.  -DBF_NODE_SEAM binds at the per-frame operator instead (weight_func of util.h:289):
// util.h's do_overlap keeps the ring buffers and the overlap-add, bf_apply_weights replaces apply_weights.
#include "rosjack.h"
#include "util.h"

#include <complex>
#include <string>

#include <beamform_b200.h>

#ifndef BF_NODE_INTERF
#define BF_NODE_INTERF 0
#endif

struct bf_node_key { const char *name; char type; };
static const bf_node_key bf_node_keys[] = BF_NODE_KEYS;

static bf_handle *bf = nullptr;
static int bf_node_smooth_size = 1;
bool READY = false;

#ifdef BF_NODE_SEAM
// apply_weights (das.cpp:47-70 and its five siblings): frame of fft_win samples per microphone in, windowed frame out
void apply_weights(jack_ringbuffer_t **in, rosjack_data *out) {
  static std::vector<std::vector<rosjack_data> > frames;
  static std::vector<const float *> ptrs;
  frames.resize(number_of_microphones);
  ptrs.resize(number_of_microphones);
  for (int i = 0; i < number_of_microphones; i++) {
    frames[i].resize(fft_win);
    jack_ringbuffer_get_read_vector(in[i], readring);   // the two chunks of util.h:217-242, without the window (the library applies it)
    size_t n0 = readring[0].len / sizeof(rosjack_data), n1 = readring[1].len / sizeof(rosjack_data);
    memcpy(frames[i].data(), readring[0].buf, n0 * sizeof(rosjack_data));
    memcpy(frames[i].data() + n0, readring[1].buf, n1 * sizeof(rosjack_data));
    ptrs[i] = frames[i].data();
  }
  if (bf_apply_weights(bf, ptrs.data(), out, fft_win) != BF_OK) {
    ROS_ERROR("%s", bf_last_error());
    for (unsigned j = 0; j < fft_win; j++) out[j] = 0.0;
  }
}
#endif

int jack_callback(jack_nframes_t nframes, void *arg) {
  rosjack_data out[nframes];
  rosjack_data **in = input_from_rosjack(nframes);              // rosjack.cpp:538 (JACK port buffers)
  if (!READY) {
    for (jack_nframes_t i = 0; i < nframes; i++) out[i] = 0.0;   // das.cpp:81-85
  } else {
#ifdef BF_NODE_SEAM
    do_overlap(in, out, nframes, apply_weights);                 // util.h:289-314, unchanged
    if (BF_NODE_ALGO == BF_ALGO_PHASEMPF) {
      // at this seam the output smoother stays in the callback, as in phasempf.cpp:331-334: every sample becomes the mean
      // of the last smooth_size overlap-added samples (zero-initialised double history, phasempf.cpp:510)
      static std::vector<double> past(bf_node_smooth_size, 0.0);
      static size_t head = 0;
      for (jack_nframes_t j = 0; j < nframes; j++) {
        past[head] = out[j];
        head = (head + 1) % past.size();
        double sum = 0.0;
        for (size_t q = 0; q < past.size(); q++) sum += past[(head + q) % past.size()];   // oldest first, like get_mean
        out[j] = sum / (double)past.size();
      }
    }
#else
    if (bf_process_hop(bf, (const float *const *)in, out, nframes) != BF_OK) {
      ROS_ERROR("%s", bf_last_error());
      for (jack_nframes_t i = 0; i < nframes; i++) out[i] = 0.0;
    }
#endif
#if BF_NODE_INTERF
    {   // keep the node's own copy of the list (logging, util.h:36) in step with the library's
      double a[BF_MAX_INTERF];
      uint32_t k = 0;
      if (bf_get_interferences(bf, a, BF_MAX_INTERF, &k) == BF_OK) interference_angles.assign(a, a + k);
    }
#endif
  }
  output_to_rosjack(out, nframes, output_type);                  // rosjack.cpp:351
  return 0;
}

void theta_roscallback(const std_msgs::Float32::ConstPtr &msg) {
  ROS_INFO("Updating weights for angle: %f", msg->data);
  bf_set_theta(bf, msg->data);                                   // das.cpp:94-99: angle = msg->data; update_weights();
}

#if BF_NODE_INTERF
void interf_theta_roscallback(const beamform::InterfTheta::ConstPtr &msg) {
  bf_set_interference(bf, msg->id, msg->angle);                  // lcmv.cpp:258-309: the list logic lives in the library
}
#endif

int main(int argc, char *argv[]) {
  ros::init(argc, argv, client_name);
  ros::NodeHandle n;
  handle_params(&n);                                             // util.h:52-134: angle, array_geometry, interference_angles
  ros::Subscriber theta_subscriber = n.subscribe("theta", 1000, theta_roscallback);
#if BF_NODE_INTERF
  ros::Subscriber interf_subscriber = n.subscribe("theta_interference", 1000, interf_theta_roscallback);
#endif
  // rosjack_ref opens a single JACK input (jack_ref.cpp:68)
  const int n_inputs = (BF_NODE_ALGO == BF_ALGO_REF) ? 1 : number_of_microphones;
  if (rosjack_create(ROSJACK_READ, &n, "jackaudio", client_name, n_inputs, jack_callback)) {
    ROS_ERROR("JACK agent could not be created.");
    ros::shutdown();
    exit(1);
  }

  bf_config cfg;
  bf_config_init(&cfg, BF_NODE_ALGO);                            // the getParam fall-backs of <algo>_handle_params
  cfg.sample_rate = rosjack_sample_rate;                         // rosjack.cpp:134
  cfg.hop = rosjack_window_size;                                 // JACK period; fft_win = 2 * hop (util.h:261)
  cfg.initial_angle = angle;
  cfg.n_mics = n_inputs;                                         // the library reads every port the node opened
  for (int i = 0; i < n_inputs; i++) {
    // util.h:116-119 re-referenced x,y to microphone 0 after computing dist/angle from the raw values (SURVEY B-6):
    // the library wants the RAW yaml coordinates
    cfg.mic_x[i] = array_geometry[i]["x"] + (i ? array_geometry[0]["x"] : 0.0);
    cfg.mic_y[i] = array_geometry[i]["y"] + (i ? array_geometry[0]["y"] : 0.0);
  }
  cfg.n_angle_interf = (int)interference_angles.size();
  for (size_t k = 0; k < interference_angles.size() && k < BF_MAX_INTERF; k++) cfg.angle_interf[k] = interference_angles[k];
  const std::string node_name = ros::this_node::getName();
  for (const bf_node_key &k : bf_node_keys) {                    // same keys, same types as the node's own handle_params
    if (!k.name) continue;
    char val[64];
    bool have = false;
    if (k.type == 'b') { bool v; if ((have = n.getParam(node_name + "/" + k.name, v))) snprintf(val, sizeof(val), "%s", v ? "true" : "false"); }
    else if (k.type == 'i') { int v; if ((have = n.getParam(node_name + "/" + k.name, v))) snprintf(val, sizeof(val), "%d", v); }
    else { double v; if ((have = n.getParam(node_name + "/" + k.name, v))) snprintf(val, sizeof(val), "%.17g", v); }
    if (have) bf_config_set(&cfg, k.name, val);
  }
  bf_node_smooth_size = cfg.smooth_size;
  if (bf_create(&bf, &cfg, 1) != BF_OK) {                        // no CPU fallback: without a B200 the node gives up like a failed rosjack_create
    ROS_ERROR("%s", bf_last_error());
    ros::shutdown();
    exit(1);
  }
#ifdef BF_NODE_SEAM
  prepare_overlap_and_add();                                     // util.h:257-287: ring buffers, out_buff
#endif
  READY = true;
  ros::spin();
  bf_destroy(bf);
  exit(0);
}
