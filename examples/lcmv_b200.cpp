// lcmv_b200.cpp - the reference's lcmv node with its DSP on the B200 (drop-in for lcmv.cpp; see node_b200.h).
#define BF_NODE_ALGO BF_ALGO_LCMV
#define BF_NODE_INTERF 1
#define BF_NODE_KEYS { {"past_windows", 'd'}, {"freq_mag_threshold", 'd'}, {"freq_max", 'd'}, {"freq_min", 'd'}, {"out_amp", 'd'}, {"interf_angle_threshold", 'd'} }   /* rosparam keys of lcmv.cpp:170-219 */
#include "node_b200.h"
