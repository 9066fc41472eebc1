// gss_b200.cpp - the reference's gss node with its DSP on the B200 (drop-in for gss.cpp; see node_b200.h).
#define BF_NODE_ALGO BF_ALGO_GSS
#define BF_NODE_INTERF 1
#define BF_NODE_KEYS { {"freq_mag_threshold", 'd'}, {"freq_max", 'd'}, {"freq_min", 'd'}, {"out_amp", 'd'}, {"interf_angle_threshold", 'd'}, {"mu", 'd'}, {"lambda", 'd'} }   /* rosparam keys of gss.cpp:177-240 */
#include "node_b200.h"
