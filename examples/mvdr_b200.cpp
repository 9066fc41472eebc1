// mvdr_b200.cpp - the reference's mvdr node with its DSP on the B200 (drop-in for mvdr.cpp; see node_b200.h).
#define BF_NODE_ALGO BF_ALGO_MVDR
#define BF_NODE_INTERF 0
#define BF_NODE_KEYS { {"past_windows", 'd'}, {"freq_mag_threshold", 'd'}, {"freq_max", 'd'}, {"freq_min", 'd'}, {"out_amp", 'd'} }   /* rosparam keys of mvdr.cpp:146-187 */
#include "node_b200.h"
