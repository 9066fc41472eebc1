// gsc_b200.cpp - the reference's gsc node with its DSP on the B200 (drop-in for gsc.cpp; see node_b200.h).
#define BF_NODE_ALGO BF_ALGO_GSC
#define BF_NODE_INTERF 0
#define BF_NODE_KEYS { {"use_vad", 'b'}, {"vad_threshold", 'd'}, {"mu0", 'd'}, {"mu_max", 'd'}, {"filter_size", 'i'} }   /* rosparam keys of gsc.cpp:199-260 */
#include "node_b200.h"
