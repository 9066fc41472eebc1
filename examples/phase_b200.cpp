// phase_b200.cpp - the reference's phase node with its DSP on the B200 (drop-in for phase.cpp; see node_b200.h).
#define BF_NODE_ALGO BF_ALGO_PHASE
#define BF_NODE_INTERF 0
#define BF_NODE_KEYS { {"min_phase", 'd'}, {"mag_mult", 'd'}, {"mag_threshold", 'd'} }   /* rosparam keys of phase.cpp:165-191 */
#include "node_b200.h"
