// mcra_b200.cpp - the reference's mcra node with its DSP on the B200 (drop-in for mcra.cpp; see node_b200.h).
#define BF_NODE_ALGO BF_ALGO_MCRA
#define BF_NODE_INTERF 0
#define BF_NODE_KEYS { {"alphaS", 'd'}, {"alphaD", 'd'}, {"alphaD2", 'd'}, {"delta", 'd'}, {"L", 'i'}, {"out_amp", 'd'}, {"out_only_noise", 'b'} }   /* rosparam keys of mcra.cpp:176-226 */
#include "node_b200.h"
