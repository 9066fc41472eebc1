// jack_ref_b200.cpp - the reference's jack_ref node with its DSP on the B200 (drop-in for jack_ref.cpp; see node_b200.h).
#define BF_NODE_ALGO BF_ALGO_REF
#define BF_NODE_INTERF 0
#define BF_NODE_KEYS { {nullptr, 0} }   /* rosparam keys of jack_ref.cpp */
#include "node_b200.h"
