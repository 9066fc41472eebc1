// das_b200.cpp - the reference's das node with its DSP on the B200 (drop-in for das.cpp; see node_b200.h).
#define BF_NODE_ALGO BF_ALGO_DAS
#define BF_NODE_INTERF 0
#define BF_NODE_KEYS { {nullptr, 0} }   /* rosparam keys of das.cpp */
#include "node_b200.h"
