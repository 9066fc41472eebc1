// phasempf_b200.cpp - the reference's phasempf node with its DSP on the B200 (drop-in for phasempf.cpp; see node_b200.h).
#define BF_NODE_ALGO BF_ALGO_PHASEMPF
#define BF_NODE_INTERF 0
#define BF_NODE_KEYS { {"min_phase", 'd'}, {"min_mag", 'd'}, {"smooth_size", 'i'}, {"MCRA_alphaS", 'd'}, {"MCRA_alphaD", 'd'}, {"MCRA_alphaD2", 'd'}, {"MCRA_delta", 'd'}, {"MCRA_L", 'i'}, {"MPF_alphaS", 'd'}, {"MPF_eta", 'd'}, {"MPF_rev_gamma", 'd'}, {"MPF_rev_delta", 'd'}, {"out_amp", 'd'}, {"noise_floor", 'd'}, {"out_only_noise", 'b'}, {"out_only_mcra", 'b'} }   /* rosparam keys of phasempf.cpp:355-472 */
#include "node_b200.h"
