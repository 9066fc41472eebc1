// TEST INFRASTRUCTURE: one translation unit = the reference node source, unmodified and included from
// where it lies (-DNODE_SRC="/root/reference/beamform/src/<node>.cpp"), followed by the offline driver.
#include NODE_SRC
#include "shim_driver.inl"
