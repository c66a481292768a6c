// Kernels of libnnb.so specialised for hidden_dim = 32.
#include "nnb_launch.inc"

template struct LaunchH<32>;
