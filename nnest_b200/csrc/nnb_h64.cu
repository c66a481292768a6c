// Kernels of libnnb.so specialised for hidden_dim = 64.
#include "nnb_launch.inc"

template struct LaunchH<64>;
