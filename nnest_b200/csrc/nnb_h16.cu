// Kernels of libnnb.so specialised for hidden_dim = 16.
#include "nnb_launch.inc"

template struct LaunchH<16>;
