// tensor-core MCMC kernel with x_dim = 2 fixed at compile time (see nnb_tc_launch.cuh)
#define NNB_TC_DIM 2
#include "nnb_tc_fixed.inc"
