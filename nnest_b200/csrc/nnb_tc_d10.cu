// tensor-core MCMC kernel with x_dim = 10 fixed at compile time (see nnb_tc_launch.cuh)
#define NNB_TC_DIM 10
#include "nnb_tc_fixed.inc"
