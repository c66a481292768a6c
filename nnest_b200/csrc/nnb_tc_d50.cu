// tensor-core MCMC kernel with x_dim = 50 fixed at compile time (see nnb_tc_launch.cuh)
#define NNB_TC_DIM 50
#include "nnb_tc_fixed.inc"
