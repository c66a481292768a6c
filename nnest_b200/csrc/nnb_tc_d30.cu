// tensor-core MCMC kernel with x_dim = 30 fixed at compile time (see nnb_tc_launch.cuh)
#define NNB_TC_DIM 30
#include "nnb_tc_fixed.inc"
