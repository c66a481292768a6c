// nnb_warp_desc.h -- weight layout descriptor of the 16-lanes-per-chain MCMC kernel (nnb_warp.cuh)
#pragma once
#include "../../include/nnb.h"

namespace nnb {

// Packed weights: per block, every (scale, translate) weight pair interleaved as float2
//   W1 [nin][16] f2 | b1 [16] f2 | L x { W2 [16 (k)][16 (j)] f2 | b2 [16] f2 } | W3 [16 (k)][NO] f2 | b3 [NO] f2
// with NO = round16(nout).  W1[a][j] = (Ws1[j][i0 + 2a], Wt1[j][i0 + 2a]), W2[k][j] = (Ws2[j][k], Wt2[j][k]),
// W3[k][o] = (Ws3[o0 + 2o][k], Wt3[o0 + 2o][k]).
struct WarpFlowDesc {
  int d, L, B;
  int total_floats;
  int off[NNB_MAX_BLOCKS];
};

}  // namespace nnb
