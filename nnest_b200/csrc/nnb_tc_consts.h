// nnb_tc_consts.h -- the block of warp-uniform float32 constants that mcmc_tc_kernel receives as a kernel parameter
// (constant bank 0); layout and rationale: nnb_tc_kernels.cuh.
#pragma once
namespace nnb {
constexpr int kTcConstFloats = 896;   // with the other parameters: below the 4 KB of the classic parameter space (see nnb_tc_launch.cuh)
struct TcConsts {
  float v[kTcConstFloats];
};
}  // namespace nnb
