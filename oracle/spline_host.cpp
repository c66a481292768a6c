// spline_host.cpp -- TEST INFRASTRUCTURE ONLY: runs the per-sample arithmetic of nnest_b200/csrc/nnb_spline.cuh on the CPU
// (the header is __host__ __device__) so that it can be checked against goldens recorded from the reference before a CUDA
// kernel is wrapped around it.  Built by oracle/build_spline_host.py with g++ into oracle/_build/.
#include <stdint.h>

#include <vector>

#include "../nnest_b200/csrc/nnb_spline.cuh"

extern "C" int spline_host_flow(const float* packed, int d, int hidden, int blocks, int num_bins, float bound, int inverse,
                                const float* in, float* out, float* logdet, int64_t n) {
  using namespace nnb::spline;
  if (d < 2 || d > 128 || hidden > kMaxHidden || num_bins > kMaxBins) return -1;
  Shape sh{d, hidden, blocks, num_bins, bound};
  std::vector<float> tmp((size_t)d);
  for (int64_t r = 0; r < n; ++r) {
    float* x = out + r * d;
    for (int i = 0; i < d; ++i) x[i] = in[r * d + i];
    logdet[r] = inverse ? flow_inverse(sh, packed, x, tmp.data()) : flow_forward(sh, packed, x, tmp.data());
  }
  return 0;
}

extern "C" int spline_host_block_floats(int d, int hidden, int num_bins) {
  nnb::spline::Shape sh{d, hidden, 1, num_bins, 3.f};
  return sh.block_floats();
}
