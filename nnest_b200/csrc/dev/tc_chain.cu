// tc_chain.cu -- development probe: how long does a chain of small tcgen05.mma (M = 128, N = 16 / 32, K = 8, tf32, A in
// TMEM) take from the first issue to the mbarrier wake-up, as a function of (a) the number of MMAs, (b) whether they
// accumulate into the SAME D columns (dependent chain, as in 3xTF32) or into different ones?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o nnest_b200/lib/tc_chain nnest_b200/csrc/dev/tc_chain.cu
#include <cstdio>
#include "../nnb_tc.cuh"
using namespace nnb::tc;

__global__ void chain_kernel(long long* out, int count, int n, int dstride, int iters) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t mbar;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 0.001f * (i % 7);
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (threadIdx.x == 0) { mbar_init(&mbar, 1); mbar_fence_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base;
  const uint32_t lane_addr = tb + (((uint32_t)(warp & 3) * 32u) << 16);
  uint32_t v[8];
  for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(0.01f * (threadIdx.x + j)) & 0xffffe000u;
  for (int c = 0; c < 64; c += 8) tmem_st8(lane_addr + c, v);
  wait_st();
  fence_before_sync();
  __syncthreads();
  uint32_t phase = 0;
  long long t_issue = 0, t_wake = 0;
  const uint32_t idesc = idesc_tf32_m128(n);
  const uint64_t desc = smem_desc_kmajor(smem_u32(sm), (n / 8) * 128u, 128u);
  for (int it = 0; it < iters; ++it) {
    if (warp == 0) {
      fence_after_sync();
      long long c0 = clock64(), c1 = 0;
      if (elect_one()) {
        for (int g = 0; g < count; ++g) mma_tf32_ts(tb + 64 + dstride * g, tb + 8 * (g & 3), desc, idesc, g > 0 && dstride == 0);
        mma_commit(&mbar);
        c1 = clock64();
      }
      __syncwarp();
      mbar_wait(&mbar, phase);
      long long c2 = clock64();
      if (it >= 10 && threadIdx.x == 0) { t_wake += c2 - c0; }
      c1 = __shfl_sync(0xffffffffu, c1, 0);   // (elected lane is lane 0 in practice; only used as an indication)
      if (it >= 10 && threadIdx.x == 0 && c1) { t_issue += c1 - c0; }
    }
    phase ^= 1u;
    fence_before_sync();
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = t_issue; out[1] = t_wake; }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  long long h[2];
  const int iters = 210;
  for (int n : {16, 32}) {
    for (int dstride : {0, 32}) {
      for (int count : {0, 1, 2, 3, 6, 12}) {
        if (dstride && 64 + dstride * count > 512) continue;
        chain_kernel<<<1, 128, 40960>>>(d, count, n, dstride, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        double k = 1.0 / (iters - 10);
        printf("N=%2d  %-22s  %2d MMAs: issue %6.0f  first issue -> wake %6.0f cycles\n", n,
               dstride ? "independent D columns" : "same D (accumulate)", count, h[0] * k, h[1] * k);
      }
    }
  }
  return 0;
}
