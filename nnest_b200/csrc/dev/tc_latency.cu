// tc_latency.cu -- development probe: cycles of one tensor-core round trip as used by mcmc_tc_kernel
// (tcgen05.st A -> fence/barrier -> issue n MMAs -> commit -> mbarrier wait -> barrier -> tcgen05.ld D).
#include <cstdio>
#include <vector>
#include "../nnb_tc.cuh"
using namespace nnb::tc;

template <int NWARPS_WAIT_ALL>
__global__ void rt_kernel(long long* out, int nmma_groups, int n, int ksteps, int iters, int threads_per_tile) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t mbar;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 0.001f * (i % 7);
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 128);
  if (threadIdx.x == 0) { mbar_init(&mbar, 1); mbar_fence_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base;
  const uint32_t lane_addr = tb + (((uint32_t)(warp & 3) * 32u) << 16);
  uint32_t phase = 0;
  long long t_st = 0, t_issue = 0, t_wait = 0, t_ld = 0;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    long long c0 = clock64();
    uint32_t hi[8], lo[8];
    for (int j = 0; j < 8; ++j) split_tf32(0.01f * (threadIdx.x + j + it) + acc * 1e-9f, hi[j], lo[j]);
    tmem_st8(lane_addr, hi); tmem_st8(lane_addr + 8, hi); tmem_st8(lane_addr + 32, lo); tmem_st8(lane_addr + 40, lo);
    wait_st();
    fence_before_sync();
    named_bar_sync(1, threads_per_tile);
    long long c1 = clock64();
    if (warp == 0) {
      fence_after_sync();
      if (elect_one()) {
        for (int g = 0; g < nmma_groups; ++g)
          mma_3xtf32(tb + 64 + 16 * (g & 1), tb, tb + 32, smem_u32(sm) + 4096 * g, smem_u32(sm) + 4096 * g + 2048, ksteps, n, false);
        mma_commit(&mbar);
      }
      __syncwarp();
    }
    long long c2 = clock64();
    if (NWARPS_WAIT_ALL || warp == 0) mbar_wait(&mbar, phase);
    phase ^= 1u;
    __syncwarp();
    if (!NWARPS_WAIT_ALL) named_bar_sync(1, threads_per_tile);
    fence_after_sync();
    long long c3 = clock64();
    uint32_t r[8];
    tmem_ld8(lane_addr + 64, r);
    wait_ld();
    acc += __uint_as_float(r[0]);
    long long c4 = clock64();
    if (it >= 10) { t_st += c1 - c0; t_issue += c2 - c1; t_wait += c3 - c2; t_ld += c4 - c3; }
  }
  if (threadIdx.x == 0) { out[0] = t_st; out[1] = t_issue; out[2] = t_wait; out[3] = t_ld; out[4] = (long long)acc; }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 128);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  long long h[5];
  const int iters = 210;
  struct Cfg { int groups, n, ks, threads; const char* name; } cfgs[] = {
      {1, 32, 2, 128, "L1: 6 MMA N=32 (128 thr)"}, {2, 16, 2, 128, "L2/L3: 12 MMA N=16 (128 thr)"},
      {1, 32, 2, 256, "L1: 6 MMA N=32 (256 thr)"}, {2, 16, 2, 256, "L2/L3: 12 MMA N=16 (256 thr)"},
      {1, 16, 1, 128, "3 MMA N=16 (128 thr)"}, {0, 16, 1, 128, "0 MMA, commit only (128 thr)"}};
  for (auto& c : cfgs) {
    for (int mode = 0; mode < 2; ++mode) {
      if (mode == 0) rt_kernel<0><<<1, c.threads, 40960>>>(d, c.groups, c.n, c.ks, iters, c.threads);
      else rt_kernel<1><<<1, c.threads, 40960>>>(d, c.groups, c.n, c.ks, iters, c.threads);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 40, cudaMemcpyDeviceToHost);
      double k = 1.0 / (iters - 10);
      printf("%-32s %s  st+bar %6.0f  issue %6.0f  commit->wake %6.0f  ld %5.0f  total %6.0f cycles\n", c.name,
             mode ? "all-warps-poll " : "one-warp-polls", h[0] * k, h[1] * k, h[2] * k, h[3] * k, (h[0] + h[1] + h[2] + h[3]) * k);
    }
  }
  return 0;
}
