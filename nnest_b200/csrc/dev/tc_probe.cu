// tc_probe.cu -- development probe (not part of libnnb.so): checks on real hardware that the TMEM-A /
// smem-B layouts and the 3xTF32 scheme of nnb_tc.cuh compute D = A * B^T to FP32 accuracy.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o nnest_b200/lib/tc_probe nnest_b200/csrc/dev/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../nnb_tc.cuh"

using namespace nnb::tc;

// A: (128, K) row-major global; Bhi/Blo: packed (N*K floats each); D: (128, N) row-major
template <int MODE>  // 0: single TF32 (hi*hi only), 1: 3xTF32
__global__ void probe(const float* A, const float* Bhi, const float* Blo, float* D, int K, int N) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t mbar;
  float* bh = sm;
  float* bl = sm + N * K;
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) { bh[i] = Bhi[i]; bl[i] = Blo[i]; }
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 128);
  if (threadIdx.x == 0) { mbar_init(&mbar, 1); mbar_fence_init(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tb = tmem_base;
  const uint32_t lane_addr = tb + (((uint32_t)(warp & 3) * 32u) << 16);
  // A hi -> cols [0,32), lo -> cols [32,64); D -> cols [64, 64+N)
  for (int k0 = 0; k0 < K; k0 += 8) {
    uint32_t hi[8], lo[8];
    for (int j = 0; j < 8; ++j) split_tf32(A[threadIdx.x * K + k0 + j], hi[j], lo[j]);
    tmem_st8(lane_addr + k0, hi);
    tmem_st8(lane_addr + 32 + k0, lo);
  }
  wait_st();
  fence_before_sync();
  // make the generic-proxy smem writes of B visible to the async (tensor core) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    fence_after_sync();
    if (MODE == 1) {
      mma_3xtf32(tb + 64, tb, tb + 32, smem_u32(bh), smem_u32(bl), K / 8, N, false);
    } else {
      const uint32_t nN = N >> 3;
      for (int ks = 0; ks < K / 8; ++ks)
        mma_tf32_ts(tb + 64, tb + 8 * ks, smem_desc_kmajor(smem_u32(bh) + ks * 2 * nN * 128, nN * 128, 128),
                    idesc_tf32_m128(N), ks > 0);
    }
    mma_commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  fence_after_sync();
  for (int n0 = 0; n0 < N; n0 += 8) {
    uint32_t r[8];
    tmem_ld8(lane_addr + 64 + n0, r);
    wait_ld();
    for (int j = 0; j < 8; ++j) D[threadIdx.x * N + n0 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 128);
}

int main() {
  int fails = 0;
  const int cfgs[][2] = {{8, 16}, {16, 16}, {16, 32}, {32, 32}, {24, 16}, {32, 16}};
  for (auto& c : cfgs) {
    const int K = c[0], N = c[1];
    std::vector<float> A(128 * K), W(N * K), Bh(N * K), Bl(N * K), D(128 * N);
    srand(K * 100 + N);
    for (auto& v : A) v = 2.f * rand() / RAND_MAX - 1.f;
    for (auto& v : W) v = 2.f * rand() / RAND_MAX - 1.f;
    host_pack_b(W.data(), N, K, K, N, K, Bh.data(), Bl.data());
    float *dA, *dBh, *dBl, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dBh, Bh.size() * 4); cudaMalloc(&dBl, Bl.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dBh, Bh.data(), Bh.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dBl, Bl.data(), Bl.size() * 4, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; ++mode) {
      cudaMemset(dD, 0, D.size() * 4);
      if (mode == 0) probe<0><<<1, 128, 2 * N * K * 4>>>(dA, dBh, dBl, dD, K, N);
      else probe<1><<<1, 128, 2 * N * K * 4>>>(dA, dBh, dBl, dD, K, N);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("K=%d N=%d mode=%d CUDA error %s\n", K, N, mode, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)W[n * K + k];
          maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
          maxref = fmax(maxref, fabs(ref));
        }
      double rel = maxerr / maxref;
      bool ok = mode == 0 ? rel < 5e-3 : rel < 2e-6;
      printf("K=%2d N=%2d %s  max_abs_err %.3e  rel %.3e  %s\n", K, N, mode ? "3xTF32" : "1xTF32", maxerr, rel,
             ok ? "OK" : "FAIL");
      fails += !ok;
    }
    cudaFree(dA); cudaFree(dBh); cudaFree(dBl); cudaFree(dD);
  }
  printf(fails ? "PROBE FAILED\n" : "PROBE OK\n");
  return fails ? 1 : 0;
}
