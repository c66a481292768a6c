// nnb_stats.cu -- chain diagnostics on the device trace (reference nnest/utils/evaluation.py:6-73, called from
// Sampler._chain_stats, nnest/sampler.py:474-492): acceptance rate, mean jump distance, per-dimension moments and the
// lagged autocorrelation sums behind the effective sample size.  The reference walks chains x steps in Python; here the
// trace never leaves the GPU.  Layout: trace_x float32 [T][d][N] (chain-minor, as written by nnb_mcmc_run); statistics are
// taken on v = x * t_scale + t_shift in float64 (the reference evaluates them on the transformed float64 samples).
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "nnb_host.h"

namespace {

constexpr int kStatThreads = 128;
constexpr int kLagBlock = 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one thread per chain: steps whose point differs from the previous one in any coordinate (evaluation.py:42-56) and the
// Euclidean length of every step (evaluation.py:59-73).  out[0] += moved steps, out[1] += sum of step lengths
__global__ void __launch_bounds__(kStatThreads) chain_moves_kernel(const float* __restrict__ x, int T, int d, long long n,
                                                                   const double* __restrict__ ts, double* __restrict__ out) {
  const long long c = (long long)blockIdx.x * kStatThreads + threadIdx.x;
  double cnt = 0.0, jump = 0.0;
  if (c < n) {
    for (int t = 1; t < T; ++t) {
      const float* a = x + ((size_t)(t - 1) * d) * n + c;
      const float* b = a + (size_t)d * n;
      bool moved = false;
      double d2 = 0.0;
      for (int i = 0; i < d; ++i) {
        const float av = a[(size_t)i * n], bv = b[(size_t)i * n];
        moved |= av != bv;
        const double df = ((double)bv - (double)av) * ts[i];
        d2 = fma(df, df, d2);
      }
      cnt += moved ? 1.0 : 0.0;
      jump += sqrt(d2);
    }
  }
  cnt = warp_sum(cnt);
  jump = warp_sum(jump);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, cnt);
    atomicAdd(out + 1, jump);
  }
}

// grid (chain blocks, d): out[i] += sum v, out[d + i] += sum v^2 over all chains and steps of dimension i
__global__ void __launch_bounds__(kStatThreads) chain_moments_kernel(const float* __restrict__ x, int T, int d, long long n,
                                                                     const double* __restrict__ ts,
                                                                     const double* __restrict__ tb, double* __restrict__ out) {
  const int i = blockIdx.y;
  const long long c = (long long)blockIdx.x * kStatThreads + threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  if (c < n) {
    const double a = ts[i], b = tb[i];
    for (int t = 0; t < T; ++t) {
      const double v = fma((double)x[((size_t)t * d + i) * n + c], a, b);
      s1 += v;
      s2 = fma(v, v, s2);
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + i, s1);
    atomicAdd(out + d + i, s2);
  }
}

// grid (chain blocks, d): out[l * d + i] += sum over chains and t of y[t] * y[t - lag0 - l], y = v - mean, for the lags
// lag0 .. lag0 + 31 (evaluation.py:6-14 before the division by (T - s), the variance and the number of chains).  One
// thread walks one chain's series once, keeping the last 32 delayed values in a register ring (fully unrolled).
__global__ void __launch_bounds__(kStatThreads) chain_autocorr_kernel(const float* __restrict__ x, int T, int d, long long n,
                                                                      const double* __restrict__ ts,
                                                                      const double* __restrict__ tb,
                                                                      const double* __restrict__ mu, int lag0,
                                                                      double* __restrict__ out) {
  const int i = blockIdx.y;
  const long long c = (long long)blockIdx.x * kStatThreads + threadIdx.x;
  double acc[kLagBlock], w[kLagBlock];
#pragma unroll
  for (int l = 0; l < kLagBlock; ++l) acc[l] = w[l] = 0.0;
  if (c < n) {
    const double a = ts[i], b = tb[i] - mu[i];
    const float* xi = x + (size_t)i * n + c;
    const size_t st = (size_t)d * n;
    for (int t0 = lag0; t0 < T; t0 += kLagBlock) {
#pragma unroll
      for (int r = 0; r < kLagBlock; ++r) {
        const int t = t0 + r;
        const bool ok = t < T;
        const double cur = ok ? fma((double)xi[(size_t)t * st], a, b) : 0.0;
        w[r] = ok ? fma((double)xi[(size_t)(t - lag0) * st], a, b) : 0.0;
#pragma unroll
        for (int l = 0; l < kLagBlock; ++l) acc[l] = fma(cur, w[(r - l + kLagBlock) % kLagBlock], acc[l]);
      }
    }
  }
#pragma unroll
  for (int l = 0; l < kLagBlock; ++l) {
    const double v = warp_sum(acc[l]);
    if ((threadIdx.x & 31) == 0) atomicAdd(out + (size_t)l * d + i, v);
  }
}

// workspace: [ts d][tb d][mu d][out 34 d + 2] doubles
int stats_ws(nnb_handle* h, int d, const double* t_scale, const double* t_shift, const double* mean, size_t out_doubles,
             cudaStream_t st) {
  const size_t need = (size_t)(3 + kLagBlock + 2) * NNB_MAX_DIM + 2;
  if (!h->d_stats_ws) NNB_CUDA(h, cudaMalloc(&h->d_stats_ws, need * sizeof(double)));
  std::vector<double> host((size_t)3 * d);
  for (int i = 0; i < d; ++i) {
    host[i] = t_scale ? t_scale[i] : 1.0;
    host[d + i] = t_shift ? t_shift[i] : 0.0;
    host[2 * d + i] = mean ? mean[i] : 0.0;
  }
  NNB_CUDA(h, cudaMemcpyAsync(h->d_stats_ws, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  NNB_CUDA(h, cudaStreamSynchronize(st));   // `host` goes out of scope
  NNB_CUDA(h, cudaMemsetAsync(h->d_stats_ws + 3 * d, 0, out_doubles * sizeof(double), st));
  return NNB_OK;
}

}  // namespace

extern "C" int nnb_chain_stats(nnb_handle* h, const float* trace_x, int64_t T, int d, int64_t n, const double* t_scale,
                               const double* t_shift, double* moved_out, double* jump_sum_out, double* sum_out,
                               double* sumsq_out, void* stream) {
  if (!h || !trace_x || T < 1 || T > (1 << 30) || d < 1 || d > NNB_MAX_DIM || n < 1) return NNB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  NNB_CUDA(h, cudaSetDevice(h->device));
  const size_t nout = (size_t)2 * d + 2;
  int rc = stats_ws(h, d, t_scale, t_shift, nullptr, nout, st);
  if (rc) return rc;
  double* ts = h->d_stats_ws;
  double* out = ts + 3 * d;
  const unsigned int blocks = (unsigned int)((n + kStatThreads - 1) / kStatThreads);
  chain_moves_kernel<<<blocks, kStatThreads, 0, st>>>(trace_x, (int)T, d, n, ts, out);
  if (sum_out || sumsq_out)
    chain_moments_kernel<<<dim3(blocks, d), kStatThreads, 0, st>>>(trace_x, (int)T, d, n, ts, ts + d, out + 2);
  NNB_CUDA(h, cudaGetLastError());
  std::vector<double> host(nout);
  NNB_CUDA(h, cudaMemcpyAsync(host.data(), out, nout * sizeof(double), cudaMemcpyDeviceToHost, st));
  NNB_CUDA(h, cudaStreamSynchronize(st));
  if (moved_out) *moved_out = host[0];
  if (jump_sum_out) *jump_sum_out = host[1];
  for (int i = 0; i < d; ++i) {
    if (sum_out) sum_out[i] = host[2 + i];
    if (sumsq_out) sumsq_out[i] = host[2 + d + i];
  }
  return NNB_OK;
}

extern "C" int nnb_chain_autocorr(nnb_handle* h, const float* trace_x, int64_t T, int d, int64_t n, const double* t_scale,
                                  const double* t_shift, const double* mean, int lag0, int nlags, double* out_host,
                                  void* stream) {
  if (!h || !trace_x || !mean || !out_host || T < 1 || T > (1 << 30) || d < 1 || d > NNB_MAX_DIM || n < 1 || lag0 < 0 ||
      nlags < 1 || nlags > kLagBlock)
    return NNB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  NNB_CUDA(h, cudaSetDevice(h->device));
  const size_t nout = (size_t)kLagBlock * d;
  int rc = stats_ws(h, d, t_scale, t_shift, mean, nout, st);
  if (rc) return rc;
  double* ts = h->d_stats_ws;
  double* out = ts + 3 * d;
  const unsigned int blocks = (unsigned int)((n + kStatThreads - 1) / kStatThreads);
  chain_autocorr_kernel<<<dim3(blocks, d), kStatThreads, 0, st>>>(trace_x, (int)T, d, n, ts, ts + d, ts + 2 * d, lag0, out);
  NNB_CUDA(h, cudaGetLastError());
  NNB_CUDA(h, cudaMemcpyAsync(out_host, out, (size_t)nlags * d * sizeof(double), cudaMemcpyDeviceToHost, st));
  NNB_CUDA(h, cudaStreamSynchronize(st));
  return NNB_OK;
}
