// nnb_spline.cu -- the reference's DEFAULT flow, flow='spline' (SingleSpeedSpline: [ActNorm, Invertible1x1Conv, NSF_CL] x
// num_blocks, nnest/networks.py:393-705), on the device: batched flow maps, chain start and the fused MCMC step.
//
// Same contracts as flow_kernel / mcmc_init_kernel / mcmc_kernel (nnb_kernels.cuh; reference nnest/trainer.py:247-269,
// nnest/sampler.py:262-444) with the per-sample arithmetic of nnb_spline.cuh (pinned to reference goldens on the CPU).
// One chain per thread, 128 chains per CTA, chain vectors as shared-memory columns.  The parameters (46 k floats at
// d = 30, hidden 16: they do not fit next to the chain vectors) stay in global memory: every lane of a warp reads the
// same weight at the same time, i.e. one L1-resident broadcast line per load.
#include <cuda_runtime.h>

#include <string>

#include "nnb_host.h"
#include "nnb_spline.cuh"

using namespace nnb;

namespace {

struct SplineSmem {
  TargetSmem tg;
  float* y;
  float* zp;
  float* tmp;
  unsigned int* red;
};

__host__ __device__ inline size_t spline_smem_bytes(int tdoubles, int d) {
  return (size_t)tdoubles * 8 + (size_t)3 * d * kBlockThreads * 4 + 32 * 4;
}

__device__ __forceinline__ SplineSmem spline_carve(unsigned char* base, const TargetDesc& td, const double* tgt_g, int d) {
  SplineSmem v;
  double* td_s = reinterpret_cast<double*>(base);
  const int nd = tgt_g ? target_doubles(td.d, td.n_params) : 0;
  for (int i = threadIdx.x; i < nd; i += blockDim.x) td_s[i] = tgt_g[i];
  target_bind(v.tg, td, td_s);
  v.y = reinterpret_cast<float*>(td_s + nd);
  v.zp = v.y + (size_t)d * kBlockThreads;
  v.tmp = v.zp + (size_t)d * kBlockThreads;
  v.red = reinterpret_cast<unsigned int*>(v.tmp + (size_t)d * kBlockThreads);
  return v;
}

template <bool INVERSE>
__global__ void __launch_bounds__(kBlockThreads)
spline_flow_kernel(spline::Shape sh, const float* __restrict__ packed, const float* __restrict__ in, long long in_rs,
                   long long in_cs, float* __restrict__ out, long long out_rs, long long out_cs,
                   float* __restrict__ logdet, int* __restrict__ empty_out, long long n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TargetDesc td{};
  SplineSmem sv = spline_carve(smem_raw, td, nullptr, sh.d);
  float* y = sv.y + threadIdx.x;
  float* tmp = sv.tmp + threadIdx.x;
  const long long ntiles = (n + kBlockThreads - 1) / kBlockThreads;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long c = tile * kBlockThreads + threadIdx.x;
    if (c >= n) continue;
    for (int i = 0; i < sh.d; ++i) y[i * kBlockThreads] = in[c * in_rs + i * in_cs];
    int empty = 0;
    const float ld = INVERSE ? spline::flow_inverse(sh, packed, y, tmp, kBlockThreads, &empty)
                             : spline::flow_forward(sh, packed, y, tmp, kBlockThreads, &empty);
    for (int i = 0; i < sh.d; ++i) out[c * out_rs + i * out_cs] = y[i * kBlockThreads];
    if (logdet) logdet[c] = ld;
    if (empty_out) empty_out[c] = empty;
  }
}

__global__ void __launch_bounds__(kBlockThreads)
spline_init_kernel(spline::Shape sh, const float* __restrict__ packed, TargetDesc td, const double* __restrict__ tgt_g,
                   InitParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SplineSmem sv = spline_carve(smem_raw, td, tgt_g, sh.d);
  __syncthreads();
  const int d = sh.d;
  const long long n = p.n;
  const long long c = (long long)blockIdx.x * kBlockThreads + threadIdx.x;
  const bool active = c < n;
  float* y = sv.y + threadIdx.x;
  float* tmp = sv.tmp + threadIdx.x;
  unsigned int ncall = 0, nbad = 0;
  if (active) {
    if (p.init_u) {  // sampler.py:264-266: z = forward(u); x = inverse(z) "due to numerical precision"
      for (int i = 0; i < d; ++i) y[i * kBlockThreads] = p.init_u[(long long)i * n + c];
      spline::flow_forward(sh, packed, y, tmp, kBlockThreads);
    } else if (p.init_z) {
      for (int i = 0; i < d; ++i) y[i * kBlockThreads] = p.init_z[(long long)i * n + c];
    } else {  // sampler.py:276: z ~ N(0, I)
      const unsigned int chain = (unsigned int)(p.chain_offset + (unsigned long long)c);
      for (int j = 0; j < (d + 3) / 4; ++j) {
        float nrm[4];
        philox_normals4(j, p.start_try, chain, kTagInit, p.seed_lo, p.seed_hi, nrm);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (4 * j + q < d) y[(4 * j + q) * kBlockThreads] = nrm[q];
      }
    }
    for (int i = 0; i < d; ++i) p.z[(long long)i * n + c] = y[i * kBlockThreads];
    const float ld = spline::flow_inverse(sh, packed, y, tmp, kBlockThreads);
    for (int i = 0; i < d; ++i) p.x[(long long)i * n + c] = y[i * kBlockThreads];
    p.logdet[c] = ld;
    SmemRow row{y};
    double l;
    if (p.init_logl) {
      l = p.init_logl[c];
    } else {
      l = loglike_any(sv.tg, row, false);
      ncall = 1;
    }
    p.logl[c] = l;
    p.logp[c] = prior_any(sv.tg, row, false);
    nbad = !(l > -1e30);
  }
  unsigned int tc = block_sum_u32(ncall, sv.red);
  unsigned int tb = block_sum_u32(nbad, sv.red);
  if (threadIdx.x == 0) {
    if (tc) atomicAdd(&p.ctrl->ncall, (unsigned long long)tc);
    if (tb) atomicAdd(&p.ctrl->nbad, (unsigned long long)tb);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kBlockThreads)
spline_mcmc_kernel(spline::Shape sh, const float* __restrict__ packed, TargetDesc td, const double* __restrict__ tgt_g,
                   McmcParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SplineSmem sv = spline_carve(smem_raw, td, tgt_g, sh.d);
  __syncthreads();
  const int d = sh.d;
  const long long n = p.n;
  const long long c = (long long)blockIdx.x * kBlockThreads + threadIdx.x;
  const bool active = c < n;
  float* y = sv.y + threadIdx.x;
  float* zp = sv.zp + threadIdx.x;
  float* tmp = sv.tmp + threadIdx.x;
  const unsigned int chain = (unsigned int)(p.chain_offset + (unsigned long long)c);
  unsigned int acc_total = 0, ncall_total = 0;
  float ld_cur = 0.f;
  double logl_cur = 0.0, logp_cur = 0.0;
  if (active) {
    ld_cur = p.logdet[c];
    logl_cur = p.logl[c];
    logp_cur = p.logp[c];
  }
  for (int s = p.s0 + 1; s <= p.s0 + p.nsteps; ++s) {
    const float scale_f = (float)(*reinterpret_cast<volatile double*>(&p.ctrl->scale));
    const unsigned int step_abs = p.step_offset + (unsigned int)s;
    bool accept = false;
    unsigned int ncall = 0;
    if (active) {
      // ---- proposal z' = z + scale * N(0, I) -------------------------------------------- sampler.py:310-316
      if (p.replay_normals) {
        const float* nr = p.replay_normals + ((long long)(s - 1) * n + c) * d;
        for (int i = 0; i < d; ++i) {
          const float v = __fadd_rn(p.z[(long long)i * n + c], __fmul_rn(nr[i], scale_f));
          y[i * kBlockThreads] = v;
          zp[i * kBlockThreads] = v;
        }
      } else {
        for (int j = 0; j < (d + 3) / 4; ++j) {
          float nrm[4];
          philox_normals4(j, step_abs, chain, kTagNormal, p.seed_lo, p.seed_hi, nrm);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int i = 4 * j + q;
            if (i < d) {
              const float v = __fadd_rn(p.z[(long long)i * n + c], __fmul_rn(nrm[q], scale_f));
              y[i * kBlockThreads] = v;
              zp[i * kBlockThreads] = v;
              if (p.dump_normals) p.dump_normals[((long long)(s - 1) * n + c) * d + i] = nrm[q];
            }
          }
        }
      }
      // ---- x', log|det J|' = flow.inverse(z') -------------------------------------------- sampler.py:320-324
      int empty = 0;
      const float ld_prop = spline::flow_inverse(sh, packed, y, tmp, kBlockThreads, &empty);
      float u01;
      if (p.replay_uniforms) {
        u01 = p.replay_uniforms[(long long)(s - 1) * n + c];
      } else {
        uint4 r = philox4x32_10(0u, step_abs, chain, kTagUniform, p.seed_lo, p.seed_hi);
        u01 = uniform01(r.x);
        if (p.dump_uniforms) p.dump_uniforms[(long long)(s - 1) * n + c] = u01;
      }
      // a ONE-chain batch whose proposal has a coupling half entirely outside the tail bound: the reference's inverse
      // raises ValueError and the proposal is skipped (sampler.py:320-324).  With more than one chain the reference only
      // raises when no chain at all has a coordinate inside, which does not happen in practice: identity map as coded.
      const bool skip = empty != 0 && n == 1;
      SmemRow row{y};
      double lp = 0.0, logp_prop = 0.0;
      if (skip) {
      } else if (MODE == NNB_MODE_HARD) {
        float lr = __fsub_rn(ld_prop, ld_cur);
        logp_prop = prior_any(sv.tg, row, false);
        if (logp_prop < -1e30) lr = -INFINITY;
        float ratio = expf(lr);
        if (ratio > 1.0f) ratio = 1.0f;
        const bool m1 = u01 < ratio;
        if (m1) {
          lp = loglike_any(sv.tg, row, false);
          ncall = 1;
          accept = isfinite(lp) && (lp > p.loglstar);
        }
      } else {
        lp = loglike_any(sv.tg, row, false);
        ncall = 1;
        logp_prop = prior_any(sv.tg, row, false);
        double lr = (double)__fsub_rn(ld_prop, ld_cur) + (lp - logl_cur) + (logp_prop - logp_cur);
        double ratio = exp(lr);
        if (ratio > 1.0) ratio = 1.0;
        accept = (double)u01 < ratio;
      }
      if (accept) {
        for (int i = 0; i < d; ++i) {
          p.z[(long long)i * n + c] = zp[i * kBlockThreads];
          p.x[(long long)i * n + c] = y[i * kBlockThreads];
        }
        ld_cur = ld_prop;
        logl_cur = lp;
        logp_cur = logp_prop;
      }
      if (p.trace_z) {
        float* tz = p.trace_z + (long long)s * d * n + c;
        float* tx = p.trace_x + (long long)s * d * n + c;
        if (accept) {
          for (int i = 0; i < d; ++i) {
            tz[(long long)i * n] = zp[i * kBlockThreads];
            tx[(long long)i * n] = y[i * kBlockThreads];
          }
        } else {
          for (int i = 0; i < d; ++i) {
            tz[(long long)i * n] = p.z[(long long)i * n + c];
            tx[(long long)i * n] = p.x[(long long)i * n + c];
          }
        }
        p.trace_logl[(long long)s * n + c] = logl_cur;
      }
    }
    acc_total += accept ? 1u : 0u;
    ncall_total += ncall;
    if (p.dynamic) {   // one launch per step: the last CTA to finish owns the scale update (sampler.py:418-430)
      unsigned int blk = block_count(accept);
      if (threadIdx.x == 0) {
        if (blk) atomicAdd(&p.ctrl->step_acc, blk);
        __threadfence();
        unsigned int t = atomicAdd(&p.ctrl->ticket, 1u);
        if (t == gridDim.x - 1) {
          __threadfence();
          unsigned int na = atomicExch(&p.ctrl->step_acc, 0u);
          p.ctrl->ticket = 0u;
          int a = p.ctrl->accept, r = p.ctrl->reject;
          if (2ull * na > (unsigned long long)n) a += 1; else r += 1;
          double sc = p.ctrl->scale;
          if (a > r) sc *= exp(1.0 / (1 + a));
          if (a < r) sc /= exp(1.0 / (1 + r));
          p.ctrl->accept = a;
          p.ctrl->reject = r;
          p.ctrl->scale = sc;
        }
      }
    }
  }
  if (active) {
    p.logdet[c] = ld_cur;
    p.logl[c] = logl_cur;
    p.logp[c] = logp_cur;
  }
  unsigned int ta = block_sum_u32(acc_total, sv.red);
  unsigned int tc = block_sum_u32(ncall_total, sv.red);
  if (threadIdx.x == 0) {
    if (ta) atomicAdd(&p.ctrl->naccept, (unsigned long long)ta);
    if (tc) atomicAdd(&p.ctrl->ncall, (unsigned long long)tc);
  }
}

spline::Shape shape_of(const nnb_handle* h) {
  spline::Shape sh;
  sh.d = h->spline_d; sh.H = h->spline_hidden; sh.blocks = h->spline_blocks; sh.K = h->spline_bins;
  sh.bound = h->spline_bound;
  return sh;
}

}  // namespace

extern "C" int nnb_set_flow_spline(nnb_handle* h, int d, int hidden, int num_blocks, int num_bins, double tail_bound,
                                   const float* packed, size_t n_floats) {
  if (!h) return NNB_ERR_ARG;
  if (d < 2 || d > NNB_MAX_DIM) return nnb_fail(h, NNB_ERR_ARG, "x_dim must be in [2, NNB_MAX_DIM]");
  if (hidden < 1 || hidden > spline::kMaxHidden) return nnb_fail(h, NNB_ERR_UNSUPPORTED, "hidden_dim must be <= 64");
  if (num_bins < 2 || num_bins > spline::kMaxBins) return nnb_fail(h, NNB_ERR_UNSUPPORTED, "num_bins must be in [2, 16]");
  if (num_blocks < 1 || num_blocks > NNB_MAX_BLOCKS) return nnb_fail(h, NNB_ERR_ARG, "num_blocks out of range");
  if (!(tail_bound > 0.0)) return nnb_fail(h, NNB_ERR_ARG, "tail_bound must be positive");
  spline::Shape sh;
  sh.d = d; sh.H = hidden; sh.blocks = num_blocks; sh.K = num_bins; sh.bound = (float)tail_bound;
  if (!packed || n_floats != (size_t)num_blocks * sh.block_floats())
    return nnb_fail(h, NNB_ERR_ARG, "parameter buffer size does not match (d, hidden, num_blocks, num_bins)");
  if (spline_smem_bytes(target_doubles(d, NNB_MAX_LIKE_PARAMS), d) > (size_t)h->max_smem)
    return nnb_fail(h, NNB_ERR_UNSUPPORTED, "x_dim too large for one CTA's shared memory");
  NNB_CUDA(h, cudaSetDevice(h->device));
  NNB_CUDA(h, nnb_reserve(&h->d_weights_spline, &h->weights_spline_cap, n_floats));
  NNB_CUDA(h, cudaMemcpy(h->d_weights_spline, packed, n_floats * sizeof(float), cudaMemcpyHostToDevice));
  h->spline_d = d; h->spline_hidden = hidden; h->spline_blocks = num_blocks; h->spline_bins = num_bins;
  h->spline_bound = (float)tail_bound;
  h->flow_is_spline = true;
  h->has_flow = true;
  h->flow.d = d;            // x_dim checks of the MCMC entry points
  h->tc_ok = false;
  h->warp_ok = false;
  return NNB_OK;
}

int nnb_spline_flow(nnb_handle* h, bool inverse, const float* in, int64_t irs, int64_t ics, float* out, int64_t ors,
                    int64_t ocs, float* logdet, int* empty_out, int64_t n, cudaStream_t st) {
  const spline::Shape sh = shape_of(h);
  const size_t sm = spline_smem_bytes(0, sh.d);
  const int grid = nnb_grid_for(h, n, 4);
  if (inverse) {
    NNB_CUDA(h, nnb_set_smem(spline_flow_kernel<true>, sm));
    spline_flow_kernel<true><<<grid, kBlockThreads, sm, st>>>(sh, h->d_weights_spline, in, irs, ics, out, ors, ocs, logdet,
                                                              empty_out, n);
  } else {
    NNB_CUDA(h, nnb_set_smem(spline_flow_kernel<false>, sm));
    spline_flow_kernel<false><<<grid, kBlockThreads, sm, st>>>(sh, h->d_weights_spline, in, irs, ics, out, ors, ocs, logdet,
                                                               empty_out, n);
  }
  NNB_CUDA(h, cudaGetLastError());
  return NNB_OK;
}

int nnb_spline_init(nnb_handle* h, const InitParams& p, cudaStream_t st) {
  const spline::Shape sh = shape_of(h);
  const size_t sm = spline_smem_bytes(target_doubles(h->tdesc.d, h->tdesc.n_params), sh.d);
  NNB_CUDA(h, nnb_set_smem(spline_init_kernel, sm));
  const int grid = (int)((p.n + kBlockThreads - 1) / kBlockThreads);
  spline_init_kernel<<<grid, kBlockThreads, sm, st>>>(sh, h->d_weights_spline, h->tdesc, h->d_target, p);
  NNB_CUDA(h, cudaGetLastError());
  return NNB_OK;
}

template <int MODE>
static int spline_mcmc_mode(nnb_handle* h, McmcParams p, int steps, cudaStream_t st) {
  const spline::Shape sh = shape_of(h);
  const size_t sm = spline_smem_bytes(target_doubles(h->tdesc.d, h->tdesc.n_params), sh.d);
  NNB_CUDA(h, nnb_set_smem(spline_mcmc_kernel<MODE>, sm));
  const int grid = (int)((p.n + kBlockThreads - 1) / kBlockThreads);
  if (p.dynamic) {
    for (int s = 0; s < steps; ++s) {
      p.s0 = s; p.nsteps = 1;
      spline_mcmc_kernel<MODE><<<grid, kBlockThreads, sm, st>>>(sh, h->d_weights_spline, h->tdesc, h->d_target, p);
    }
  } else {
    p.s0 = 0; p.nsteps = steps;
    spline_mcmc_kernel<MODE><<<grid, kBlockThreads, sm, st>>>(sh, h->d_weights_spline, h->tdesc, h->d_target, p);
  }
  NNB_CUDA(h, cudaGetLastError());
  h->last_launches = p.dynamic ? steps : 1;
  return NNB_OK;
}

int nnb_spline_mcmc(nnb_handle* h, McmcParams p, int steps, cudaStream_t st) {
  return p.mode == NNB_MODE_MH ? spline_mcmc_mode<NNB_MODE_MH>(h, p, steps, st)
                               : spline_mcmc_mode<NNB_MODE_HARD>(h, p, steps, st);
}

extern "C" int nnb_flow_empty_halves(nnb_handle* h, const float* z, int64_t z_rs, int64_t z_cs, int inverse, int* flags,
                                     int64_t n, void* stream) {
  if (!h || !flags || (n > 0 && !z) || n < 0) return NNB_ERR_ARG;
  if (!h->has_flow) return nnb_fail(h, NNB_ERR_STATE, "no flow set");
  NNB_CUDA(h, cudaSetDevice(h->device));
  if (!h->flow_is_spline) {   // affine coupling layers are defined everywhere
    NNB_CUDA(h, cudaMemsetAsync(flags, 0, sizeof(int) * n, (cudaStream_t)stream));
    return NNB_OK;
  }
  if (n == 0) return NNB_OK;
  // scratch output: the flags come from a full evaluation of the map
  float* scratch = nullptr;
  NNB_CUDA(h, cudaMallocAsync(&scratch, sizeof(float) * (size_t)n * h->spline_d, (cudaStream_t)stream));
  int rc = nnb_spline_flow(h, inverse != 0, z, z_rs, z_cs, scratch, h->spline_d, 1, nullptr, flags, n, (cudaStream_t)stream);
  cudaFreeAsync(scratch, (cudaStream_t)stream);
  return rc;
}
