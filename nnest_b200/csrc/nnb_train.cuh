// nnb_train.cuh -- fused flow fitting: one launch = one epoch of Trainer._train + Trainer._validate
// (reference nnest/trainer.py:384-418): for every mini-batch  data = x[perm] + jitter * N(0, I),
// loss = -mean(log p(data)) (networks.py:71-76: N(0, I) log-density of the forward flow + log-det),
// backward pass, Adam step (torch.optim.Adam semantics incl. L2 weight decay); then the validation loss.
//
// Design (sm_100a, FP32):
//   * persistent kernel, 256 threads per CTA, 128 of them own one sample each in the per-sample phases; a mini-batch is
//     spread over ceil(batch/128) CTAs (at most one per SM).  With the reference's default batch_size = 100 that
//     is ONE CTA walking the whole epoch with weights, gradient and sample vectors resident in shared memory --
//     no launches, no host round trips between iterations.
//   * light-weight backward: a coupling block is invertible, so the backward pass walks the blocks in reverse and
//     recovers each block's input as (y - t) * exp(-s) instead of storing it; the hidden activations of the s/t MLPs
//     are kept from the forward pass in per-thread local memory (L1 resident, 2 (L+1) H floats per block).  Per sample
//     only the current vector and its gradient live in shared memory (2 d floats).
//   * weight gradients are batch contractions dW[o][i] = sum_b dpre[b][o] * in[b][i]: every thread stages its
//     sample's `in` / `dpre` vectors as columns of two shared-memory matrices ([feature][sample], rows 16-byte
//     aligned), then each (o, i) pair is owned by one thread, which runs down the sample axis with float4 loads.
//     No atomics inside a CTA.
//   * several CTAs per mini-batch (cooperative launch): every CTA stores its partial gradient in its own global slot;
//     after a grid barrier CTA c sums slice c of all slots IN CTA ORDER (a fixed-order reduce-scatter: the result does
//     not depend on which CTA arrives first, so a seed reproduces bit for bit), a second grid barrier publishes the
//     summed slices, and every CTA applies the identical Adam update to its own shared-memory copy of the weights
//     (optimizer moments: one private global copy per CTA, L2 resident).  Losses are summed the same way: per-CTA
//     partials in a fixed warp order, added up by the host in CTA order.
// Parameter layout: the caller's flat vector is netG.state_dict() order ("natural", as nnb_set_flow); in shared
// memory each net starts at a multiple of 4 floats and the first layer is stored transposed (W1T[i][j]) so that
// every inner loop reads float4 rows.  Adam is element-wise, so only the copy in / copy out permutes.
#pragma once
#include "nnb_kernels.cuh"

namespace nnb {

constexpr int kTrainThreads = 128;   // samples per slice = worker threads (one sample each)
constexpr int kTrainCta = 256;       // threads per CTA
constexpr int kStageStride = 132;   // floats per staged row: 128 samples + 4 (rows 16-byte aligned, bank-conflict free)
enum { kTagTrain = 3 };

struct TrainCtrl {
  double train_loss;     // (nnb_mean_nn_distance accumulator)
  double val_loss;
  unsigned int ticket;   // grid-barrier arrivals (monotonic)
  unsigned int pad;
};

struct TrainParams {
  int d, B, P, netP;            // P = 2 B netP natural floats
  const float* x_train;         // (n_train, d) row-major
  long long n_train;
  const long long* perm;        // visiting order of the epoch (n_train) or NULL = identity
  int batch_size;
  const float* x_valid;         // (n_valid, d) row-major
  long long n_valid;
  const float* noise;           // optional N(0,1) draws (n_train, d) in visiting order (replay); NULL = Philox
  float jitter;
  unsigned int seed_lo, seed_hi, epoch;
  float lr, beta1, beta2, eps, weight_decay;
  long long step0;              // Adam steps taken before this call
  float* params;                // [P] in / out (natural layout)
  float* adam_m;                // [P] in / out
  float* adam_v;                // [P] in / out
  float* mv_priv;               // [grid][2][P] workspace when grid > 1
  float* gpart;                 // [grid][Psm] workspace when grid > 1: every CTA's partial gradient of the mini-batch
  float* gsum;                  // [Psm] workspace when grid > 1: the partials summed in CTA order
  double* loss_part;            // [grid][2]: per-CTA sums of the mini-batch mean losses (trainer.py:396) and of the
                                // validation -log p (trainer.py:414), zeroed by the host
  TrainCtrl* ctrl;
  float* grad_out;              // optional [P]: data gradient of the LAST mini-batch (natural layout)
  int do_train;
  int grad_only;                // data-parallel mode: one mini-batch, gradient only (no Adam, no validation)
  int batch_total;              // grad_only: samples of the whole mini-batch over all ranks (loss / gradient scale)
};

__host__ __device__ inline int train_net_floats(int d, int H, int L) { return H * d + H + L * (H * H + H) + d * H + d; }
__host__ __device__ inline int train_psm(int d, int H, int L, int B) { return 2 * B * round4(train_net_floats(d, H, L)); }
__host__ __device__ inline int train_stage_rows(int d, int H) {
  const int half = (d + 1) / 2;
  const int im = (H > half ? H : half) + 1, om = H > half ? H : half;
  return 2 * (im + om);   // A and D matrices of both nets
}
__host__ inline size_t train_smem_bytes(int d, int H, int L, int B) {
  return (size_t)2 * train_psm(d, H, L, B) * 4 + (size_t)2 * d * kTrainThreads * 4 +
         (size_t)train_stage_rows(d, H) * kStageStride * 4 + 64;
}

// natural index -> shared-memory index
template <int H>
__device__ __forceinline__ int train_perm_index(int p, int d, int netP, int netPp) {
  const int net = p / netP, r = p - net * netP;
  if (r < H * d) {
    const int j = r / d, i = r - j * d;
    return net * netPp + i * H + j;
  }
  return net * netPp + r;
}

template <int ACT>
__device__ __forceinline__ float train_act(float v) { return ACT == 0 ? tanhf(v) : fmaxf(v, 0.0f); }
// derivative of the activation expressed through its OUTPUT
template <int ACT>
__device__ __forceinline__ float train_dact(float h) { return ACT == 0 ? (1.0f - h * h) : (h > 0.0f ? 1.0f : 0.0f); }

// hidden activations h[0..L] of one net for the sample whose vector is the shared-memory column y
template <int H, int L, int ACT>
__device__ __forceinline__ void train_mlp_hidden(const float* __restrict__ w, int d, int nin, int i0,
                                                 const float* __restrict__ y, float (&h)[L + 1][H]) {
  const float* b1 = w + H * d;
  float pre[H];
#pragma unroll
  for (int j = 0; j < H; ++j) pre[j] = b1[j];
  for (int a = 0; a < nin; ++a) {
    const int i = i0 + 2 * a;
    const float v = y[i * kTrainThreads];
    const float4* wr = reinterpret_cast<const float4*>(w + i * H);
#pragma unroll
    for (int j4 = 0; j4 < H / 4; ++j4) {
      const float4 q = wr[j4];
      pre[4 * j4 + 0] = fmaf(q.x, v, pre[4 * j4 + 0]);
      pre[4 * j4 + 1] = fmaf(q.y, v, pre[4 * j4 + 1]);
      pre[4 * j4 + 2] = fmaf(q.z, v, pre[4 * j4 + 2]);
      pre[4 * j4 + 3] = fmaf(q.w, v, pre[4 * j4 + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < H; ++j) h[0][j] = train_act<ACT>(pre[j]);
  const float* p = b1 + H;
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const float* b2 = p + H * H;
#pragma unroll
    for (int j = 0; j < H; ++j) h[l + 1][j] = train_act<ACT>(dot_row<H>(p + j * H, h[l]) + b2[j]);
    p = b2 + H;
  }
}

// acc[j] += row[j] * v  (row: 16-byte aligned shared memory)
template <int H>
__device__ __forceinline__ void axpy_row(const float* __restrict__ row, float v, float (&acc)[H]) {
  const float4* wr = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int j4 = 0; j4 < H / 4; ++j4) {
    const float4 q = wr[j4];
    acc[4 * j4 + 0] = fmaf(q.x, v, acc[4 * j4 + 0]);
    acc[4 * j4 + 1] = fmaf(q.y, v, acc[4 * j4 + 1]);
    acc[4 * j4 + 2] = fmaf(q.z, v, acc[4 * j4 + 2]);
    acc[4 * j4 + 3] = fmaf(q.w, v, acc[4 * j4 + 3]);
  }
}

// G[base + o*so + i*si] += sum_b D[o][b] * A[i][b]  (i < nI);  G[bbase + o*sb] += sum_b D[o][b]   (bias = row nI of A == 1)
__device__ __forceinline__ void train_stage_gemm(float* __restrict__ G, const float* __restrict__ A,
                                                 const float* __restrict__ D, int nO, int nI, int base, int so, int si,
                                                 int bbase, int sb) {
  const int nI1 = nI + 1, total = nO * nI1;
  for (int e = threadIdx.x; e < total; e += kTrainCta) {
    const int o = e / nI1, i = e - o * nI1;
    const float4* a4 = reinterpret_cast<const float4*>(A + i * kStageStride);
    const float4* d4 = reinterpret_cast<const float4*>(D + o * kStageStride);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
    for (int b = 0; b < kTrainThreads / 4; ++b) {
      const float4 a = a4[b], q = d4[b];
      s0 = fmaf(a.x, q.x, s0);
      s1 = fmaf(a.y, q.y, s1);
      s2 = fmaf(a.z, q.z, s2);
      s3 = fmaf(a.w, q.w, s3);
    }
    const int idx = i < nI ? base + o * so + i * si : bbase + o * sb;
    G[idx] += (s0 + s1) + (s2 + s3);
  }
}

__device__ __forceinline__ void train_grid_barrier(TrainCtrl* c, unsigned int& phase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&c->ticket, 1u);
    ++phase;
    const unsigned int target = phase * gridDim.x;
    while (*reinterpret_cast<volatile unsigned int*>(&c->ticket) < target) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
}

// forward flow of the thread's sample, in place on the column y; returns -log p(x).  KEEP: the hidden activations of
// every block are kept for the backward pass in `acts` (per-thread local memory, L1 resident: 2 (L+1) H floats per block
// -- cheaper to reload than to recompute)
template <int H, int L, bool KEEP>
__device__ __forceinline__ float train_forward_nll(const float* __restrict__ W, int d, int B, int netPp, float* y,
                                                   float (*acts)[2][L + 1][H]) {
  float ld = 0.f;
  const int w3 = H * d + H + L * (H * H + H), b3 = w3 + d * H;
  for (int k = 0; k < B; ++k) {
    const int nin = blk_nin(d, k), i0 = blk_i0(k), nout = blk_nout(d, k), o0 = blk_o0(k);
    const float* ws = W + (2 * k) * netPp;
    const float* wt = ws + netPp;
    float hs[L + 1][H], ht[L + 1][H];
    train_mlp_hidden<H, L, 0>(ws, d, nin, i0, y, hs);
    train_mlp_hidden<H, L, 1>(wt, d, nin, i0, y, ht);
    if (KEEP) {
#pragma unroll
      for (int l = 0; l <= L; ++l)
#pragma unroll
        for (int j = 0; j < H; ++j) {
          acts[k][0][l][j] = hs[l][j];
          acts[k][1][l][j] = ht[l][j];
        }
    }
    for (int o = 0; o < nout; ++o) {
      const int i = o0 + 2 * o;
      const float s = dot_row<H>(ws + w3 + i * H, hs[L]) + ws[b3 + i];
      const float t = dot_row<H>(wt + w3 + i * H, ht[L]) + wt[b3 + i];
      y[i * kTrainThreads] = fmaf(y[i * kTrainThreads], expf(s), t);   // networks.py:296-297
      ld += s;
    }
  }
  float q = 0.f;
  for (int i = 0; i < d; ++i) q = fmaf(y[i * kTrainThreads], y[i * kTrainThreads], q);
  return 0.5f * q + 0.9189385332046727f * (float)d - ld;
}

template <int H, int L>
__global__ void __launch_bounds__(kTrainCta, 1) train_epoch_kernel(TrainParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = p.d, B = p.B, netP = p.netP, netPp = round4(netP), Psm = 2 * B * netPp;
  float* W = reinterpret_cast<float*>(smem_raw);
  float* G = W + Psm;
  float* y_all = G + Psm;
  float* gy_all = y_all + d * kTrainThreads;
  float* stage = gy_all + d * kTrainThreads;
  const int half = (d + 1) / 2;
  const int IM = (H > half ? H : half) + 1, OM = H > half ? H : half;
  float* As = stage;
  float* At = As + IM * kStageStride;
  float* Ds = At + IM * kStageStride;
  float* Dt = Ds + OM * kStageStride;
  double* red_d = reinterpret_cast<double*>(Dt + OM * kStageStride);   // 8 doubles behind the staging matrices
  float* red_s = reinterpret_cast<float*>(red_d);
  const int tid = threadIdx.x;
  // warps 0-3 own one sample each ("workers"); warps 4-7 only join the CTA-wide phases (weight-gradient contractions,
  // Adam, copies), which are latency bound with four warps
  const bool worker = tid < kTrainThreads;
  float* y = y_all + (worker ? tid : 0);
  float* gy = gy_all + (worker ? tid : 0);
  const int w2off = H * d + H, w3off = w2off + L * (H * H + H), b3off = w3off + d * H, b1off = H * d;

  for (int q = tid; q < Psm; q += kTrainCta) W[q] = 0.f;
  __syncthreads();
  for (int q = tid; q < p.P; q += kTrainCta) W[train_perm_index<H>(q, d, netP, netPp)] = p.params[q];
  const bool multi = gridDim.x > 1;
  float* m_ptr = multi ? p.mv_priv + (size_t)blockIdx.x * 2 * p.P : p.adam_m;
  float* v_ptr = multi ? m_ptr + p.P : p.adam_v;
  if (multi && p.do_train && !p.grad_only) {
    for (int q = tid; q < p.P; q += kTrainCta) {
      m_ptr[q] = p.adam_m[q];
      v_ptr[q] = p.adam_v[q];
    }
  }
  __syncthreads();

  unsigned int phase = 0;
  float acts[NNB_MAX_BLOCKS][2][L + 1][H];   // local memory: hidden activations of the forward pass, per block
  const long long bs = p.batch_size;
  const long long nmb = p.do_train ? (p.n_train + bs - 1) / bs : 0;
  for (long long mb = 0; mb < nmb; ++mb) {
    const long long cnt = (p.n_train - mb * bs) < bs ? (p.n_train - mb * bs) : bs;   // DataLoader keeps the short tail
    const float inv_bs = 1.0f / (float)(p.grad_only && p.batch_total > 0 ? (long long)p.batch_total : cnt);
    for (int q = tid; q < Psm; q += kTrainCta) G[q] = 0.f;
    float loss_t = 0.f;
    for (long long s0 = (long long)blockIdx.x * kTrainThreads; s0 < cnt; s0 += (long long)gridDim.x * kTrainThreads) {
      const long long s = s0 + tid;
      const bool valid = worker && s < cnt;
      const float vscale = valid ? 1.0f : 0.0f;
      if (worker) {
      // ---- load the sample: x[perm[pos]] + jitter * N(0, I)    (trainer.py:390) ---------------------------------
      if (valid) {
        const long long pos = mb * bs + s;
        const long long row = p.perm ? p.perm[pos] : pos;
        const float* xr = p.x_train + row * d;
        if (p.jitter != 0.0f) {
          if (p.noise) {
            const float* nr = p.noise + pos * d;
            for (int i = 0; i < d; ++i) y[i * kTrainThreads] = fmaf(p.jitter, nr[i], xr[i]);
          } else {
            for (int j = 0; j < (d + 3) / 4; ++j) {
              float nrm[4];
              philox_normals4(j, p.epoch, (unsigned int)pos, kTagTrain, p.seed_lo, p.seed_hi, nrm);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (4 * j + q < d) y[(4 * j + q) * kTrainThreads] = fmaf(p.jitter, nrm[q], xr[4 * j + q]);
            }
          }
        } else {
          for (int i = 0; i < d; ++i) y[i * kTrainThreads] = xr[i];
        }
      } else {
        for (int i = 0; i < d; ++i) y[i * kTrainThreads] = 0.f;
      }
      // ---- forward: z = f(x), loss contribution, dL/dz = z / batch ------------------------------------------------------
      const float nll = train_forward_nll<H, L, true>(W, d, B, netPp, y, acts);
      if (valid) loss_t += nll * inv_bs;
      for (int i = 0; i < d; ++i) gy[i * kTrainThreads] = y[i * kTrainThreads] * inv_bs * vscale;
      }
      // ---- backward, blocks in reverse; y = output of block k on entry, its input on exit -----------------------------
      for (int k = B - 1; k >= 0; --k) {
        const int nin = blk_nin(d, k), i0 = blk_i0(k), nout = blk_nout(d, k), o0 = blk_o0(k);
        const int sbase = (2 * k) * netPp, tbase = sbase + netPp;
        const float* ws = W + sbase;
        const float* wt = W + tbase;
        float hs[L + 1][H], ht[L + 1][H];
        if (worker) {
#pragma unroll
          for (int l = 0; l <= L; ++l)
#pragma unroll
            for (int j = 0; j < H; ++j) {
              hs[l][j] = acts[k][0][l][j];
              ht[l][j] = acts[k][1][l][j];
            }
        }
        float dhs[H], dht[H];
#pragma unroll
        for (int j = 0; j < H; ++j) dhs[j] = dht[j] = 0.f;
        // output layer: z_i = x_i e^{s} + t, log-det += s   ->   ds = g (z_i - t) - 1/batch, dt = g, dx = g e^{s}
        __syncthreads();   // previous GEMM finished with the staging buffers
        if (worker) {
        for (int o = 0; o < nout; ++o) {
          const int i = o0 + 2 * o;
          const float s = dot_row<H>(ws + w3off + i * H, hs[L]) + ws[b3off + i];
          const float t = dot_row<H>(wt + w3off + i * H, ht[L]) + wt[b3off + i];
          const float zv = y[i * kTrainThreads], g = gy[i * kTrainThreads];
          const float ds = (g * (zv - t) - inv_bs) * vscale, dt = g;
          y[i * kTrainThreads] = (zv - t) * expf(-s);
          gy[i * kTrainThreads] = g * expf(s);
          axpy_row<H>(ws + w3off + i * H, ds, dhs);
          axpy_row<H>(wt + w3off + i * H, dt, dht);
          Ds[o * kStageStride + tid] = ds;
          Dt[o * kStageStride + tid] = dt;
        }
#pragma unroll
        for (int j = 0; j < H; ++j) {
          As[j * kStageStride + tid] = hs[L][j];
          At[j * kStageStride + tid] = ht[L][j];
        }
        As[H * kStageStride + tid] = 1.0f;
        At[H * kStageStride + tid] = 1.0f;
        }
        __syncthreads();
        train_stage_gemm(G, As, Ds, nout, H, sbase + w3off + o0 * H, 2 * H, 1, sbase + b3off + o0, 2);
        train_stage_gemm(G, At, Dt, nout, H, tbase + w3off + o0 * H, 2 * H, 1, tbase + b3off + o0, 2);
        // hidden layers, last to first
#pragma unroll
        for (int l = L - 1; l >= 0; --l) {
          float dps[H], dpt[H];
          if (worker) {
#pragma unroll
            for (int j = 0; j < H; ++j) {
              dps[j] = dhs[j] * train_dact<0>(hs[l + 1][j]);
              dpt[j] = dht[j] * train_dact<1>(ht[l + 1][j]);
              dhs[j] = dht[j] = 0.f;
            }
          }
          const int wl = w2off + l * (H * H + H);
          __syncthreads();
          if (worker) {
#pragma unroll
          for (int j = 0; j < H; ++j) {
            axpy_row<H>(ws + wl + j * H, dps[j], dhs);
            axpy_row<H>(wt + wl + j * H, dpt[j], dht);
            Ds[j * kStageStride + tid] = dps[j];
            Dt[j * kStageStride + tid] = dpt[j];
            As[j * kStageStride + tid] = hs[l][j];
            At[j * kStageStride + tid] = ht[l][j];
          }
          As[H * kStageStride + tid] = 1.0f;
          At[H * kStageStride + tid] = 1.0f;
          }
          __syncthreads();
          train_stage_gemm(G, As, Ds, H, H, sbase + wl, H, 1, sbase + wl + H * H, 1);
          train_stage_gemm(G, At, Dt, H, H, tbase + wl, H, 1, tbase + wl + H * H, 1);
        }
        // input layer
        {
          float dps[H], dpt[H];
          if (worker) {
#pragma unroll
            for (int j = 0; j < H; ++j) {
              dps[j] = dhs[j] * train_dact<0>(hs[0][j]);
              dpt[j] = dht[j] * train_dact<1>(ht[0][j]);
            }
          }
          __syncthreads();
          if (worker) {
#pragma unroll
          for (int j = 0; j < H; ++j) {
            Ds[j * kStageStride + tid] = dps[j];
            Dt[j * kStageStride + tid] = dpt[j];
          }
          for (int a = 0; a < nin; ++a) {
            const int i = i0 + 2 * a;
            const float xv = y[i * kTrainThreads];
            As[a * kStageStride + tid] = xv;
            gy[i * kTrainThreads] += dot_row<H>(ws + i * H, dps) + dot_row<H>(wt + i * H, dpt);
          }
          As[nin * kStageStride + tid] = 1.0f;
          }
          __syncthreads();
          // both nets share the input: A = As for both
          train_stage_gemm(G, As, Ds, H, nin, sbase + i0 * H, 1, 2 * H, sbase + b1off, 1);
          train_stage_gemm(G, As, Dt, H, nin, tbase + i0 * H, 1, 2 * H, tbase + b1off, 1);
        }
      }
      __syncthreads();
    }
    // ---- loss of the mini-batch (trainer.py:396) ----------------------------------------------------------------------------
    {
      float v = loss_t;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0) red_s[tid >> 5] = v;
      __syncthreads();
      if (tid == 0) {   // fixed warp order; the slot belongs to this CTA alone
        double t = 0.0;
        for (int w = 0; w < kTrainThreads / 32; ++w) t += (double)red_s[w];
        p.loss_part[2 * blockIdx.x] += t;
      }
    }
    __syncthreads();
    // ---- several CTAs per mini-batch: sum the partial gradients through global memory --------------------------------
    if (multi) {
      float* gp = p.gpart + (size_t)blockIdx.x * Psm;
      for (int q = tid; q < Psm; q += kTrainCta) gp[q] = G[q];
      train_grid_barrier(p.ctrl, phase);
      // fixed-order reduce-scatter: this CTA owns one slice of the parameters and adds the partials in CTA order
      const int chunk = (Psm + gridDim.x - 1) / gridDim.x;
      const int q_hi = min(Psm, (int)(blockIdx.x + 1) * chunk);
      for (int q = blockIdx.x * chunk + tid; q < q_hi; q += kTrainCta) {
        float acc = 0.f;
#pragma unroll 8
        for (unsigned int c = 0; c < gridDim.x; ++c) acc += __ldcg(p.gpart + (size_t)c * Psm + q);
        p.gsum[q] = acc;
      }
      // the slots are rewritten only after the next mini-batch's compute, i.e. after every CTA has passed this barrier
      train_grid_barrier(p.ctrl, phase);
      for (int q = tid; q < Psm; q += kTrainCta) G[q] = __ldcg(p.gsum + q);
      __syncthreads();
    }
    if (p.grad_out && mb == nmb - 1 && blockIdx.x == 0)
      for (int q = tid; q < p.P; q += kTrainCta) p.grad_out[q] = G[train_perm_index<H>(q, d, netP, netPp)];
    // ---- Adam (torch.optim.Adam, weight decay added to the gradient), identical in every CTA -----------------------------
    if (!p.grad_only) {
      const double step = (double)(p.step0 + mb + 1);
      const float bc1 = (float)(1.0 - pow((double)p.beta1, step));
      const float bc2s = (float)sqrt(1.0 - pow((double)p.beta2, step));
      const float step_size = p.lr / bc1;
      // walk the parameters in shared-memory order (no integer division to map indices: H is a power of two) and
      // fetch the moments of four parameters before touching any of them, so that their L2 latencies overlap
      float* __restrict__ mp = m_ptr;
      float* __restrict__ vp = v_ptr;
      for (int net = 0; net < 2 * B; ++net) {
        const int qbase = net * netP, wbase = net * netPp;
        for (int r0 = tid; r0 < netP; r0 += 4 * kTrainCta) {
          int qi[4];
          float mv[4], vv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = r0 + u * kTrainCta;
            // first layer: shared memory holds W1T[i][j] (r = i H + j), the caller's vector W1[j][i] (j d + i)
            qi[u] = r < netP ? qbase + (r < H * d ? (r % H) * d + r / H : r) : -1;
            mv[u] = qi[u] >= 0 ? mp[qi[u]] : 0.f;
            vv[u] = qi[u] >= 0 ? vp[qi[u]] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (qi[u] < 0) continue;
            const int w = wbase + r0 + u * kTrainCta;
            const float wv = W[w];
            const float g = fmaf(p.weight_decay, wv, G[w]);
            const float m = mv[u] + (g - mv[u]) * (1.0f - p.beta1);     // exp_avg.lerp_(grad, 1 - beta1)
            const float v = fmaf(g * g, 1.0f - p.beta2, vv[u] * p.beta2);   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
            mp[qi[u]] = m;
            vp[qi[u]] = v;
            const float denom = sqrtf(v) / bc2s + p.eps;
            W[w] = wv - step_size * (m / denom);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- validation: -log p of the held-out samples with the epoch's final weights (trainer.py:405-418) -----------------
  {
    float loss_t = 0.f;
    for (long long s0 = (long long)blockIdx.x * kTrainThreads; s0 < p.n_valid; s0 += (long long)gridDim.x * kTrainThreads) {
      const long long s = s0 + tid;
      const bool valid = worker && s < p.n_valid;
      if (worker) {
        const float* xr = p.x_valid + (valid ? s : 0) * d;
        for (int i = 0; i < d; ++i) y[i * kTrainThreads] = valid ? xr[i] : 0.f;
        const float nll = train_forward_nll<H, L, false>(W, d, B, netPp, y, nullptr);
        if (valid) loss_t += nll;
      }
    }
    double v = (double)loss_t;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) red_d[tid >> 5] = v;
    __syncthreads();
    if (tid == 0 && p.n_valid > 0) {
      double t = 0.0;
      for (int w = 0; w < kTrainThreads / 32; ++w) t += red_d[w];
      p.loss_part[2 * blockIdx.x + 1] = t;
    }
  }
  // ---- publish ----------------------------------------------------------------------------------------------------------------
  if (blockIdx.x == 0 && p.do_train && !p.grad_only) {
    for (int q = tid; q < p.P; q += kTrainCta) {
      p.params[q] = W[train_perm_index<H>(q, d, netP, netPp)];
      if (multi) {
        p.adam_m[q] = m_ptr[q];
        p.adam_v[q] = v_ptr[q];
      }
    }
  }
}

// ---- mean nearest-neighbour distance, float32 prefilter + exact float64 refinement ------------------------------------------
// The float64 brute force below costs 25 ms at 65 536 x 30 and runs before each of the 272 flow fits of a config-4 run.
// Here every candidate is first rated in float32 (direct differences, FP32 FMA pipe); only candidates whose float32 distance
// is within the float32 error bound of the running minimum are re-evaluated exactly in float64 from the original rows.
// The result is EXACT (the same minimum the float64 brute force finds):
//   |sqrt(d32) - sqrt(d64)| <= eps := eps_abs + eps_rel * sqrt(d64)   for every pair, with eps_abs = 2.5e-7 sqrt(d) max|x|
//   (float32 rounding of the coordinates and of their differences) and eps_rel = 2e-6 (d + 2 roundings of 2^-24), so when the
//   true nearest neighbour c* comes by, sqrt(d32(c*)) <= sqrt(d64(c*)) + eps <= sqrt(d64(c_best)) + eps <= sqrt(best32) + 2 eps,
//   where c_best is the candidate that set the running float32 minimum: c* passes the test and is evaluated exactly.
// nn_prepare_kernel: float32 copy of the rows, zero padded to DP floats, and max |x| (as the bits of a non-negative float).
__global__ void nn_prepare_kernel(const double* __restrict__ x, long long n, int d, int DP, float* __restrict__ xf,
                                  unsigned int* __restrict__ maxabs_bits) {
  const long long total = n * DP;
  float m = 0.f;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / DP;
    const int i = (int)(e - r * DP);
    const float v = i < d ? (float)x[r * d + i] : 0.f;
    xf[e] = v;
    m = fmaxf(m, fabsf(v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(maxabs_bits, __float_as_uint(m));
}

// grid = (query blocks, candidate splits): block (bx, by) rates the candidates of split by for the 128 queries of bx and
// writes the exact minimum squared distance it found to best_out[by][row]; nn_reduce_kernel takes the minimum over the
// splits (one thread per query does not fill a B200: 65 536 queries are 14 warps per SM).
template <int DP>
__global__ void __launch_bounds__(128) nn_min_dist_f32_kernel(const double* __restrict__ x, const float* __restrict__ xf,
                                                              long long n, int d, const unsigned int* __restrict__ maxabs_bits,
                                                              double* __restrict__ best_out) {
  __shared__ __align__(16) float c[128 * DP];
  const int tid = threadIdx.x;
  const long long row = (long long)blockIdx.x * 128 + tid;
  const bool valid = row < n;
  float q[DP];
#pragma unroll
  for (int i = 0; i < DP; ++i) q[i] = valid ? xf[row * DP + i] : 0.f;
  const float eps_abs = 2.5e-7f * sqrtf((float)d) * __uint_as_float(*maxabs_bits) + 1e-30f;
  float best32 = INFINITY, thr = INFINITY;
  double best64 = INFINITY;
  const long long tiles = (n + 127) / 128, per = (tiles + gridDim.y - 1) / gridDim.y;
  const long long c_lo = (long long)blockIdx.y * per * 128;
  const long long c_hi = c_lo + per * 128 < n ? c_lo + per * 128 : n;
  for (long long c0 = c_lo; c0 < c_hi; c0 += 128) {
    const int m = (int)((c_hi - c0) < 128 ? (c_hi - c0) : 128);
    __syncthreads();
    {
      const float4* src = reinterpret_cast<const float4*>(xf + c0 * DP);
      float4* dst = reinterpret_cast<float4*>(c);
      for (int e = tid; e < m * (DP / 4); e += 128) dst[e] = src[e];
    }
    __syncthreads();
    for (int j = 0; j < m; ++j) {
      const float4* cj = reinterpret_cast<const float4*>(c + j * DP);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int i = 0; i < DP / 4; ++i) {
        const float4 v = cj[i];
        const float t0 = q[4 * i] - v.x, t1 = q[4 * i + 1] - v.y, t2 = q[4 * i + 2] - v.z, t3 = q[4 * i + 3] - v.w;
        s0 = fmaf(t0, t0, s0);
        s1 = fmaf(t1, t1, s1);
        s2 = fmaf(t2, t2, s2);
        s3 = fmaf(t3, t3, s3);
      }
      const float s = (s0 + s1) + (s2 + s3);
      if (s <= thr && valid && c0 + j != row) {   // rare: within the float32 error bound of the running minimum
        const double* a = x + row * d;
        const double* b = x + (c0 + j) * d;
        double e0 = 0.0, e1 = 0.0;
        int i = 0;
        for (; i + 1 < d; i += 2) {      // same pairing as the float64 kernel (even / odd coordinates)
          const double u0 = a[i] - b[i], u1 = a[i + 1] - b[i + 1];
          e0 = fma(u0, u0, e0);
          e1 = fma(u1, u1, e1);
        }
        if (i < d) {
          const double u0 = a[i] - b[i];
          e0 = fma(u0, u0, e0);
        }
        const double ex = e0 + e1;
        if (ex < best64) best64 = ex;
        if (s < best32) {
          best32 = s;
          const float r = sqrtf(best32) * (1.0f + 4e-6f) + 2.0f * eps_abs;
          thr = r * r * (1.0f + 1e-6f);
        }
      }
    }
  }
  if (valid) best_out[(long long)blockIdx.y * n + row] = best64;
}

// minimum over the candidate splits, square root, per-block partial sums in a fixed order
__global__ void __launch_bounds__(128) nn_reduce_kernel(const double* __restrict__ best, long long n, int splits,
                                                        double* __restrict__ sum_out) {
  const int tid = threadIdx.x;
  const long long row = (long long)blockIdx.x * 128 + tid;
  double v = 0.0;
  if (row < n && n > 1) {
    double b = best[row];
    for (int s = 1; s < splits; ++s) b = fmin(b, best[(long long)s * n + row]);
    v = sqrt(b);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ double wsum[4];
  if ((tid & 31) == 0) wsum[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) sum_out[blockIdx.x] = ((wsum[0] + wsum[1]) + wsum[2]) + wsum[3];
}

// ---- mean nearest-neighbour distance (training jitter, trainer.py:147-150), float64 brute force (d > 32) ----------------
// One query row per thread, candidates streamed through shared memory in tiles of 128 rows; float64 like the reference's
// cKDTree on float64 samples.  DREG > 0: the query lives in registers (d <= DREG, zero padded) and a candidate coordinate
// is one broadcast shared-memory load per two dimensions; DREG == 0: any d, query coordinates as shared-memory columns.
template <int DREG>
__global__ void __launch_bounds__(128, 2) nn_min_dist_kernel(const double* __restrict__ x, long long n, int d,
                                                             double* __restrict__ sum_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const long long row = (long long)blockIdx.x * 128 + tid;
  const bool valid = row < n;
  double best = INFINITY;
  if (DREG > 0) {
    double* c = reinterpret_cast<double*>(smem_raw);   // [128][DREG]
    double q[DREG > 0 ? DREG : 1];
#pragma unroll
    for (int i = 0; i < DREG; ++i) q[i] = (valid && i < d) ? x[row * d + i] : 0.0;
    for (long long c0 = 0; c0 < n; c0 += 128) {
      const int m = (int)((n - c0) < 128 ? (n - c0) : 128);
      __syncthreads();
      for (int e = tid; e < m * DREG; e += 128) {
        const int j = e / DREG, i = e - j * DREG;
        c[e] = i < d ? x[(c0 + j) * d + i] : 0.0;
      }
      __syncthreads();
      for (int j = 0; j < m; ++j) {
        const double2* cj = reinterpret_cast<const double2*>(c + j * DREG);
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < DREG / 2; ++i) {
          const double2 v = cj[i];
          const double t0 = q[2 * i] - v.x, t1 = q[2 * i + 1] - v.y;
          s0 = fma(t0, t0, s0);
          s1 = fma(t1, t1, s1);
        }
        const double s = s0 + s1;
        if (c0 + j != row && s < best) best = s;
      }
    }
  } else {
    double* q = reinterpret_cast<double*>(smem_raw);   // [d][128]
    double* c = q + (size_t)d * 128;                   // [128][d]
    for (int i = 0; i < d; ++i) q[i * 128 + tid] = valid ? x[row * d + i] : 0.0;
    for (long long c0 = 0; c0 < n; c0 += 128) {
      const int m = (int)((n - c0) < 128 ? (n - c0) : 128);
      __syncthreads();
      for (int e = tid; e < m * d; e += 128) c[e] = x[c0 * d + e];
      __syncthreads();
      for (int j = 0; j < m; ++j) {
        const double* cj = c + j * d;
        double s = 0.0;
        for (int i = 0; i < d; ++i) {
          const double t = q[i * 128 + tid] - cj[i];
          s = fma(t, t, s);
        }
        if (c0 + j != row && s < best) best = s;
      }
    }
  }
  // per-block partial in a fixed order (the host adds the blocks in order): the jitter of a seeded run is reproducible
  double v = valid && n > 1 ? sqrt(best) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ double wsum[4];
  if ((tid & 31) == 0) wsum[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) sum_out[blockIdx.x] = ((wsum[0] + wsum[1]) + wsum[2]) + wsum[3];
}

}  // namespace nnb
