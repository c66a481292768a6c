// nnb_train.cu -- host side of the fused flow-fitting kernel (nnb_train.cuh): nnb_train_epoch, nnb_mean_nn_distance.
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>
#include <vector>

#include "nnb_host.h"
#include "nnb_train.cuh"
#include "nnb_nn_tc.cuh"

using namespace nnb;

namespace {

// control block + per-CTA loss partials ([sm_count][2] doubles) in one allocation, mirrored in pinned host memory
constexpr size_t kLossOff = 32;
size_t train_ctrl_bytes(const nnb_handle* h) { return kLossOff + (size_t)2 * h->sm_count * sizeof(double); }

int ensure_train_ctrl(nnb_handle* h) {
  if (!h->d_train_ctrl) NNB_CUDA(h, cudaMalloc(&h->d_train_ctrl, train_ctrl_bytes(h)));
  if (!h->h_train_ctrl) NNB_CUDA(h, cudaMallocHost(&h->h_train_ctrl, 2 * train_ctrl_bytes(h)));
  for (cudaEvent_t& e : h->train_ev)
    if (!e) NNB_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return NNB_OK;
}

template <int H, int L>
int launch_train(nnb_handle* h, TrainParams& p, int grid, size_t smem, cudaStream_t st) {
  NNB_CUDA(h, nnb_set_smem(train_epoch_kernel<H, L>, smem));
  if (grid > 1) {
    void* args[] = {(void*)&p};
    NNB_CUDA(h, cudaLaunchCooperativeKernel((const void*)train_epoch_kernel<H, L>, dim3(grid), dim3(kTrainCta), args,
                                            smem, st));
  } else {
    train_epoch_kernel<H, L><<<1, kTrainCta, smem, st>>>(p);
    NNB_CUDA(h, cudaGetLastError());
  }
  return NNB_OK;
}

}  // namespace

extern "C" int nnb_train_supported(int x_dim, int hidden_dim, int num_layers, int num_blocks, int max_smem_bytes) {
  if (x_dim < 2 || x_dim > NNB_MAX_DIM || num_blocks < 1 || num_blocks > NNB_MAX_BLOCKS) return 0;
  if (!((hidden_dim == 16 || hidden_dim == 32) && (num_layers == 1 || num_layers == 2))) return 0;
  return train_smem_bytes(x_dim, hidden_dim, num_layers, num_blocks) <= (size_t)max_smem_bytes ? 1 : 0;
}

// The epoch is queued on the stream and its losses travel to a pinned slot; nnb_train_epoch_end collects them.  Up to two
// epochs may be in flight: the caller queues epoch e + 1 BEFORE it reads the losses of epoch e whenever the outcome of
// epoch e cannot end the fit, so the device never waits for the host between epochs (include/nnb.h).
extern "C" int nnb_train_epoch_begin(nnb_handle* h, const nnb_train_args* a, void* stream) {
  if (!h || !a) return NNB_ERR_ARG;
  if (h->train_begun - h->train_ended >= 2)
    return nnb_fail(h, NNB_ERR_STATE, "nnb_train_epoch_begin: two epochs are already in flight (call nnb_train_epoch_end)");
  cudaStream_t st = (cudaStream_t)stream;
  const int d = a->x_dim, H = a->hidden_dim, L = a->num_layers, B = a->num_blocks;
  if (!nnb_train_supported(d, H, L, B, h->max_smem))
    return nnb_fail(h, NNB_ERR_UNSUPPORTED,
                    "nnb_train_epoch supports hidden_dim in {16, 32}, num_layers in {1, 2}, scale == '' and flows that "
                    "fit one CTA's shared memory");
  const int netP = train_net_floats(d, H, L);
  const int P = 2 * B * netP;
  if (a->n_params != (size_t)P) return nnb_fail(h, NNB_ERR_ARG, "n_params does not match (x_dim, hidden_dim, num_layers, num_blocks)");
  if (!a->params) return nnb_fail(h, NNB_ERR_ARG, "params is NULL");
  if (a->do_train) {
    if (!a->grad_only && (!a->adam_m || !a->adam_v)) return nnb_fail(h, NNB_ERR_ARG, "adam_m / adam_v are NULL");
    if (a->n_train < 0 || (a->n_train > 0 && !a->x_train)) return nnb_fail(h, NNB_ERR_ARG, "x_train");
    if (a->batch_size < 1) return nnb_fail(h, NNB_ERR_ARG, "batch_size must be >= 1");
    if (a->n_train >= (1ll << 32)) return nnb_fail(h, NNB_ERR_ARG, "n_train must be < 2^32");
  }
  if (a->n_valid < 0 || (a->n_valid > 0 && !a->x_valid)) return nnb_fail(h, NNB_ERR_ARG, "x_valid");
  NNB_CUDA(h, cudaSetDevice(h->device));

  const bool grad_only = a->do_train && a->grad_only;
  if (grad_only && (!a->grad_out || a->n_train <= 0)) return nnb_fail(h, NNB_ERR_ARG, "grad_only needs grad_out and n_train > 0");
  long long work = a->do_train ? (grad_only ? (long long)a->n_train : (long long)a->batch_size) : (long long)a->n_valid;
  if (a->do_train && a->n_train < work) work = a->n_train;
  long long g = (work + kTrainThreads - 1) / kTrainThreads;
  if (g < 1) g = 1;
  if (g > h->sm_count) g = h->sm_count;
  if (g > 1 && !h->coop_supported) g = 1;
  const int grid = (int)g;
  const int Psm = train_psm(d, H, L, B);

  { int rc0 = ensure_train_ctrl(h); if (rc0) return rc0; }
  NNB_CUDA(h, cudaMemsetAsync(h->d_train_ctrl, 0, train_ctrl_bytes(h), st));
  if (grid > 1) {
    // [grid][Psm] partial gradients | [Psm] their fixed-order sum | [grid][2][P] private Adam moments
    const size_t need = (size_t)(grid + 1) * Psm + (size_t)grid * 2 * P;
    if (h->train_ws_floats < need) {
      if (h->d_train_ws) cudaFree(h->d_train_ws);
      h->d_train_ws = nullptr;
      NNB_CUDA(h, cudaMalloc(&h->d_train_ws, need * sizeof(float)));
      h->train_ws_floats = need;
    }
  }

  TrainParams p{};
  p.d = d; p.B = B; p.P = P; p.netP = netP;
  p.x_train = a->x_train; p.n_train = a->n_train; p.perm = (const long long*)a->perm;
  p.batch_size = grad_only ? (int)a->n_train : a->batch_size;
  p.x_valid = grad_only ? nullptr : a->x_valid; p.n_valid = grad_only ? 0 : a->n_valid;
  p.grad_only = grad_only ? 1 : 0; p.batch_total = a->batch_total;
  p.noise = a->noise; p.jitter = (float)a->jitter;
  p.seed_lo = (unsigned int)(a->seed & 0xffffffffu); p.seed_hi = (unsigned int)(a->seed >> 32); p.epoch = a->epoch;
  p.lr = (float)a->lr; p.beta1 = (float)a->beta1; p.beta2 = (float)a->beta2; p.eps = (float)a->eps;
  p.weight_decay = (float)a->weight_decay;
  p.step0 = a->step0;
  p.params = a->params; p.adam_m = a->adam_m; p.adam_v = a->adam_v;
  p.gpart = grid > 1 ? h->d_train_ws : nullptr;
  p.gsum = grid > 1 ? h->d_train_ws + (size_t)grid * Psm : nullptr;
  p.mv_priv = grid > 1 ? h->d_train_ws + (size_t)(grid + 1) * Psm : nullptr;
  p.ctrl = (TrainCtrl*)h->d_train_ctrl;
  p.loss_part = reinterpret_cast<double*>(reinterpret_cast<char*>(h->d_train_ctrl) + kLossOff);
  p.grad_out = a->grad_out;
  p.do_train = a->do_train ? 1 : 0;

  const size_t smem = train_smem_bytes(d, H, L, B);
  int rc;
  if (H == 16 && L == 1) rc = launch_train<16, 1>(h, p, grid, smem, st);
  else if (H == 16 && L == 2) rc = launch_train<16, 2>(h, p, grid, smem, st);
  else if (H == 32 && L == 1) rc = launch_train<32, 1>(h, p, grid, smem, st);
  else rc = launch_train<32, 2>(h, p, grid, smem, st);
  if (rc) return rc;
  const int slot = (int)(h->train_begun & 1u);
  NNB_CUDA(h, cudaMemcpyAsync(static_cast<char*>(h->h_train_ctrl) + slot * train_ctrl_bytes(h), h->d_train_ctrl,
                              train_ctrl_bytes(h), cudaMemcpyDeviceToHost, st));
  NNB_CUDA(h, cudaEventRecord(h->train_ev[slot], st));
  h->train_slot_grid[slot] = grid;
  ++h->train_begun;
  return NNB_OK;
}

extern "C" int nnb_train_epoch_end(nnb_handle* h, double* train_loss_sum, double* val_nll_sum, int* grid_out) {
  if (!h) return NNB_ERR_ARG;
  if (h->train_begun == h->train_ended) return nnb_fail(h, NNB_ERR_STATE, "nnb_train_epoch_end: no epoch in flight");
  const int slot = (int)(h->train_ended & 1u);
  ++h->train_ended;                                   // the slot is given up even if the wait fails
  NNB_CUDA(h, cudaEventSynchronize(h->train_ev[slot]));
  const int grid = h->train_slot_grid[slot];
  // losses: per-CTA partials added in CTA order (deterministic, unlike floating-point atomics)
  const double* lp = reinterpret_cast<const double*>(static_cast<const char*>(h->h_train_ctrl) +
                                                     slot * train_ctrl_bytes(h) + kLossOff);
  double tl = 0.0, vl = 0.0;
  for (int c = 0; c < grid; ++c) { tl += lp[2 * c]; vl += lp[2 * c + 1]; }
  if (train_loss_sum) *train_loss_sum = tl;
  if (val_nll_sum) *val_nll_sum = vl;
  if (grid_out) *grid_out = grid;
  return NNB_OK;
}

extern "C" int nnb_train_epoch(nnb_handle* h, const nnb_train_args* a, void* stream) {
  if (!h || !a) return NNB_ERR_ARG;
  if (h->train_begun != h->train_ended)
    return nnb_fail(h, NNB_ERR_STATE, "nnb_train_epoch: asynchronous epochs are in flight (call nnb_train_epoch_end first)");
  const int rc = nnb_train_epoch_begin(h, a, stream);
  if (rc) return rc;
  return nnb_train_epoch_end(h, a->train_loss_sum_out, a->val_nll_sum_out, a->grid_out);
}

extern "C" int nnb_mean_nn_distance(nnb_handle* h, const double* x, int64_t n, int d, double* out, void* stream) {
  if (!h || !out || n < 0 || (n > 0 && !x) || d < 1 || d > NNB_MAX_DIM) return NNB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  NNB_CUDA(h, cudaSetDevice(h->device));
  *out = 0.0;
  if (n < 2) return NNB_OK;
  const int grid = (int)((n + 127) / 128);
  if (h->nn_part_cap < grid) {
    if (h->d_nn_part) cudaFree(h->d_nn_part);
    h->d_nn_part = nullptr;
    NNB_CUDA(h, cudaMalloc(&h->d_nn_part, sizeof(double) * grid));
    h->nn_part_cap = grid;
  }
  double* acc = h->d_nn_part;
  static const bool f64_only = getenv("NNB_NN_F64") != nullptr;
  static const bool no_tc = getenv("NNB_NN_NO_TC") != nullptr;
  const int K = nn_tc_k(d);
  // query tiles per CTA (4, 2 or 1) and ring stages (4 or 3): the largest that fit the shared memory
  int nn_qt = 0, nn_stages = 3;
  for (int qt : {4, 2, 1})
    if (!nn_qt && nn_tc_smem_bytes(K, qt, 3) <= (size_t)h->max_smem) nn_qt = qt;
  if (nn_qt && nn_tc_smem_bytes(K, nn_qt, 4) <= (size_t)h->max_smem) nn_stages = 4;
  if (const char* e = getenv("NNB_NN_QT")) {          // development override: query tiles per CTA
    const int q = atoi(e);
    if ((q == 1 || q == 2 || q == 4) && nn_tc_smem_bytes(K, q, 3) <= (size_t)h->max_smem) {
      nn_qt = q;
      nn_stages = nn_tc_smem_bytes(K, q, 4) <= (size_t)h->max_smem ? 4 : 3;
    }
  }
  if (const char* e = getenv("NNB_NN_STAGES")) {
    const int q = atoi(e);
    if ((q == 3 || q == 4) && nn_qt && nn_tc_smem_bytes(K, nn_qt, q) <= (size_t)h->max_smem) nn_stages = q;
  }
  if (n >= 8192 && nn_qt && !f64_only && !no_tc) {
    // tensor-core ratings (3xTF32, one augmented dot product per pair) + exact float64 refinement: nnb_nn_tc.cuh
    const int tiles = grid;   // 128 rows per tile
    const int qgroups = (tiles + nn_qt - 1) / nn_qt;
    // candidate splits: ONE wave of CTAs.  Every (query, split, column half) costs at least one exact evaluation, so more
    // splits than it takes to fill the SMs only add work (65 536 x 30, QT = 4: 3.4 / 3.9 / 4.2 ms at 1 / 2 / 3 splits)
    int splits = h->sm_count / qgroups;
    if (splits < 1) splits = 1;
    if (splits > 32) splits = 32;
    if (const char* e = getenv("NNB_NN_SPLITS")) splits = atoi(e) > 0 ? atoi(e) : splits;   // development override
    const int tps = (tiles + splits - 1) / splits;
    splits = (tiles + tps - 1) / tps;
    const int nparts = 64;
    const size_t tile_fl = (size_t)tiles * 128 * K;
    // workspace (floats): 4 packed arrays | centre partials (doubles) | R2 bits | best [splits][n] (doubles)
    const size_t part_off = 4 * tile_fl, r2_off = part_off + (size_t)2 * nparts * d + 2, best_off = (r2_off + 4 + 1) & ~(size_t)1;
    const size_t need = best_off + (size_t)2 * (2 * splits) * n + 4;   // two column halves per split
    if (h->nn_ws_floats < need) {
      if (h->d_nn_ws) cudaFree(h->d_nn_ws);
      h->d_nn_ws = nullptr;
      h->nn_ws_floats = 0;
      NNB_CUDA(h, cudaMalloc(&h->d_nn_ws, need * sizeof(float)));
      h->nn_ws_floats = need;
    }
    float* ws = h->d_nn_ws;
    float *a_hi = ws, *a_lo = ws + tile_fl, *b_hi = ws + 2 * tile_fl, *b_lo = ws + 3 * tile_fl;
    double* part = reinterpret_cast<double*>(ws + part_off);
    unsigned int* r2 = reinterpret_cast<unsigned int*>(ws + r2_off);
    double* best = reinterpret_cast<double*>(ws + best_off);
    NNB_CUDA(h, cudaMemsetAsync(r2, 0, sizeof(unsigned int), st));
    nn_tc_centre_kernel<<<nparts, 256, 0, st>>>(x, n, d, part);
    nn_tc_pack_kernel<<<tiles, 128, (size_t)d * sizeof(double), st>>>(x, n, d, K, part, nparts, a_hi, a_lo, b_hi, b_lo, r2);
    const size_t smem = nn_tc_smem_bytes(K, nn_qt, nn_stages);
    const dim3 g2(qgroups, splits);
#define NNB_NN_LAUNCH(QT)                                                                                             \
  do {                                                                                                               \
    NNB_CUDA(h, nnb_set_smem(nn_tc_kernel<QT>, smem));                                                               \
    nn_tc_kernel<QT><<<g2, kNnThreads, smem, st>>>(a_hi, a_lo, b_hi, b_lo, x, n, d, K, r2, tiles, tps, nn_stages, best); \
  } while (0)
    if (nn_qt == 4) NNB_NN_LAUNCH(4);
    else if (nn_qt == 2) NNB_NN_LAUNCH(2);
    else NNB_NN_LAUNCH(1);
#undef NNB_NN_LAUNCH
    nn_reduce_kernel<<<grid, 128, 0, st>>>(best, n, 2 * splits, acc);
  } else if (d <= 32 && !f64_only) {
    // float32 prefilter + exact float64 refinement (same result as the brute force, ~2.5x faster)
    const int DP = d <= 8 ? 8 : (d <= 16 ? 16 : 32);
    // candidate splits: enough blocks for ~8 per SM
    int splits = (int)((8ll * h->sm_count + grid - 1) / grid);
    if (splits < 1) splits = 1;
    if (splits > 16) splits = 16;
    const size_t nbest = (size_t)splits * n * 2;     // doubles, counted in floats
    const size_t need = (size_t)n * DP + 8 + nbest;
    if (h->nn_ws_floats < need) {
      if (h->d_nn_ws) cudaFree(h->d_nn_ws);
      h->d_nn_ws = nullptr;
      NNB_CUDA(h, cudaMalloc(&h->d_nn_ws, need * sizeof(float)));
      h->nn_ws_floats = need;
    }
    unsigned int* mx = reinterpret_cast<unsigned int*>(h->d_nn_ws + (size_t)n * DP);
    NNB_CUDA(h, cudaMemsetAsync(mx, 0, sizeof(unsigned int), st));
    nn_prepare_kernel<<<h->sm_count * 4, 256, 0, st>>>(x, n, d, DP, h->d_nn_ws, mx);
    double* best = reinterpret_cast<double*>(h->d_nn_ws + (((size_t)n * DP + 4 + 1) & ~(size_t)1));   // 8-byte aligned
    const dim3 g2(grid, splits);
    if (DP == 8) nn_min_dist_f32_kernel<8><<<g2, 128, 0, st>>>(x, h->d_nn_ws, n, d, mx, best);
    else if (DP == 16) nn_min_dist_f32_kernel<16><<<g2, 128, 0, st>>>(x, h->d_nn_ws, n, d, mx, best);
    else nn_min_dist_f32_kernel<32><<<g2, 128, 0, st>>>(x, h->d_nn_ws, n, d, mx, best);
    nn_reduce_kernel<<<grid, 128, 0, st>>>(best, n, splits, acc);
  } else if (d <= 16) {
    nn_min_dist_kernel<16><<<grid, 128, 128 * 16 * sizeof(double), st>>>(x, n, d, acc);
  } else if (d <= 32) {
    nn_min_dist_kernel<32><<<grid, 128, 128 * 32 * sizeof(double), st>>>(x, n, d, acc);
  } else {
    const size_t smem = (size_t)2 * d * 128 * sizeof(double);
    NNB_CUDA(h, nnb_set_smem(nn_min_dist_kernel<0>, smem));
    nn_min_dist_kernel<0><<<grid, 128, smem, st>>>(x, n, d, acc);
  }
  NNB_CUDA(h, cudaGetLastError());
  std::vector<double> part((size_t)grid);
  NNB_CUDA(h, cudaMemcpyAsync(part.data(), acc, sizeof(double) * grid, cudaMemcpyDeviceToHost, st));
  NNB_CUDA(h, cudaStreamSynchronize(st));
  double sum = 0.0;
  for (int b = 0; b < grid; ++b) sum += part[(size_t)b];   // block order: deterministic
  *out = sum / (double)n;
  return NNB_OK;
}
