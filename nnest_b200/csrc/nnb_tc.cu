// nnb_tc.cu -- host side of the tensor-core MCMC kernel: weight packing into the UMMA layout and launch.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nnb_tc_launch.cuh"

using namespace nnb;

// natural (state_dict) order -> per block { B1 | bias1 | L x (B2s B2t bias2) | B3s B3t bias3 }, see nnb_tc_kernels.cuh
int nnb_tc_pack(nnb_handle* h, const float* weights) {
  const FlowDesc& f = h->flow;
  h->tc_ok = false;
  if (!tc_supported(f)) return NNB_OK;
  const int d = f.d, H = 16, L = f.L, B = f.B;
  TcFlowDesc t{};
  t.d = d; t.L = L; t.B = B;
  int off = 0;
  for (int k = 0; k < B; ++k) { t.off[k] = off; off += tc_block_floats(d, L, k); }
  t.total_floats = off;
  std::vector<float> buf((size_t)off, 0.f);
  TcConsts cst{};
  const size_t net_nat = (size_t)H * d + H + (size_t)L * (H * H + H) + (size_t)d * H + d;
  for (int k = 0; k < B; ++k) {
    const int nin = blk_nin(d, k), i0 = blk_i0(k), nout = blk_nout(d, k), o0 = blk_o0(k);
    const int K1 = round8(nin), N3 = round16(nout);
    const float* net[2] = {weights + (size_t)(2 * k) * net_nat, weights + (size_t)(2 * k + 1) * net_nat};
    float* dst = buf.data() + t.off[k];
    // layer 1: rows 0-15 scale net, 16-31 translate net; columns = masked inputs
    std::vector<float> w1((size_t)32 * K1, 0.f);
    // (scale net: constants of its activations folded into the weights, see kTcTanhScale / kTcExpScale)
    const float sc_h[2] = {kTcTanhScale, 1.0f}, sc_o[2] = {kTcExpScale, 1.0f};
    for (int s = 0; s < 2; ++s)
      for (int j = 0; j < H; ++j)
        for (int a = 0; a < nin; ++a) w1[(size_t)(16 * s + j) * K1 + a] = sc_h[s] * net[s][(size_t)j * d + (i0 + 2 * a)];
    tc::host_pack_b(w1.data(), 32, K1, K1, 32, K1, dst, dst + 32 * K1);
    float* bias1 = dst + 64 * K1;
    float* cb = cst.v + tc_cb_off(d, L, k);          // the same biases in the kernel-parameter layout (TcConsts)
    for (int s = 0; s < 2; ++s)
      for (int j = 0; j < H; ++j) cb[16 * s + j] = bias1[16 * s + j] = sc_h[s] * net[s][(size_t)H * d + j];
    float* o = bias1 + 32;
    size_t nat_off = (size_t)H * d + H;
    for (int l = 0; l < L; ++l) {
      for (int s = 0; s < 2; ++s) {
        std::vector<float> w2((size_t)H * H);
        for (int q = 0; q < H * H; ++q) w2[(size_t)q] = sc_h[s] * net[s][nat_off + q];
        tc::host_pack_b(w2.data(), 16, 16, 16, 16, 16, o + 512 * s, o + 512 * s + 256);
      }
      for (int s = 0; s < 2; ++s)
        for (int j = 0; j < H; ++j)
          cb[32 + 32 * l + 16 * s + j] = o[1024 + 16 * s + j] = sc_h[s] * net[s][nat_off + (size_t)H * H + j];
      o += 1056;
      nat_off += (size_t)H * H + H;
    }
    std::vector<float> w3((size_t)N3 * 16, 0.f);
    for (int s = 0; s < 2; ++s) {
      std::fill(w3.begin(), w3.end(), 0.f);
      for (int q = 0; q < nout; ++q)
        for (int j = 0; j < H; ++j) w3[(size_t)q * 16 + j] = sc_o[s] * net[s][nat_off + (size_t)(o0 + 2 * q) * H + j];
      tc::host_pack_b(w3.data(), N3, 16, 16, N3, 16, o + (size_t)32 * N3 * s, o + (size_t)32 * N3 * s + 16 * N3);
    }
    float* bias3 = o + 64 * N3;
    for (int s = 0; s < 2; ++s)
      for (int q = 0; q < nout; ++q)
        cb[32 + 32 * L + N3 * s + q] = bias3[N3 * s + q] = sc_o[s] * net[s][nat_off + (size_t)d * H + (o0 + 2 * q)];
  }
  if (tc_smem_bytes(t, target_doubles(d, NNB_MAX_LIKE_PARAMS), 1, 2) > (size_t)h->max_smem) return NNB_OK;
  NNB_CUDA(h, nnb_reserve(&h->d_weights_tc, &h->weights_tc_cap, buf.size()));
  NNB_CUDA(h, cudaMemcpy(h->d_weights_tc, buf.data(), buf.size() * sizeof(float), cudaMemcpyHostToDevice));
  h->tcflow = t;
  h->tc_consts = cst;
  h->tc_ok = true;
  return NNB_OK;
}

int nnb_launch_mcmc_tc(nnb_handle* h, McmcParams p, int steps, cudaStream_t st) {
  // Threads per chain.  One thread per chain (128 registers, the current latent point resident in shared memory, fewer
  // tile barriers) measured faster than two threads per chain at every batch size tried on the B200 (1 k - 131 k chains),
  // so it is the default; NNB_TC_NPART=2 selects the two-thread layout (the chain's threads split every per-column
  // phase and the second one draws the next step's noise during the accept test).
  const char* npart_env = getenv("NNB_TC_NPART");     // read at every launch (the tests switch layouts in-process)
  const int npart = npart_env && atoi(npart_env) == 2 ? 2 : 1;
  // the reference's default architecture at the dimensions of the named workloads: fully unrolled kernels
  static const bool generic_only = getenv("NNB_TC_GENERIC") != nullptr;
  if (!generic_only && h->tcflow.L == 1 && h->tcflow.B == 3) {
    switch (h->tcflow.d) {
      case 2: return nnb_launch_mcmc_tc_fixed<2>(h, p, steps, st, npart);
      case 10: return nnb_launch_mcmc_tc_fixed<10>(h, p, steps, st, npart);
      case 30: return nnb_launch_mcmc_tc_fixed<30>(h, p, steps, st, npart);
      case 50: return nnb_launch_mcmc_tc_fixed<50>(h, p, steps, st, npart);
      default: break;
    }
  }
  if (npart == 1)
    return p.mode == NNB_MODE_MH ? launch_tc_mode<NNB_MODE_MH, 1, 0>(h, p, steps, st)
                                 : launch_tc_mode<NNB_MODE_HARD, 1, 0>(h, p, steps, st);
  return p.mode == NNB_MODE_MH ? launch_tc_mode<NNB_MODE_MH, 2, 0>(h, p, steps, st)
                               : launch_tc_mode<NNB_MODE_HARD, 2, 0>(h, p, steps, st);
}
