// nnb_nn_tc.cuh -- mean nearest-neighbour distance (training jitter, reference nnest/trainer.py:147-150: 0.2 * mean of
// cKDTree(samples).query(samples, 2)) with the pair ratings on the tensor cores (tcgen05, 3xTF32) and an exact float64
// refinement.  The result is the one the float64 brute force finds, bit for bit.
//
// Rating = squared distance as ONE dot product of augmented rows: with xc = float32(x - centre),
//     a(q) = [ xc_1 .. xc_d, 1, |xc|^2, 0 .. ]        (query side,      K = round8(d + 2) columns)
//     b(c) = [ -2 xc_1 .. -2 xc_d, |xc|^2, 1, 0 .. ]   (candidate side)
//     a(q) . b(c) = |xc_q|^2 + |xc_c|^2 - 2 xc_q . xc_c = |xc_q - xc_c|^2,
// so a 128 x 128 tile of ratings is 3 * K / 8 tcgen05.mma (M = N = 128, kind::tf32, both operands in shared memory in the
// canonical K-major core-matrix layout of nnb_tc.cuh) and the epilogue only has to compare: one FMNMX per rating.
// Error bound: every operand enters as hi + lo (tf32 halves: 2^-21 relative), the lo * lo product is dropped (2^-22), the
// FP32 accumulation of 3 K terms rounds by at most 2^-23 of the running sum each, and the centred coordinates are rounded
// to float32 once (2^-24 each); with S = sum |a_i b_i| <= 4 R2 (R2 = max |xc|^2 over the rows)
//     |rating - d64^2| <= E := R2 * (2e-5 + 4e-6 K)        (conservative by more than 2x)
// for every pair.  TWO passes over the candidates: pass 1 finds m = the smallest rating of a candidate other than the
// query itself (pure min-tree, no branches); pass 2 rates the candidates again and re-evaluates exactly (float64, from the
// original rows) those with rating <= m + 2 E -- typically one or two per query.  The true nearest neighbour c* is among
// them: rating(c*) <= d64^2(c*) + E <= d64^2(c_m) + E <= rating(c_m) + 2 E.  (One pass with a running threshold needs an
// exact evaluation at every new record, ~ln n per query and split, each a chain of L2 latencies that the whole warp waits
// for.)  The centre only has to be SOME point near the cloud (it shrinks R2, i.e. the number of exact re-evaluations), it
// does not enter the result.
// Measured at the config-4 retrain size (65 536 x 30, B200): float64 brute force 25 ms, float32 FFMA prefilter 15.4 ms,
// this path 3.4 ms (the tensor-core work of the two passes is 1.5 ms; first version -- one pass, one query tile per CTA --
// 5.3 ms, L2-bandwidth bound).
//
// Kernel structure (one CTA per (QT x 128 queries, candidate split), 288 threads):
//   warp 8, one elected lane: TMA bulk copies (cp.async.bulk) of the packed candidate tiles into a 3- or 4-stage shared-memory
//     ring, the MMAs of one tile into one of four 128-column TMEM accumulators, tcgen05.commit to "accumulator full" and
//     "stage free" mbarriers;
//   warps 0-7 (two threads per query row = TMEM lane, one per half of the tile's columns; each half is rated like a split of
//     its own): wait "accumulator full", two tcgen05.ld of 32 columns, arrive on "accumulator free", min tree + compare, rare
//     exact path.
#pragma once
#include "nnb_tc.cuh"

namespace nnb {

constexpr int kNnMaxStages = 4;   // candidate-tile ring: as many stages (3 or 4) as fit the shared memory
__host__ __device__ inline int nn_tc_k(int d) { return (d + 2 + 7) & ~7; }
__host__ __device__ inline float nn_tc_slack(float r2, int K) { return r2 * (2e-5f + 4e-6f * (float)K) + 1e-30f; }
__host__ inline size_t nn_tc_smem_bytes(int K, int qt, int stages) { return (size_t)(2 * qt + 2 * stages) * 128 * K * 4 + 512; }

// float offset of element (row r of a 128-row tile, column k) in the packed tile (K-major core matrices, nN = 16)
__host__ __device__ inline int nn_tc_off(int r, int k) {
  return (k >> 3) * 1024 + (((k & 7) >> 2) * 16 + (r >> 3)) * 32 + (r & 7) * 4 + (k & 3);
}

// stage 1 of the centre: per-block partial sums of the coordinates in a fixed order (64 blocks x 256 threads:
// thread = (row lane 0..7, dim 0..31)); the pack kernel adds the 64 partials in block order -> reproducible
__global__ void __launch_bounds__(256) nn_tc_centre_kernel(const double* __restrict__ x, long long n, int d,
                                                           double* __restrict__ part /* [gridDim.x][d] */) {
  __shared__ double sh[8][32];
  const int dim = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per, r1 = r0 + per < n ? r0 + per : n;
  for (int i0 = 0; i0 < d; i0 += 32) {
    const int i = i0 + dim;
    double s = 0.0;
    if (i < d)
      for (long long r = r0 + lane; r < r1; r += 8) s += x[r * d + i];
    sh[lane][dim] = s;
    __syncthreads();
    if (lane == 0 && i < d) {
      double t = 0.0;
      for (int l = 0; l < 8; ++l) t += sh[l][dim];
      part[(size_t)blockIdx.x * d + i] = t;
    }
    __syncthreads();
  }
}

// one thread per (padded) row: centred float32 coordinates, augmented query / candidate rows, tf32 hi / lo halves, written
// into the packed tiles; max |xc|^2 over the rows (bits of a non-negative float)
__global__ void __launch_bounds__(128) nn_tc_pack_kernel(const double* __restrict__ x, long long n, int d, int K,
                                                         const double* __restrict__ part, int nparts,
                                                         float* __restrict__ a_hi, float* __restrict__ a_lo,
                                                         float* __restrict__ b_hi, float* __restrict__ b_lo,
                                                         unsigned int* __restrict__ r2_bits) {
  extern __shared__ double centre[];   // [d]
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += part[(size_t)p * d + i];
    centre[i] = s / (double)n;
  }
  __syncthreads();
  const long long row = (long long)blockIdx.x * 128 + threadIdx.x;   // blockIdx.x = tile
  const int r = threadIdx.x;
  const bool valid = row < n;
  const size_t tile_base = (size_t)blockIdx.x * 128 * K;
  float qq = 0.f;
  // two passes over the row keep the register footprint independent of d: first |xc|^2, then the columns
  if (valid)
    for (int i = 0; i < d; ++i) {
      const float v = (float)(x[row * d + i] - centre[i]);
      qq = fmaf(v, v, qq);
    }
  for (int k = 0; k < K; ++k) {
    float av = 0.f, bv = 0.f;
    if (valid) {
      if (k < d) {
        const float v = (float)(x[row * d + k] - centre[k]);
        av = v;
        bv = -2.0f * v;
      } else if (k == d) {
        av = 1.0f;
        bv = qq;
      } else if (k == d + 1) {
        av = qq;
        bv = 1.0f;
      }
    } else if (k == d) {
      bv = 1e30f;   // padding candidates can never be rated near a minimum (padding queries are never written)
    }
    uint32_t h, l;
    const size_t o = tile_base + nn_tc_off(r, k);
    tc::split_tf32(av, h, l);
    a_hi[o] = __uint_as_float(h);
    a_lo[o] = __uint_as_float(l);
    tc::split_tf32(bv, h, l);
    b_hi[o] = __uint_as_float(h);
    b_lo[o] = __uint_as_float(l);
  }
  float m = valid ? qq : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(r2_bits, __float_as_uint(m));
}

namespace nn_tc_detail {
// D[tmem] (+)= A[smem] * B[smem]^T, one K = 8 step; issued by ONE thread
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// descriptor of one K = 8 step of a packed 128-row tile: LBO (K direction) = 16 * 128 B, SBO (row direction) = 128 B
__device__ __forceinline__ uint64_t tile_desc(uint32_t smem_addr, int kstep) {
  return tc::smem_desc_kmajor(smem_addr + (uint32_t)kstep * 4096u, 2048u, 128u);
}
__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar) { tc::mbar_arrive(bar); }
}  // namespace nn_tc_detail

constexpr int kNnAcc = 4;            // 128-column TMEM accumulators (all 512 columns)
constexpr int kNnThreads = 288;      // 8 epilogue warps + 1 producer warp

// QT = query tiles (of 128 rows) per CTA.  Every CTA streams its split's candidate tiles from L2 -- with one query tile per
// CTA that is 512 x 16.8 MB = 17 GB per call at 65 536 x 30 and the kernel is L2-bandwidth bound (measured: 4.7 ms, tensor
// pipe 33 % busy); with QT query tiles resident the same candidate tile feeds QT sets of MMAs and the traffic drops QT-fold.
template <int QT>
__global__ void __launch_bounds__(kNnThreads, 1)
nn_tc_kernel(const float* __restrict__ a_hi, const float* __restrict__ a_lo, const float* __restrict__ b_hi,
             const float* __restrict__ b_lo, const double* __restrict__ x, long long n, int d, int K,
             const unsigned int* __restrict__ r2_bits, int tiles, int tiles_per_split, int stages,
             double* __restrict__ best_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tile_floats = 128 * K;
  float* q_all = reinterpret_cast<float*>(smem_raw);          // QT x { hi, lo }
  float* c_ring = q_all + (size_t)2 * QT * tile_floats;       // stages x { hi, lo }
  uint64_t* bars = reinterpret_cast<uint64_t*>(c_ring + (size_t)2 * stages * tile_floats);
  uint64_t* full = bars;                     // [stages] candidate tile landed
  uint64_t* empty = full + kNnMaxStages;     // [stages] MMAs that read the stage have completed
  uint64_t* tfull = empty + kNnMaxStages;    // [kNnAcc] accumulator written
  uint64_t* tempty = tfull + kNnAcc;         // [kNnAcc] accumulator read by the eight epilogue warps
  uint64_t* qbar = tempty + kNnAcc;
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(qbar + 1);

  const int warp = threadIdx.x >> 5;
  const int t0 = blockIdx.y * tiles_per_split;
  const int t1 = t0 + tiles_per_split < tiles ? t0 + tiles_per_split : tiles;
  const int nt = t1 - t0;
  const uint32_t bytes = (uint32_t)tile_floats * 4u;
  const int qt0 = blockIdx.x * QT;                                  // first query tile of this CTA
  const int nq = tiles - qt0 < QT ? tiles - qt0 : QT;               // query tiles that exist

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < kNnAcc; ++b) {
      tc::mbar_init(&tfull[b], 1);
      tc::mbar_init(&tempty[b], 8);
    }
    tc::mbar_init(qbar, 1);
    tc::mbar_fence_init();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 8) tc::tmem_alloc(tmem_base_s, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_base_s;
  // work item w = (candidate tile of the two-pass sequence) * nq + query tile; accumulator w % kNnAcc
  const int nt2 = 2 * nt;

  if (warp == 8) {
    // ---- producer: TMA loads + MMA issue ------------------------------------------------------------------------------
    if (tc::elect_one()) {
      tc::mbar_expect_tx(qbar, 2 * bytes * (uint32_t)nq);
      for (int qi = 0; qi < nq; ++qi) {
        const size_t qoff = (size_t)(qt0 + qi) * tile_floats;
        tc::tma_bulk_g2s(q_all + (size_t)2 * qi * tile_floats, a_hi + qoff, bytes, qbar);
        tc::tma_bulk_g2s(q_all + (size_t)(2 * qi + 1) * tile_floats, a_lo + qoff, bytes, qbar);
      }
      const uint32_t idesc = tc::idesc_tf32_m128(128);
      const uint32_t q0 = tc::smem_u32(q_all);
      const int ksteps = K >> 3;
      // The loads run `ahead` = stages - 2 tiles in front of the MMAs: the stage that tile i overwrites was read by the MMAs
      // of tile i - stages, issued two iterations ago.
      const int ahead = stages - 2;
      int w = 0;
      for (int i = 0; i < nt2 + ahead; ++i) {
        if (i < nt2) {   // load candidate tile i
          const int s = i % stages;
          if (i >= stages) tc::mbar_wait(&empty[s], (uint32_t)(((i / stages) - 1) & 1));
          const size_t coff = (size_t)(t0 + (i >= nt ? i - nt : i)) * tile_floats;
          float* dst = c_ring + (size_t)2 * s * tile_floats;
          tc::mbar_expect_tx(&full[s], 2 * bytes);
          tc::tma_bulk_g2s(dst, b_hi + coff, bytes, &full[s]);
          tc::tma_bulk_g2s(dst + tile_floats, b_lo + coff, bytes, &full[s]);
        }
        const int j = i - ahead;
        if (j >= 0) {   // MMAs of candidate tile j (of the sequence of both passes) against every resident query tile
          const int s = j % stages;
          if (j == 0) tc::mbar_wait(qbar, 0u);
          tc::mbar_wait(&full[s], (uint32_t)((j / stages) & 1));
          const uint32_t ch = tc::smem_u32(c_ring + (size_t)2 * s * tile_floats), cl = ch + bytes;
          for (int qi = 0; qi < nq; ++qi, ++w) {
            const int b = w % kNnAcc;
            if (w >= kNnAcc) tc::mbar_wait(&tempty[b], (uint32_t)(((w / kNnAcc) - 1) & 1));
            tc::fence_after_sync();
            const uint32_t qh = q0 + (uint32_t)(2 * qi) * bytes, ql = qh + bytes;
            const uint32_t dcol = tmem + (uint32_t)b * 128u;
            uint32_t acc = 0u;
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t dqh = nn_tc_detail::tile_desc(qh, ks), dql = nn_tc_detail::tile_desc(ql, ks);
              const uint64_t dch = nn_tc_detail::tile_desc(ch, ks), dcl = nn_tc_detail::tile_desc(cl, ks);
              nn_tc_detail::mma_tf32_ss(dcol, dql, dch, idesc, acc);   // small terms first
              nn_tc_detail::mma_tf32_ss(dcol, dqh, dcl, idesc, 1u);
              nn_tc_detail::mma_tf32_ss(dcol, dqh, dch, idesc, 1u);
              acc = 1u;
            }
            tc::mma_commit(&tfull[b]);
          }
          tc::mma_commit(&empty[s]);
        }
      }
    }
    __syncwarp();
  } else {
    // ---- epilogue: two threads per query row (TMEM lane m), one per half of a tile's 128 columns -------------------------------
    const int m = threadIdx.x & 127, half = threadIdx.x >> 7;   // warps w and w + 4 share lane quarter w
    const float E = nn_tc_slack(__uint_as_float(*r2_bits), K);
    const uint32_t lane_tmem = tmem + (((uint32_t)(m >> 5) * 32u) << 16) + (uint32_t)half * 64u;
    float best32[QT], thr[QT];
    double best64[QT];
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) {
      best32[qi] = INFINITY;
      thr[qi] = INFINITY;
      best64[qi] = INFINITY;
    }
    auto exact = [&](long long row, long long cand) -> double {   // (inlined once: see the rolled loop below)
      if (cand == row || cand >= n) return INFINITY;
      const double* a = x + row * d;
      const double* b = x + cand * d;
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
      return e0 + e1;
    };
    int w = 0;
    for (int jj = 0; jj < nt2; ++jj) {
      const int pass = jj >= nt ? 1 : 0, j = pass ? jj - nt : jj;
      const long long c0 = (long long)(t0 + j) * 128 + 64 * half;   // first candidate of this thread's 64 columns
#pragma unroll
      for (int qi = 0; qi < QT; ++qi) {
        if (qi >= nq) break;
        const int b = w % kNnAcc;
        if (jj == nt) thr[qi] = best32[qi] + 2.0f * E;   // pass 1 is complete
        tc::mbar_wait(&tfull[b], (uint32_t)((w / kNnAcc) & 1));
        ++w;
        tc::fence_after_sync();
        uint32_t r0[32], r1[32];
        tc::tmem_ld32(lane_tmem + (uint32_t)b * 128u, r0);
        tc::tmem_ld32(lane_tmem + (uint32_t)b * 128u + 32u, r1);
        tc::wait_ld();
        // the accumulator is free as soon as its values are in registers
        tc::fence_before_sync();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) nn_tc_detail::mbar_arrive_cnt(&tempty[b]);
        const long long row = (long long)(qt0 + qi) * 128 + m;
        // pass 1: the query's own rating (~0) does not count
        const long long selfq = pass == 0 ? row - c0 : -1;
        const int self = (selfq >= 0 && selfq < 64) ? (int)selfq : -1;
        float v[32];   // min tree (five levels) over the 64 ratings
#pragma unroll
        for (int q = 0; q < 32; ++q)
          v[q] = fminf(q == self ? INFINITY : __uint_as_float(r0[q]), q + 32 == self ? INFINITY : __uint_as_float(r1[q]));
#pragma unroll
        for (int ww = 16; ww > 0; ww >>= 1)
#pragma unroll
          for (int q = 0; q < ww; ++q) v[q] = fminf(v[q], v[q + ww]);
        const float mn = v[0];
        if (pass == 0) {
          best32[qi] = fminf(best32[qi], mn);
        } else if (row < n && mn <= thr[qi]) {   // some rating is within the slack of the minimum
          // (bit mask + rolled loop: ONE copy of the exact evaluation in the code -- unrolled over the 64 ratings and the
          // query tiles the kernel was 25 k instructions and ran at a tenth of the speed)
          unsigned long long mask = 0ull;
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            mask |= __uint_as_float(r0[q]) <= thr[qi] ? (1ull << q) : 0ull;
            mask |= __uint_as_float(r1[q]) <= thr[qi] ? (1ull << (q + 32)) : 0ull;
          }
          double bq = best64[qi];
#pragma unroll 1
          while (mask) {
            const int q = __ffsll((long long)mask) - 1;
            mask &= mask - 1ull;
            bq = fmin(bq, exact(row, c0 + q));
          }
          best64[qi] = bq;
        }
      }
    }
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) {
      const long long row = (long long)(qt0 + qi) * 128 + m;
      if (qi < nq && row < n) best_out[((long long)blockIdx.y * 2 + half) * n + row] = best64[qi];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

}  // namespace nnb
