// CUDA kernels (sm_100a) of the bucket-method multi-scalar multiplication and the batched scalar
// multiplication.  Thread bodies live in msm_body.cuh; this file is the launch geometry.
//
// Replaces /root/reference/src/abstract/curve.ts:863-905 `pippenger(c, points, scalars)`:
// same result point (compared as canonical affine), different schedule.  The reference walks
// unsigned c-bit windows MSB->LSB on one thread with complete projective additions; here
//
//   k_prepare      canonical affine -> Montgomery "prepared" affine, range validation     (N threads)
//   k_digits<0>    signed c-bit digit recoding + per-(window,bucket) histogram            (N threads)
//   k_scan_*       exclusive scan of the W*B histogram -> bucket offsets                  (2 passes)
//   k_digits<1>    counting-sort scatter of (point index | sign) into bucket order        (N threads)
//   k_accumulate   every thread folds exactly L consecutive sorted entries with mixed
//                  additions, emitting complete buckets or head/tail partials at bucket
//                  boundaries — constant work per thread whatever the bucket sizes are     (T/L threads)
//   k_reduce1      per-chunk running sums (2 additions per bucket), stitching the partials
//                  of buckets that straddle accumulate segments on the fly                 (W*B/K threads)
//   k_reduce2      per-window second level: suffix scan + reduction of the chunk sums
//                  (registers -> warp shuffles -> shared memory)                           (W blocks)
//   k_final        Horner over windows (c doublings each) + one inversion to affine;
//                  one warp, lanes share each formula's independent multiplications      (1 warp)
//
// No atomics touch curve points, so degenerate inputs (all scalars equal, all points equal —
// test/point.test.ts:842-853, benchmark/msm_timings.ts:45-63) stay correct; they only lengthen the
// serial stitch in k_reduce1.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "msm_body.cuh"

namespace nmsm {

template <class Cv>
__global__ void k_prepare(const uint32_t* __restrict__ pts, uint32_t n, uint32_t* __restrict__ aff,
                          unsigned int* err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) prepare_body<Cv>(i, n, pts, aff, err);
}

template <class Cv, bool SCATTER>
__global__ void k_digits(const uint32_t* __restrict__ scalars, uint32_t n, MsmPlan plan,
                         unsigned int* __restrict__ counts_or_cursor, uint32_t* __restrict__ sorted,
                         unsigned int* err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) digits_body<Cv, SCATTER>(i, n, scalars, plan, counts_or_cursor, sorted, err);
}

// Exclusive scan of the G bucket counters in two coalesced passes over SCAN_TILE-element tiles:
//   k_scan_tiles   per-tile totals
//   k_scan_apply   tile prefix (every block sums the totals before it) + in-tile scan -> offsets, cursor
static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_PER_THREAD = 8;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

__device__ __forceinline__ uint32_t block_sum_256(uint32_t v, uint32_t* sh) {
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t t = 0;
  for (int q = 0; q < SCAN_THREADS / 32; q++) t += sh[q];
  __syncthreads();
  return t;
}

static __global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tiles(const unsigned int* __restrict__ counts, uint32_t G, uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t sh[SCAN_THREADS / 32];
  const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; k++)
    if (base + k < G) v += counts[base + k];
  uint32_t t = block_sum_256(v, sh);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = t;
}

static __global__ void __launch_bounds__(SCAN_THREADS)
k_scan_apply(const unsigned int* __restrict__ counts, uint32_t G, const uint32_t* __restrict__ tile_sums,
             uint32_t* __restrict__ offsets, unsigned int* __restrict__ cursor) {
  __shared__ uint32_t sh[SCAN_THREADS / 32];
  __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
  // prefix of the tiles before this one
  uint32_t pre = 0;
  for (uint32_t k = threadIdx.x; k < blockIdx.x; k += SCAN_THREADS) pre += tile_sums[k];
  const uint32_t tile_prefix = block_sum_256(pre, sh);
  const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
  uint32_t c[SCAN_PER_THREAD];
  uint32_t mine = 0;
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; k++) {
    c[k] = (base + k < G) ? counts[base + k] : 0;
    mine += c[k];
  }
  // exclusive scan of `mine` across the block: warp inclusive scan + warp totals
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = mine;
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= (uint32_t)d) inc += o;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  uint32_t wpre = 0;
  for (uint32_t q = 0; q < warp; q++) wpre += warp_tot[q];
  uint32_t run = tile_prefix + wpre + inc - mine;
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; k++) {
    if (base + k < G) {
      offsets[base + k] = run;
      cursor[base + k] = run;
    }
    run += c[k];
  }
  // the thread that owns element G-1 also publishes the grand total
  if (base < G && base + SCAN_PER_THREAD >= G) offsets[G] = run;
}

#ifndef NMSM_ACC_MINBLOCKS
#define NMSM_ACC_MINBLOCKS 1
#endif
template <class Cv>
__global__ void __launch_bounds__(128, NMSM_ACC_MINBLOCKS)
k_accumulate(const uint32_t* __restrict__ aff, const uint32_t* __restrict__ sorted,
             const uint32_t* __restrict__ offsets, MsmPlan plan, uint32_t* __restrict__ buckets,
             uint32_t* __restrict__ heads, uint32_t* __restrict__ tails) {
  accumulate_body<Cv>(blockIdx.x * blockDim.x + threadIdx.x, aff, sorted, offsets, plan, buckets, heads, tails);
}

template <class Cv>
__global__ void __launch_bounds__(128)
k_reduce1(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ buckets,
          const uint32_t* __restrict__ heads, const uint32_t* __restrict__ tails, MsmPlan plan,
          uint32_t* __restrict__ sums, uint32_t* __restrict__ wsums) {
  uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id < (uint32_t)plan.W * plan.chunks) reduce1_body<Cv>(id, offsets, buckets, heads, tails, plan, sums, wsums);
}

template <class G>
__device__ __forceinline__ typename G::Acc shfl_down_acc(const typename G::Acc& a, int delta) {
  typename G::Acc r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&a);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int k = 0; k < G::ACC_WORDS; k++) d[k] = __shfl_down_sync(0xffffffffu, s[k], delta);
  return r;
}

// Second level of the bucket reduction, one block per window (see reduce2_serial for the maths):
//   window_out[w] = sum_k wsums_k + K * sum_k k * sums_k ,  k over the M = B/K chunks of the window.
// Thread j owns R consecutive chunks [jR, (j+1)R): serial running sums give its total s_j, its local
// weighted sum w_j = sum (k - jR) sums_k and its wsums total; the cross-thread term R * sum_j j * s_j is
// a suffix scan of s_j (warp shuffles + one shared-memory hop) followed by a block reduction.
static constexpr int REDUCE2_THREADS = 256;
template <class Cv>
__global__ void __launch_bounds__(REDUCE2_THREADS)
k_reduce2(const uint32_t* __restrict__ sums, const uint32_t* __restrict__ wsums, MsmPlan plan,
          uint32_t* __restrict__ window_out) {
  using G = typename Cv::G;
  using Acc = typename G::Acc;
  extern __shared__ uint32_t smem[];  // (REDUCE2_THREADS/32) accumulators
  const uint32_t w = blockIdx.x, tid = threadIdx.x;
  const uint32_t lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t NW = REDUCE2_THREADS / 32;
  const uint32_t M = plan.chunks;
  const uint32_t R = M >= REDUCE2_THREADS ? M / REDUCE2_THREADS : 1;  // power of two
  const uint32_t lo = tid * R;
  Acc s = G::identity(), wl = G::identity(), wt = G::identity();
  if (lo < M) {
    for (uint32_t k = lo + R; k-- > lo;) {
      const size_t id = (size_t)w * M + k;
      nl_add<G>(wt, load_acc<G>(wsums + id * G::ACC_WORDS));
      nl_add<G>(s, load_acc<G>(sums + id * G::ACC_WORDS));
      if (k > lo) nl_add<G>(wl, s);  // after the loop: sum (k - lo) * sums_k
    }
  }
  // inclusive suffix scan of s over the block: ss_j = sum_{i >= j} s_i
  Acc ss = s;
  for (int d = 1; d < 32; d <<= 1) {
    Acc o = shfl_down_acc<G>(ss, d);
    if (lane + d < 32) nl_add<G>(ss, o);
  }
  auto put = [&](uint32_t slot, const Acc& a) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&a);
    for (int k = 0; k < G::ACC_WORDS; k++) smem[slot * G::ACC_WORDS + k] = src[k];
  };
  auto get = [&](uint32_t slot) {
    Acc a;
    uint32_t* dst = reinterpret_cast<uint32_t*>(&a);
    for (int k = 0; k < G::ACC_WORDS; k++) dst[k] = smem[slot * G::ACC_WORDS + k];
    return a;
  };
  if (lane == 0) put(warp, ss);  // warp totals
  __syncthreads();
  for (uint32_t q = warp + 1; q < NW; q++) nl_add<G>(ss, get(q));
  __syncthreads();
  // per-thread contribution: wt + K * (wl + R * [j >= 1] ss_j)
  Acc v = G::identity();
  if (tid >= 1) {
    v = ss;
    for (uint32_t r = 1; r < R; r <<= 1) nl_dbl<G>(v);
  }
  nl_add<G>(v, wl);
  for (int j = 1; j < plan.K; j <<= 1) nl_dbl<G>(v);
  nl_add<G>(v, wt);
  // block reduction
  for (int d = 16; d >= 1; d >>= 1) {
    Acc o = shfl_down_acc<G>(v, d);
    nl_add<G>(v, o);
  }
  if (lane == 0) put(warp, v);
  __syncthreads();
  if (tid == 0) {
    for (uint32_t q = 1; q < NW; q++) nl_add<G>(v, get(q));
    save_acc<G>(window_out + (size_t)w * G::ACC_WORDS, v);
  }
}

// Horner over the window sums (curve.ts:901-902) by ONE warp whose lanes hold replicated state and
// split the independent multiplications of every point formula between them (ec.cuh Par4).
template <class Cv, bool AFFINE_OUT>
__global__ void __launch_bounds__(32)
k_final(const uint32_t* __restrict__ window_out, MsmPlan plan, uint32_t* __restrict__ out,
        uint32_t* __restrict__ out_inf) {
  using G = typename Cv::G;
  typename G::Acc acc = G::identity();
  for (int w = plan.W - 1; w >= 0; w--) {
    if (w != plan.W - 1)
      for (int j = 0; j < plan.c; j++) G::par_dbl(acc);
    G::par_add(acc, load_acc<G>(window_out + (size_t)w * G::ACC_WORDS));
  }
  if (AFFINE_OUT) {
    uint32_t xy[G::IN_WORDS];
    uint32_t inf;
    nl_to_affine<G>(acc, xy, &inf);
    if (threadIdx.x == 0) {
      for (int k = 0; k < G::IN_WORDS; k++) out[k] = xy[k];
      *out_inf = inf;
    }
  } else if (threadIdx.x == 0) {
    save_acc<G>(out, acc);
  }
}

template <class Cv>
__global__ void __launch_bounds__(32)
k_fold(const uint32_t* __restrict__ accs, int count, uint32_t* __restrict__ out, uint32_t* __restrict__ out_inf) {
  using G = typename Cv::G;
  typename G::Acc acc = G::identity();
  for (int i = 0; i < count; i++) G::par_add(acc, load_acc<G>(accs + (size_t)i * G::ACC_WORDS));
  uint32_t xy[G::IN_WORDS];
  uint32_t inf;
  nl_to_affine<G>(acc, xy, &inf);
  if (threadIdx.x == 0) {
    for (int k = 0; k < G::IN_WORDS; k++) out[k] = xy[k];
    *out_inf = inf;
  }
}

template <class Cv>
__global__ void __launch_bounds__(128)
k_mul_batch(const uint32_t* __restrict__ pts, const uint32_t* __restrict__ scalars, uint32_t n,
            int allow_zero, uint32_t* __restrict__ out_xy, uint32_t* __restrict__ out_inf,
            unsigned int* err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mul_body<Cv>(i, pts, scalars, allow_zero, out_xy, out_inf, err);
}

}  // namespace nmsm
