// CUDA kernels (sm_100a) of the bucket-method multi-scalar multiplication and the batched scalar
// multiplication.  Thread bodies live in msm_body.cuh; this file is the launch geometry.
//
// Replaces /root/reference/src/abstract/curve.ts:863-905 `pippenger(c, points, scalars)`:
// same result point (compared as canonical affine), different schedule.  The reference walks
// unsigned c-bit windows MSB->LSB on one thread with complete projective additions; here
//
//   k_prepare      canonical affine -> Montgomery "prepared" affine, range validation     (N threads)
//   k_digits<0>    signed c-bit digit recoding + per-(window,bucket) histogram            (N threads)
//   k_scan         exclusive scan of the W*B histogram -> bucket offsets                  (1 block)
//   k_digits<1>    counting-sort scatter of (point index | sign) into bucket order        (N threads)
//   k_accumulate   every thread folds exactly L consecutive sorted entries with mixed
//                  additions, emitting complete buckets or head/tail partials at bucket
//                  boundaries — constant work per thread whatever the bucket sizes are     (T/L threads)
//   k_fixup        stitches partials of buckets that straddle thread segments             (W*B threads)
//   k_reduce       per-chunk running sums  sum_{b}(b+1)*B_b  (2 additions per bucket)     (W*B/K threads)
//   k_window_sum   per-window tree reduction (registers -> warp shuffles -> shared)       (W blocks)
//   k_final        Horner over windows (c doublings each) + one inversion to affine       (1 thread)
//
// No atomics touch curve points, so degenerate inputs (all scalars equal, all points equal —
// test/point.test.ts:842-853, benchmark/msm_timings.ts:45-63) stay correct; they only lengthen the
// serial stitch in k_fixup.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "msm_body.cuh"

namespace nmsm {

template <class Cv>
__global__ void k_prepare(const uint32_t* __restrict__ pts, uint32_t n, uint32_t* __restrict__ aff,
                          unsigned int* err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) prepare_body<Cv>(i, pts, aff, err);
}

template <class Cv, bool SCATTER>
__global__ void k_digits(const uint32_t* __restrict__ scalars, uint32_t n, MsmPlan plan,
                         unsigned int* __restrict__ counts_or_cursor, uint32_t* __restrict__ sorted,
                         unsigned int* err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) digits_body<Cv, SCATTER>(i, scalars, plan, counts_or_cursor, sorted, err);
}

// Exclusive scan of G counters by one block; offsets[G] = total.  Also copies offsets into cursor.
static __global__ void k_scan(const unsigned int* __restrict__ counts, uint32_t G, uint32_t* __restrict__ offsets,
                              unsigned int* __restrict__ cursor) {
  __shared__ uint32_t partial[1024];
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  const uint32_t per = (G + nt - 1) / nt;
  const uint32_t lo = min(tid * per, G), hi = min(lo + per, G);
  uint32_t sum = 0;
  for (uint32_t k = lo; k < hi; k++) sum += counts[k];
  partial[tid] = sum;
  __syncthreads();
  for (uint32_t d = 1; d < nt; d <<= 1) {  // Hillis-Steele inclusive scan over the per-thread sums
    uint32_t v = (tid >= d) ? partial[tid - d] : 0;
    __syncthreads();
    partial[tid] += v;
    __syncthreads();
  }
  uint32_t run = partial[tid] - sum;
  for (uint32_t k = lo; k < hi; k++) {
    offsets[k] = run;
    cursor[k] = run;
    run += counts[k];
  }
  if (tid == nt - 1) offsets[G] = partial[nt - 1];
}

template <class Cv>
__global__ void __launch_bounds__(128)
k_accumulate(const uint32_t* __restrict__ aff, const uint32_t* __restrict__ sorted,
             const uint32_t* __restrict__ offsets, MsmPlan plan, uint32_t* __restrict__ buckets,
             uint32_t* __restrict__ heads, uint32_t* __restrict__ tails) {
  accumulate_body<Cv>(blockIdx.x * blockDim.x + threadIdx.x, aff, sorted, offsets, plan, buckets, heads, tails);
}

template <class Cv>
__global__ void k_fixup(const uint32_t* __restrict__ offsets, MsmPlan plan, uint32_t* __restrict__ buckets,
                        const uint32_t* __restrict__ heads, const uint32_t* __restrict__ tails) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < (uint32_t)plan.G) fixup_body<Cv>(g, offsets, plan, buckets, heads, tails);
}

template <class Cv>
__global__ void __launch_bounds__(128)
k_reduce(const uint32_t* __restrict__ buckets, MsmPlan plan, uint32_t* __restrict__ chunk_out) {
  uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id < (uint32_t)plan.W * plan.chunks) reduce_body<Cv>(id, buckets, plan, chunk_out);
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

// Block w: window_out[w] = sum_k chunk_out[w][k]  (strided register partial sums, warp-shuffle
// tree, then one shared-memory hop between the warps)
template <class Cv>
__global__ void __launch_bounds__(128)
k_window_sum(const uint32_t* __restrict__ chunk_out, MsmPlan plan, uint32_t* __restrict__ window_out) {
  using G = typename Cv::G;
  extern __shared__ uint32_t smem[];
  const uint32_t w = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  typename G::Acc acc = G::identity();
  for (uint32_t k = tid; k < (uint32_t)plan.chunks; k += nt)
    nl_add<G>(acc, load_acc<G>(chunk_out + ((size_t)w * plan.chunks + k) * G::ACC_WORDS));
  for (int d = 16; d >= 1; d >>= 1) {
    typename G::Acc o = shfl_down_acc<G>(acc, d);
    nl_add<G>(acc, o);
  }
  const uint32_t lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  if (lane == 0) {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&acc);
    for (int k = 0; k < G::ACC_WORDS; k++) smem[warp * G::ACC_WORDS + k] = s[k];
  }
  __syncthreads();
  if (tid == 0) {
    for (uint32_t q = 1; q < nwarps; q++) {
      typename G::Acc o;
      uint32_t* d = reinterpret_cast<uint32_t*>(&o);
      for (int k = 0; k < G::ACC_WORDS; k++) d[k] = smem[q * G::ACC_WORDS + k];
      nl_add<G>(acc, o);
    }
    save_acc<G>(window_out + (size_t)w * G::ACC_WORDS, acc);
  }
}

template <class Cv, bool AFFINE_OUT>
__global__ void k_final(const uint32_t* __restrict__ window_out, MsmPlan plan, uint32_t* __restrict__ out,
                        uint32_t* __restrict__ out_inf) {
  if (blockIdx.x == 0 && threadIdx.x == 0) final_body<Cv, AFFINE_OUT>(window_out, plan, out, out_inf);
}

template <class Cv>
__global__ void k_fold(const uint32_t* __restrict__ accs, int count, uint32_t* __restrict__ out,
                       uint32_t* __restrict__ out_inf) {
  if (blockIdx.x == 0 && threadIdx.x == 0) fold_body<Cv>(accs, count, out, out_inf);
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
