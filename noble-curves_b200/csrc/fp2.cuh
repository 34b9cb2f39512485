// Quadratic extension Fp2 = Fp[u]/(u^2 + 1) for the G2 groups of bn254 and BLS12-381.
// Follows /root/reference/src/abstract/tower.ts:305-561 `_Field2`:
//   mul :420-431 (3 base multiplications), sqr :432-438 (2), add/sub/neg :393-418, inv :458-475.
#pragma once
#include "field.cuh"
#include "inv_divsteps.cuh"

namespace nmsm {

#ifndef NMSM_FP2_LAZY
#define NMSM_FP2_LAZY 1
#endif
template <class C>
struct Fp2;
#if defined(__CUDACC__)
template <class C>
__device__ __noinline__ Fp2<C> fp2_mul_call(Fp2<C> a, Fp2<C> b);
#endif

template <class C>
struct Fp2 {
  using Base = Fp<C>;
  static constexpr int LIMBS = 2 * C::N;
  static constexpr int BASE_MULS = 3;
  static constexpr int BASE_SQRS = 2;
  using Params = C;
  Base c0, c1;

  NMSM_HD static Fp2 zero() { return Fp2{Base::zero(), Base::zero()}; }
  NMSM_HD static Fp2 one() { return Fp2{Base::one(), Base::zero()}; }
  NMSM_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  NMSM_HD bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
  NMSM_HD bool operator!=(const Fp2& o) const { return !(*this == o); }

  NMSM_HD friend Fp2 operator+(const Fp2& a, const Fp2& b) { return Fp2{a.c0 + b.c0, a.c1 + b.c1}; }
  NMSM_HD friend Fp2 operator-(const Fp2& a, const Fp2& b) { return Fp2{a.c0 - b.c0, a.c1 - b.c1}; }
  NMSM_HD Fp2 operator-() const { return Fp2{-c0, -c1}; }
  NMSM_HD friend Fp2 operator*(const Fp2& a, const Fp2& b) {
#if NMSM_FP2_LAZY
#if defined(__CUDA_ARCH__) && defined(NMSM_MUL_NOINLINE)
    return fp2_mul_call<C>(a, b);
#else
    return mul_lazy(a, b);
#endif
#else
    Base t1 = a.c0 * b.c0;
    Base t2 = a.c1 * b.c1;
    Base o0 = t1 - t2;
    Base o1 = (a.c0 + a.c1) * (b.c0 + b.c1) - (t1 + t2);
    return Fp2{o0, o1};
#endif
  }
  // Karatsuba with lazy reduction: three plain 2N-limb products, the combinations taken in double width, TWO Montgomery
  // reductions instead of three (and no modular corrections on the way): 5 N^2 + ... instead of 6 N^2 IMAD.WIDE.
  //   c0 = t1 - t2 + p R / 4   in (0, p R / 2)      t1 = a0 b0, t2 = a1 b1 < p^2 <= p R / 4
  //   c1 = t3 - t1 - t2 = a0 b1 + a1 b0 < 2 p^2      t3 = (a0 + a1)(b0 + b1), operand sums < 2p fit N limbs unreduced
  NMSM_HD static Fp2 mul_lazy(const Fp2& a, const Fp2& b) {
    constexpr int N = C::N;
    uint32_t t1[2 * N], t2[2 * N], t3[2 * N], sa[N], sb[N];
    mul_wide<C>(t1, a.c0.v, b.c0.v);
    mul_wide<C>(t2, a.c1.v, b.c1.v);
    sa[0] = add_cc(a.c0.v[0], a.c1.v[0]);
#pragma unroll
    for (int k = 1; k < N; k++) sa[k] = addc_cc(a.c0.v[k], a.c1.v[k]);
      sb[0] = add_cc(b.c0.v[0], b.c1.v[0]);
#pragma unroll
    for (int k = 1; k < N; k++) sb[k] = addc_cc(b.c0.v[k], b.c1.v[k]);
      mul_wide<C>(t3, sa, sb);
    // t3 -= t1; t3 -= t2
    t3[0] = sub_cc(t3[0], t1[0]);
#pragma unroll
    for (int k = 1; k < 2 * N; k++) t3[k] = subc_cc(t3[k], t1[k]);
    t3[0] = sub_cc(t3[0], t2[0]);
#pragma unroll
    for (int k = 1; k < 2 * N; k++) t3[k] = subc_cc(t3[k], t2[k]);
    // t1 += p R / 4; t1 -= t2
    t1[N - 1] = add_cc(t1[N - 1], p_times_quarter_r<C>(N - 1));
#pragma unroll
    for (int k = N; k < 2 * N; k++) t1[k] = addc_cc(t1[k], p_times_quarter_r<C>(k));
      t1[0] = sub_cc(t1[0], t2[0]);
#pragma unroll
    for (int k = 1; k < 2 * N; k++) t1[k] = subc_cc(t1[k], t2[k]);
    Fp2 r;
    mont_reduce_wide<C>(r.c0.v, t1);
    mont_reduce_wide<C>(r.c1.v, t3);
    return r;
  }
  // canonical layout: c0 limbs then c1 limbs (little-endian words)
  NMSM_HD static Fp2 from_canonical(const uint32_t* x) {
    return Fp2{Base::from_canonical(x), Base::from_canonical(x + C::N)};
  }
  NMSM_HD void to_canonical(uint32_t* x) const {
    c0.to_canonical(x);
    c1.to_canonical(x + C::N);
  }
  NMSM_HD static bool canonical_in_range(const uint32_t* x) {
    return Base::canonical_in_range(x) && Base::canonical_in_range(x + C::N);
  }
};

#if defined(__CUDACC__)
template <class C>
__device__ __noinline__ Fp2<C> fp2_mul_call(Fp2<C> a, Fp2<C> b) {
  return Fp2<C>::mul_lazy(a, b);
}
#endif

template <class C>
NMSM_HD Fp2<C> sqr(const Fp2<C>& a) {
  Fp<C> s = a.c0 + a.c1;
  Fp<C> d = a.c0 - a.c1;
  Fp<C> t = a.c0 + a.c0;
  return Fp2<C>{s * d, t * a.c1};
}
template <class C>
NMSM_HD Fp2<C> dbl(const Fp2<C>& a) {
  return a + a;
}
template <class C>
NMSM_HD Fp2<C> inv(const Fp2<C>& a) {
  Fp<C> factor = inv(sqr(a.c0) + sqr(a.c1));
  return Fp2<C>{factor * a.c0, factor * (-a.c1)};
}

}  // namespace nmsm
