// Quadratic extension Fp2 = Fp[u]/(u^2 + 1) for the G2 groups of bn254 and BLS12-381.
// Follows /root/reference/src/abstract/tower.ts:305-561 `_Field2`:
//   mul :420-431 (3 base multiplications), sqr :432-438 (2), add/sub/neg :393-418, inv :458-475.
#pragma once
#include "field.cuh"
#include "inv_divsteps.cuh"

namespace nmsm {

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
    Base t1 = a.c0 * b.c0;
    Base t2 = a.c1 * b.c1;
    Base o0 = t1 - t2;
    Base o1 = (a.c0 + a.c1) * (b.c0 + b.c1) - (t1 + t2);
    return Fp2{o0, o1};
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
