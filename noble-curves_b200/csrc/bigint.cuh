// Carry-chain primitives for fixed-width big integers held in 32-bit register limbs.
//
// On the device every primitive is one PTX instruction (`mad.lo.cc.u32`, `madc.hi.cc.u32`,
// `add.cc.u32`, ...).  ptxas fuses a `mad(c).lo.cc` / `madc.hi.cc` pair that shares its
// multiplicands and targets adjacent accumulator limbs into a single `IMAD.WIDE.U32(.X)` with
// carry-in / carry-out, which is what makes the even/odd column layout in field.cuh reach
// one IMAD.WIDE per 32x32 partial product (2n^2+n per Montgomery multiplication).
//
// `asm volatile` keeps the relative order of the chain; NVVM itself never emits instructions
// that touch the PTX condition-code register, so nothing can clobber CC between two links.
//
// When compiled for the host (tests/hostemu only — never part of the product library) the same
// primitives are emulated with an explicit carry flag so the limb algorithms can be verified
// bit-for-bit on a machine without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define NMSM_HD __host__ __device__ __forceinline__
#define NMSM_D __device__ __forceinline__
#else
#define NMSM_HD inline
#define NMSM_D inline
#endif

namespace nmsm {

#if defined(__CUDA_ARCH__)

#define NMSM_OP3(name, ptx)                                                      \
  NMSM_D uint32_t name(uint32_t a, uint32_t b) {                                 \
    uint32_t r;                                                                  \
    asm volatile(ptx " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                 \
    return r;                                                                    \
  }
#define NMSM_OP4(name, ptx)                                                      \
  NMSM_D uint32_t name(uint32_t a, uint32_t b, uint32_t c) {                     \
    uint32_t r;                                                                  \
    asm volatile(ptx " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));     \
    return r;                                                                    \
  }

NMSM_OP3(add_cc, "add.cc.u32")
NMSM_OP3(addc_cc, "addc.cc.u32")
NMSM_OP3(addc, "addc.u32")
NMSM_OP3(sub_cc, "sub.cc.u32")
NMSM_OP3(subc_cc, "subc.cc.u32")
NMSM_OP3(subc, "subc.u32")
NMSM_OP4(mad_lo_cc, "mad.lo.cc.u32")
NMSM_OP4(madc_lo_cc, "madc.lo.cc.u32")
NMSM_OP4(mad_hi_cc, "mad.hi.cc.u32")
NMSM_OP4(madc_hi_cc, "madc.hi.cc.u32")
NMSM_OP4(madc_hi, "madc.hi.u32")
NMSM_OP4(madc_lo, "madc.lo.u32")

#undef NMSM_OP3
#undef NMSM_OP4

#else  // host emulation of the PTX condition code (test infrastructure)

namespace emu {
static uint32_t cc = 0;
}
inline uint32_t add_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b;
  emu::cc = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
inline uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b + emu::cc;
  emu::cc = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + emu::cc; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b;
  emu::cc = (uint32_t)(t >> 63);  // borrow
  return (uint32_t)t;
}
inline uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b - emu::cc;
  emu::cc = (uint32_t)(t >> 63);
  return (uint32_t)t;
}
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - emu::cc; }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c;
  emu::cc = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c + emu::cc;
  emu::cc = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (((uint64_t)a * b) >> 32) + c;
  emu::cc = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (((uint64_t)a * b) >> 32) + c + emu::cc;
  emu::cc = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
  return (uint32_t)(((uint64_t)a * b) >> 32) + c + emu::cc;
}
inline uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) {
  return (uint32_t)((uint64_t)a * b) + c + emu::cc;
}

#endif

}  // namespace nmsm
