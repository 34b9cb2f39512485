// Batched point decoding kernels; one thread per encoding.
#include "context.h"
#include "codec.cuh"

namespace nmsm {

template <int CURVE>
__global__ void __launch_bounds__(128)
k_decode(const uint8_t* __restrict__ enc, uint32_t n, int flags, uint32_t* __restrict__ out_xy, uint8_t* __restrict__ status) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int WORDS = CURVE == NMSM_BLS12_381_G2 ? 48 : (CURVE == NMSM_BLS12_381_G1 ? 24 : 16);
  uint32_t* o = out_xy + (size_t)i * WORDS;
  int st = 0;
  if (CURVE == NMSM_SECP256K1) st = sec1_decode_secp256k1(enc + (size_t)i * 33, o);
  if (CURVE == NMSM_BLS12_381_G1) st = zcash_decode_bls12_381_g1(enc + (size_t)i * 48, o);
  if (CURVE == NMSM_BLS12_381_G2) st = zcash_decode_bls12_381_g2(enc + (size_t)i * 96, o);
  if (CURVE == NMSM_ED25519) st = ed25519_decode(enc + (size_t)i * 32, o, (flags & NMSM_DECODE_ZIP215) != 0);
  if (st == 0)  // an invalid encoding hands back zeros, never stale device memory
    for (int k = 0; k < WORDS; k++) o[k] = 0;
  status[i] = (uint8_t)st;
}

int decode_points_impl(int curve, const uint8_t* enc, uint64_t n, int flags, uint8_t* out_xy, uint8_t* out_status) {
  Context& X = g_ctx;
  Slot& C = X.slot[0];
  int enc_bytes, pt_bytes;
  switch (curve) {
    case NMSM_SECP256K1: enc_bytes = 33; pt_bytes = 64; break;
    case NMSM_BLS12_381_G1: enc_bytes = 48; pt_bytes = 96; break;
    case NMSM_BLS12_381_G2: enc_bytes = 96; pt_bytes = 192; break;
    case NMSM_ED25519: enc_bytes = 32; pt_bytes = 64; break;
    default: return fail(NMSM_ERR_ARG, "nmsm_points_decode: no decoder for this curve (secp256k1, ed25519, BLS12-381 G1/G2 only)");
  }
  if (n == 0) return NMSM_OK;
  if (n >= (1ull << 31)) return fail(NMSM_ERR_ARG, "n must be < 2^31");
  CK(C.in_pts.ensure(n * enc_bytes + 64));
  CK(C.mul_out.ensure(n * (pt_bytes + 1) + 64));
  uint32_t* d_xy = (uint32_t*)C.mul_out.p;
  uint8_t* d_st = (uint8_t*)C.mul_out.p + n * pt_bytes;
  cudaStream_t st = C.stream;
  CK(cudaMemcpyAsync(C.in_pts.p, enc, n * enc_bytes, cudaMemcpyHostToDevice, st));
  const unsigned int blocks = (unsigned int)((n + 127) / 128);
  if (curve == NMSM_SECP256K1) k_decode<NMSM_SECP256K1><<<blocks, 128, 0, st>>>((const uint8_t*)C.in_pts.p, (uint32_t)n, flags, d_xy, d_st);
  if (curve == NMSM_BLS12_381_G1) k_decode<NMSM_BLS12_381_G1><<<blocks, 128, 0, st>>>((const uint8_t*)C.in_pts.p, (uint32_t)n, flags, d_xy, d_st);
  if (curve == NMSM_BLS12_381_G2) k_decode<NMSM_BLS12_381_G2><<<blocks, 128, 0, st>>>((const uint8_t*)C.in_pts.p, (uint32_t)n, flags, d_xy, d_st);
  if (curve == NMSM_ED25519) k_decode<NMSM_ED25519><<<blocks, 128, 0, st>>>((const uint8_t*)C.in_pts.p, (uint32_t)n, flags, d_xy, d_st);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out_xy, d_xy, n * pt_bytes, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(out_status, d_st, n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return NMSM_OK;
}

}  // namespace nmsm
