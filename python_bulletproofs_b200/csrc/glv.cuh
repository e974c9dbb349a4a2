// glv.cuh -- secp256k1 endomorphism split of a scalar:  k = k1 + k2*lambda (mod q), |k1|, |k2| < 2^128,
// with phi(x, y) = (beta*x, y) = lambda*(x, y).  Used by the MSM so that the window recoding covers
// 129 bits instead of 256: half as many windows (bucket sets to reduce) and half as many doublings in
// the serial Horner tail, for the same number of bucket additions.  The result of the MSM is
// unchanged (same group element), so parity with Pippenger.multiexp (pippenger.py:22-61) is untouched.
//
// Lattice basis (a1, b1), (a2, b2) with a_i + b_i*lambda = 0 (mod q), b2 = a1, b1 < 0:
//   c1 = round(k*b2/q), c2 = round(-k*b1/q)   -- one 256x256-bit product + shift each, with g_i = round(2^384*b/q)
//   k1 = k - c1*a1 - c2*a2,  k2 = -c1*b1 - c2*b2 = c1*|b1| - c2*a1        (exact small integers)
// evaluated modulo 2^256 in two's complement: the sign is bit 255, no modular reduction is needed.
// Checked against Python big ints over 3*10^5 random scalars (max |k_i| and max c_i: 128 bits).
#pragma once
#include "fq.cuh"

namespace bp {

BP_HD Fq glv_g1() { Fq r = {{0x45DBB031u, 0xE893209Au, 0x71E8CA7Fu, 0x3DAA8A14u, 0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u}}; return r; }
BP_HD Fq glv_g2() { Fq r = {{0x8AC47F71u, 0x1571B4AEu, 0x9DF506C6u, 0x221208ACu, 0x0ABFE4C4u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u}}; return r; }
// beta (mod p), little-endian limbs
#define BP_BETA_LIMBS {0x719501EEu, 0xC1396C28u, 0x12F58995u, 0x9CF04975u, 0xAC3434E9u, 0x6E64479Eu, 0x657C0710u, 0x7AE96A2Bu}

// round(k * g / 2^384) as 5 limbs: bits 384.. of the 512-bit product plus the rounding bit 383
BP_HD void glv_mul_shift384(const Fq& k, const Fq& gq, uint32_t c[5]) {
  uint32_t t[16];
#if defined(__CUDA_ARCH__) && defined(BP_HAVE_FP_MUL_WIDE)
  mul_wide(t, k.v, gq.v);                        // the carry-chained IMAD.WIDE product of fp.cuh (a third of the instructions)
#else
  for (int i = 0; i < 16; i++) t[i] = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t cy = 0;
    for (int j = 0; j < 8; j++) { cy += (uint64_t)k.v[i] * gq.v[j] + t[i + j]; t[i + j] = (uint32_t)cy; cy >>= 32; }
    t[i + 8] = (uint32_t)cy;
  }
#endif
  uint64_t cy = (t[11] >> 31) & 1u;              // bit 383
  for (int i = 0; i < 4; i++) { cy += t[12 + i]; c[i] = (uint32_t)cy; cy >>= 32; }
  c[4] = (uint32_t)cy;
}

// acc (8 limbs, mod 2^256) += sign * a (na limbs) * b (nb limbs)
template <int NA, int NB>
BP_HD void glv_mac256(uint32_t acc[8], const uint32_t* a, const uint32_t* b, bool subtract) {
  uint32_t prod[8];
  for (int i = 0; i < 8; i++) prod[i] = 0;
  for (int i = 0; i < NA; i++) {
    uint64_t cy = 0;
    for (int j = 0; j < NB && i + j < 8; j++) { cy += (uint64_t)a[i] * b[j] + prod[i + j]; prod[i + j] = (uint32_t)cy; cy >>= 32; }
    if (i + NB < 8) prod[i + NB] = (uint32_t)cy;
  }
  if (subtract) {
    int64_t bw = 0;
    for (int i = 0; i < 8; i++) { bw += (int64_t)acc[i] - (int64_t)prod[i]; acc[i] = (uint32_t)bw; bw >>= 32; }
  } else {
    uint64_t cy = 0;
    for (int i = 0; i < 8; i++) { cy += (uint64_t)acc[i] + prod[i]; acc[i] = (uint32_t)cy; cy >>= 32; }
  }
}

// k (reduced mod q) -> magnitudes |k1|, |k2| (< 2^128, as Fq limbs) and their signs
BP_HD void glv_split(const Fq& k, Fq& m1, bool& neg1, Fq& m2, bool& neg2) {
  const uint32_t A1[4] = {0x9284EB15u, 0xE86C90E4u, 0xA7D46BCDu, 0x3086D221u};
  const uint32_t MB1[4] = {0x0ABFE4C3u, 0x6F547FA9u, 0x010E8828u, 0xE4437ED6u};
  const uint32_t A2[5] = {0x9D44CFD8u, 0x57C1108Du, 0xA8E2F3F6u, 0x14CA50F7u, 0x00000001u};
  uint32_t c1[5], c2[5];
  glv_mul_shift384(k, glv_g1(), c1);
  glv_mul_shift384(k, glv_g2(), c2);
  uint32_t k1[8], k2[8];
  for (int i = 0; i < 8; i++) { k1[i] = k.v[i]; k2[i] = 0; }
  glv_mac256<5, 4>(k1, c1, A1, true);     // k1 = k - c1*a1 - c2*a2
  glv_mac256<5, 5>(k1, c2, A2, true);
  glv_mac256<5, 4>(k2, c1, MB1, false);   // k2 = c1*|b1| - c2*a1
  glv_mac256<5, 4>(k2, c2, A1, true);
  neg1 = (k1[7] >> 31) != 0;
  neg2 = (k2[7] >> 31) != 0;
  uint64_t cy = 1;
  for (int i = 0; i < 8; i++) { uint32_t v = neg1 ? ~k1[i] : k1[i]; if (neg1) { cy += v; v = (uint32_t)cy; cy >>= 32; } m1.v[i] = v; }
  cy = 1;
  for (int i = 0; i < 8; i++) { uint32_t v = neg2 ? ~k2[i] : k2[i]; if (neg2) { cy += v; v = (uint32_t)cy; cy >>= 32; } m2.v[i] = v; }
}

}  // namespace bp
