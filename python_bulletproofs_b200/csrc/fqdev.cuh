// fqdev.cuh -- device-only fast multiplication in Z_q (q = order of secp256k1) in STANDARD form.
//
// Replaces utils.ModP.__mul__ / __pow__ / inv (/root/reference/src/utils/utils.py:39-44,57-58,66-72) where the verifier's
// scalar preparation runs on the GPU (verify.cuh).  fq.cuh's CIOS Montgomery product is one long dependent carry chain
// (~1 us per product for a lone warp); here the 8x8 product is fp.cuh's carry-chained IMAD.WIDE schoolbook (mul_wide, full
// instruction-level parallelism) and the reduction folds with q = 2^256 - c, c = 2^128 + CL (129 bits):
//   hi*2^256 + lo  =  lo + hi*CL + (hi << 128)   (mod q)        32 + 20 + 4 multiplications for the three folds
// 120 multiplications per product against 136, no Montgomery domain (no conversions at either end), canonical output.
#pragma once
#include "fp.cuh"
#include "fq.cuh"

namespace bp {

// r (nout limbs) = lo (8 limbs) + hi (NH limbs) * c,  c = 2^128 + CL;  per-limb 64-bit accumulators, one carry pass
template <int NH, int NOUT>
BP_DI void fq_fold_step(u32* r, const u32* lo, const u32* hi) {
  const u32 CL[4] = {0x2FC9BEBFu, 0x402DA173u, 0x50B75FC4u, 0x45512319u};
  u64 acc[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; k++) acc[k] = k < 8 ? lo[k] : 0;
#pragma unroll
  for (int i = 0; i < NH; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const u64 p = (u64)hi[i] * CL[j];
      acc[i + j] += (u32)p;
      if (i + j + 1 < NOUT) acc[i + j + 1] += p >> 32;
    }
    if (i + 4 < NOUT) acc[i + 4] += hi[i];
  }
  u64 c = 0;
#pragma unroll
  for (int k = 0; k < NOUT; k++) { c += acc[k]; r[k] = (u32)c; c >>= 32; }
}

// t (16 limbs, any 512-bit value) mod q, canonical
BP_DI Fq fq_fold512(const u32 t[16]) {
  u32 s[13], r[9], w[9];
  fq_fold_step<8, 13>(s, t, t + 8);          // < 2^256 + 2^385
  fq_fold_step<5, 9>(r, s, s + 8);           // hi < 2^129:  < 2^256 + 2^258.4  ->  r[8] <= 6
  fq_fold_step<1, 9>(w, r, r + 8);           // < 2^256 + 2^132: w[8] in {0, 1}
  Fq x;
#pragma unroll
  for (int i = 0; i < 8; i++) x.v[i] = w[i];
  if (w[8]) {                                // value = 2^256 + x with x < 2^133: subtract q, i.e. add c (no further carry)
    const Fq c = fq_const_r();
    Fq y; fq_raw_add(y, x, c); x = y;
  }
  return fq_reduce(x);
}

BP_DI Fq fq_mul_dev(const Fq& a, const Fq& b) {
  u32 t[16];
  mul_wide(t, a.v, b.v);
  return fq_fold512(t);
}
BP_DI Fq fq_sqr_dev(const Fq& a) {
  u32 t[16];
  sqr_wide(t, a.v);
  return fq_fold512(t);
}
__device__ __noinline__ Fq fq_mul_ni(const Fq& a, const Fq& b) { return fq_mul_dev(a, b); }

// a^-1 = a^(q-2), 4-bit fixed windows (256 squarings + 64 multiplications + 14 for the table); 0 -> 0
__device__ __noinline__ Fq fq_inv_dev(const Fq& a) {
  Fq tab[16];
  tab[0] = fq_one(); tab[1] = a;
#pragma unroll 1
  for (int i = 2; i < 16; i++) tab[i] = fq_mul_ni(tab[i - 1], a);
  Fq e = fq_const_q(); e.v[0] -= 2;
  Fq acc = fq_one();
#pragma unroll 1
  for (int w = 63; w >= 0; w--) {
#pragma unroll 1
    for (int s = 0; s < 4; s++) acc = fq_mul_ni(acc, acc);
    const u32 d = (e.v[w >> 3] >> (4 * (w & 7))) & 15u;
    if (d) acc = fq_mul_ni(acc, tab[d]);
  }
  return acc;
}

// a^-1 mod q by the binary extended Euclidean algorithm (shifts, adds, subtractions; data-dependent trip counts): for a LONE
// lane this is several times quicker than the 334-product Fermat chain; used where one lane per warp inverts on behalf of
// the others (k_rp_invert).  Input reduced (< q); 0 -> 0.  Invariants: x1*a = u, x2*a = v (mod q).
__device__ __noinline__ Fq fq_inv_gcd(const Fq& a) {
  const u32 Q_[8] = BP_Q_LIMBS;
  u32 u[8], v[8], x1[8], x2[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { u[i] = a.v[i]; v[i] = Q_[i]; x1[i] = i == 0 ? 1u : 0u; x2[i] = 0u; }
  if (fq_is_zero(a)) return fq_zero();
  for (;;) {
    u32 ou = u[0] ^ 1u, ov = v[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < 8; i++) { ou |= u[i]; ov |= v[i]; }
    if (ou == 0u || ov == 0u) {                 // u == 1 or v == 1
      Fq r;
#pragma unroll
      for (int i = 0; i < 8; i++) r.v[i] = ou == 0u ? x1[i] : x2[i];
      return r;
    }
    while (!(u[0] & 1u)) { shr1_256(u, 0); u32 top = 0; if (x1[0] & 1u) top = add256(x1, x1, Q_); shr1_256(x1, top); }
    while (!(v[0] & 1u)) { shr1_256(v, 0); u32 top = 0; if (x2[0] & 1u) top = add256(x2, x2, Q_); shr1_256(x2, top); }
    u32 d[8];
    if (sub256(d, u, v) == 0u) {                // u >= v
#pragma unroll
      for (int i = 0; i < 8; i++) u[i] = d[i];
      if (sub256(x1, x1, x2)) add256(x1, x1, Q_);
    } else {
      sub256(v, v, u);
      if (sub256(x2, x2, x1)) add256(x2, x2, Q_);
    }
  }
}

}  // namespace bp
