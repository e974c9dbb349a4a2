// fq.cuh -- scalar field Z_q of secp256k1 (q = group order), host + device.
//
// Replaces: utils.ModP (+, -, *, neg, pow, inv) and inner_product
// (/root/reference/src/utils/utils.py:24-81,134-137), egcd (:15-21).
//
// Vectors (a, b, xs, ...) are stored in STANDARD form, 8 x 32-bit little-endian limbs, always
// reduced (< q).  Products use Montgomery REDC (CIOS, R = 2^256, -q^-1 mod 2^32 = 0x5588B13F):
//   mont(a, b)      = a*b*R^-1
//   mul(a, b)       = mont(mont(a, b), R^2)            (standard in, standard out)
//   mont(a, xR)     = a*x  when xR = x*R is the Montgomery form of a shared challenge
// The same code runs on the host (scalar preparation next to the Fiat-Shamir transcript) and in
// kernels (a/b folding, inner products, s-vector); it is O(n) work, not the IMAD-bound part.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define BP_HD __host__ __device__ __forceinline__
#else
#define BP_HD inline
#endif

namespace bp {

struct alignas(16) Fq { uint32_t v[8]; };

#define BP_Q_LIMBS {0xD0364141u, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}
#define BP_Q_NINV 0x5588B13Fu

BP_HD Fq fq_const_q() { Fq r = {BP_Q_LIMBS}; return r; }
BP_HD Fq fq_const_r() { Fq r = {{0x2FC9BEBFu, 0x402DA173u, 0x50B75FC4u, 0x45512319u, 1u, 0u, 0u, 0u}}; return r; }
BP_HD Fq fq_const_r2() { Fq r = {{0x67D7D140u, 0x896CF214u, 0x0E7CF878u, 0x741496C2u, 0x5BCD07C6u, 0xE697F5E4u, 0x81C69BC5u, 0x9D671CD5u}}; return r; }
BP_HD Fq fq_const_half() { Fq r = {{0x681B20A0u, 0xDFE92F46u, 0x57A4501Du, 0x5D576E73u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x7FFFFFFFu}}; return r; }
BP_HD Fq fq_zero() { Fq r = {{0, 0, 0, 0, 0, 0, 0, 0}}; return r; }
BP_HD Fq fq_one() { Fq r = {{1, 0, 0, 0, 0, 0, 0, 0}}; return r; }
BP_HD Fq fq_from_u64(uint64_t x) { Fq r = fq_zero(); r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32); return r; }

BP_HD bool fq_is_zero(const Fq& a) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.v[i]; return o == 0; }
BP_HD bool fq_eq(const Fq& a, const Fq& b) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i]; return o == 0; }
// a >= b as 256-bit integers
BP_HD bool fq_geq(const Fq& a, const Fq& b) {
  for (int i = 7; i >= 0; i--) { if (a.v[i] > b.v[i]) return true; if (a.v[i] < b.v[i]) return false; }
  return true;
}
BP_HD uint32_t fq_raw_add(Fq& r, const Fq& a, const Fq& b) {
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
  return (uint32_t)c;
}
BP_HD uint32_t fq_raw_sub(Fq& r, const Fq& a, const Fq& b) {
  int64_t c = 0;
  for (int i = 0; i < 8; i++) { c += (int64_t)a.v[i] - (int64_t)b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
  return (uint32_t)(c & 1);
}
// any 256-bit value -> [0, q): 2^256 < 2q so one conditional subtraction (pippenger.py:26)
BP_HD Fq fq_reduce(const Fq& a) { Fq q = fq_const_q(), r = a; if (fq_geq(a, q)) fq_raw_sub(r, a, q); return r; }
BP_HD Fq fq_add(const Fq& a, const Fq& b) {
  Fq r, q = fq_const_q(); uint32_t c = fq_raw_add(r, a, b);
  if (c || fq_geq(r, q)) { Fq t; fq_raw_sub(t, r, q); r = t; }
  return r;
}
BP_HD Fq fq_sub(const Fq& a, const Fq& b) {
  Fq r, q = fq_const_q(); uint32_t bw = fq_raw_sub(r, a, b);
  if (bw) { Fq t; fq_raw_add(t, r, q); r = t; }
  return r;
}
BP_HD Fq fq_neg(const Fq& a) { return fq_is_zero(a) ? a : fq_sub(fq_zero(), a); }
// a*b*R^-1 mod q, inputs < q
BP_HD Fq fq_mont(const Fq& a, const Fq& b) {
  const uint32_t qv[8] = BP_Q_LIMBS;
  uint32_t t[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 8; j++) { c += (uint64_t)a.v[j] * b.v[i] + t[j]; t[j] = (uint32_t)c; c >>= 32; }
    c += t[8]; t[8] = (uint32_t)c; t[9] = (uint32_t)(c >> 32);
    uint32_t m = t[0] * BP_Q_NINV;
    c = (uint64_t)m * qv[0] + t[0]; c >>= 32;
    for (int j = 1; j < 8; j++) { c += (uint64_t)m * qv[j] + t[j]; t[j - 1] = (uint32_t)c; c >>= 32; }
    c += t[8]; t[7] = (uint32_t)c; t[8] = t[9] + (uint32_t)(c >> 32);
  }
  Fq r; for (int i = 0; i < 8; i++) r.v[i] = t[i];
  Fq q = fq_const_q();
  if (t[8] || fq_geq(r, q)) { Fq s; fq_raw_sub(s, r, q); r = s; }
  return r;
}
BP_HD Fq fq_to_mont(const Fq& a) { return fq_mont(a, fq_const_r2()); }      // a*R
BP_HD Fq fq_from_mont(const Fq& a) { return fq_mont(a, fq_one()); }         // a*R^-1
BP_HD Fq fq_mul(const Fq& a, const Fq& b) { return fq_mont(fq_mont(a, b), fq_const_r2()); }
BP_HD Fq fq_sqr(const Fq& a) { return fq_mul(a, a); }
// a^e, e given as 8 little-endian limbs (standard form in/out)
BP_HD Fq fq_pow(const Fq& a, const Fq& e) {
  Fq am = fq_to_mont(a), acc = fq_const_r();   // 1 in Montgomery form
  for (int i = 255; i >= 0; i--) {
    acc = fq_mont(acc, acc);
    if ((e.v[i >> 5] >> (i & 31)) & 1) acc = fq_mont(acc, am);
  }
  return fq_from_mont(acc);
}
BP_HD Fq fq_pow_u64(const Fq& a, uint64_t e) { return fq_pow(a, fq_from_u64(e)); }
// a^-1 = a^(q-2); a == 0 has no inverse (caller checks: utils.py:69-70)
BP_HD Fq fq_inv(const Fq& a) {
  Fq e = fq_const_q(); e.v[0] -= 2;   // low limb 0xD0364141 - 2, no borrow
  return fq_pow(a, e);
}

// Host-only fast path for the per-round challenge inverse (the prover loop waits on it between launches):
// the same exponentiation on 4 x 64-bit limbs with 128-bit products, ~7x quicker than the portable 32-bit code.
namespace fq64 {
typedef unsigned __int128 u128;
static const uint64_t Q64[4] = {0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL};
static const uint64_t NINV64 = 0x4B0DFF665588B13FULL;          // -q^-1 mod 2^64
inline void mont(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * NINV64;
    c = (u128)m * Q64[0] + t[0]; c >>= 64;
    for (int j = 1; j < 4; j++) { c += (u128)m * Q64[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  bool ge = t[4] != 0;
  if (!ge) { ge = true; for (int i = 3; i >= 0; i--) { if (t[i] > Q64[i]) break; if (t[i] < Q64[i]) { ge = false; break; } } }
  if (ge) { u128 bw = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - Q64[i] - (uint64_t)bw; t[i] = (uint64_t)d; bw = (d >> 64) & 1; } }
  for (int i = 0; i < 4; i++) r[i] = t[i];
}
}  // namespace fq64
// a^-1 mod q on the host by the binary extended Euclidean algorithm over 4 x 64-bit limbs (~2-3 us; the 384-product
// Fermat chain it replaces took ~35 us and sat on the critical path of every IPA round).  a reduced, non-zero.
inline Fq fq_inv_host(const Fq& a) {
  using namespace fq64;
  auto is_one = [](const uint64_t* x) { return x[0] == 1 && !(x[1] | x[2] | x[3]); };
  auto shr1 = [](uint64_t* x, uint64_t top) { x[0] = x[0] >> 1 | x[1] << 63; x[1] = x[1] >> 1 | x[2] << 63; x[2] = x[2] >> 1 | x[3] << 63; x[3] = x[3] >> 1 | top << 63; };
  auto add = [](uint64_t* r, const uint64_t* y) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r[i] + y[i]; r[i] = (uint64_t)c; c >>= 64; } return (uint64_t)c; };
  auto sub = [](uint64_t* r, const uint64_t* y) { u128 bw = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)r[i] - y[i] - (uint64_t)bw; r[i] = (uint64_t)d; bw = (d >> 64) & 1; } return (uint64_t)bw; };
  auto geq = [](const uint64_t* x, const uint64_t* y) { for (int i = 3; i >= 0; i--) { if (x[i] != y[i]) return x[i] > y[i]; } return true; };
  auto half = [&](uint64_t* x) { uint64_t top = 0; if (x[0] & 1) top = add(x, Q64); shr1(x, top); };      // x / 2 mod q
  uint64_t u[4], v[4] = {Q64[0], Q64[1], Q64[2], Q64[3]}, x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; i++) u[i] = (uint64_t)a.v[2 * i] | (uint64_t)a.v[2 * i + 1] << 32;
  Fq r = fq_zero();
  if (!(u[0] | u[1] | u[2] | u[3])) return r;
  while (!is_one(u) && !is_one(v)) {
    while (!(u[0] & 1)) { shr1(u, 0); half(x1); }
    while (!(v[0] & 1)) { shr1(v, 0); half(x2); }
    if (geq(u, v)) { sub(u, v); if (sub(x1, x2)) add(x1, Q64); }
    else { sub(v, u); if (sub(x2, x1)) add(x2, Q64); }
  }
  const uint64_t* res = is_one(u) ? x1 : x2;
  for (int i = 0; i < 4; i++) { r.v[2 * i] = (uint32_t)res[i]; r.v[2 * i + 1] = (uint32_t)(res[i] >> 32); }
  return r;
}

}  // namespace bp
