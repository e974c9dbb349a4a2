// rp_algebra.h -- the prover's O(nm) polynomial algebra over Z_q on the host, 4 x 64-bit limbs.
//
// Replaces the Python big-int loops of _get_polynomial_coeffs / _final_compute
// (/root/reference/src/rangeproofs/rangeproof_prover.py:93-112, rangeproof_aggreg_prover.py:117-146) and the scalar lists
// of the P multiexp (:78-86): given the bit vector aL, the blinding vectors sL, sR and the challenges, it returns
//   phase 1 (after y, z):   t1 = sum sL_i (y^i (aR_i + z) + zz_i) + sum (aL_i - z) y^i sR_i,   t2 = sum sL_i y^i sR_i
//   phase 2 (after x):      l_i = aL_i - z + sL_i x,  r_i = y^i (aR_i + z + sR_i x) + zz_i,  t_hat = <l, r>,
//                           y^-i,  and the h-generator scalars z + zz_i y^-i of the P multiexp
// with aR = aL - 1 and zz_i = z^(2 + i / n) 2^(i mod n).  All values canonical (< q), 32-byte little-endian.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "fq.cuh"

namespace bp {
namespace rpa {

struct H4 { uint64_t v[4]; };
using fq64::Q64;
using fq64::u128;

inline H4 ld(const uint8_t* b) { H4 r; memcpy(r.v, b, 32); return r; }
inline void st(uint8_t* b, const H4& a) { memcpy(b, a.v, 32); }
inline bool geq_q(const H4& a) {
  for (int i = 3; i >= 0; i--) { if (a.v[i] > Q64[i]) return true; if (a.v[i] < Q64[i]) return false; }
  return true;
}
inline H4 reduce(H4 a) {              // any 256-bit value -> [0, q): 2^256 < 2q
  if (geq_q(a)) { u128 bw = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)a.v[i] - Q64[i] - (uint64_t)bw; a.v[i] = (uint64_t)d; bw = (d >> 64) & 1; } }
  return a;
}
inline H4 add(const H4& a, const H4& b) {
  H4 r; u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a.v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
  if (c || geq_q(r)) { u128 bw = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)r.v[i] - Q64[i] - (uint64_t)bw; r.v[i] = (uint64_t)d; bw = (d >> 64) & 1; } }
  return r;
}
inline H4 sub(const H4& a, const H4& b) {
  H4 r; u128 bw = 0;
  for (int i = 0; i < 4; i++) { u128 d = (u128)a.v[i] - b.v[i] - (uint64_t)bw; r.v[i] = (uint64_t)d; bw = (d >> 64) & 1; }
  if (bw) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r.v[i] + Q64[i]; r.v[i] = (uint64_t)c; c >>= 64; } }
  return r;
}
inline H4 mont(const H4& a, const H4& b) { H4 r; fq64::mont(r.v, a.v, b.v); return r; }
inline H4 r2() { Fq t = fq_const_r2(); H4 r; memcpy(r.v, t.v, 32); return r; }
inline H4 one() { H4 r = {{1, 0, 0, 0}}; return r; }
inline H4 zero() { H4 r = {{0, 0, 0, 0}}; return r; }
inline H4 to_m(const H4& a) { return mont(a, r2()); }                 // a * R
inline H4 from_m(const H4& a) { return mont(a, one()); }              // a * R^-1
// standard * Montgomery -> standard:  mont(a, bR) = a * b
inline H4 mul_sm(const H4& a, const H4& bm) { return mont(a, bm); }

struct Consts {                       // y^i (Montgomery), zz_i (standard), per position
  std::vector<H4> ym, zz;
};
inline Consts position_constants(const H4& y, const H4& z, size_t n, size_t m) {
  const size_t nm = n * m;
  Consts c; c.ym.resize(nm); c.zz.resize(nm);
  const H4 yM = to_m(y), zM = to_m(z);
  H4 acc = to_m(one());
  for (size_t i = 0; i < nm; i++) { c.ym[i] = acc; acc = mont(acc, yM); }        // y^i * R
  std::vector<H4> two(n);
  H4 t = one();
  for (size_t i = 0; i < n; i++) { two[i] = t; t = add(t, t); }                   // 2^i mod q
  H4 zj = from_m(mont(zM, zM));                                                   // z^2, standard
  for (size_t j = 0; j < m; j++) {
    const H4 zjM = to_m(zj);
    for (size_t i = 0; i < n; i++) c.zz[j * n + i] = mul_sm(two[i], zjM);
    zj = mul_sm(zj, zM);
  }
  return c;
}

}  // namespace rpa
}  // namespace bp
