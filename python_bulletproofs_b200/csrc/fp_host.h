// fp_host.h -- F_p (secp256k1 base field) on the HOST, 4 x 64-bit limbs: just enough to turn an XYZZ point into its canonical
// affine form.  The IPA prover's rounds end in the host's Fiat-Shamir hash anyway (inner_product_prover.py:102-106), and one
// modular inversion is ~2 us on a CPU core against ~45 us as a lone GPU thread at the end of a latency chain, so the round
// kernels hand L and R over in XYZZ coordinates and the host finishes them.  Same canonical (x, y) as ec.cuh:xyzz_to_affine.
#pragma once
#include <cstdint>
#include <cstring>

namespace bp {
namespace fph {
typedef unsigned __int128 u128;
static const uint64_t P64[4] = {0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL};
static const uint64_t PC = 0x1000003D1ULL;           // 2^256 = PC (mod p)

inline bool geq_p(const uint64_t* x) { for (int i = 3; i >= 0; i--) { if (x[i] != P64[i]) return x[i] > P64[i]; } return true; }
inline void canon(uint64_t* x) {                     // x < 2^256 -> [0, p)
  if (!geq_p(x)) return;
  u128 bw = 0;
  for (int i = 0; i < 4; i++) { u128 d = (u128)x[i] - P64[i] - (uint64_t)bw; x[i] = (uint64_t)d; bw = (d >> 64) & 1; }
}
inline bool is_zero(const uint64_t* x) { return !(x[0] | x[1] | x[2] | x[3]); }
// r = a * b mod p, canonical
inline void mul(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
  uint64_t t[8] = {0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a[i] * b[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
    t[i + 4] = (uint64_t)c;
  }
  // fold the high half: hi * 2^256 = hi * PC
  uint64_t s[5];
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)t[4 + i] * PC + t[i]; s[i] = (uint64_t)c; c >>= 64; }
  s[4] = (uint64_t)c;                                // < 2^34
  c = (u128)s[4] * PC;
  for (int i = 0; i < 4; i++) { c += s[i]; r[i] = (uint64_t)c; c >>= 64; }
  if ((uint64_t)c) {                                 // one more wrap (the sum is then tiny)
    c = PC;
    for (int i = 0; i < 4; i++) { c += r[i]; r[i] = (uint64_t)c; c >>= 64; }
  }
  canon(r);
}
// a^-1 mod p (a canonical, non-zero) by the binary extended Euclidean algorithm
inline void inv(uint64_t r[4], const uint64_t a[4]) {
  auto is_one = [](const uint64_t* x) { return x[0] == 1 && !(x[1] | x[2] | x[3]); };
  auto shr1 = [](uint64_t* x, uint64_t top) { x[0] = x[0] >> 1 | x[1] << 63; x[1] = x[1] >> 1 | x[2] << 63; x[2] = x[2] >> 1 | x[3] << 63; x[3] = x[3] >> 1 | top << 63; };
  auto add = [](uint64_t* x, const uint64_t* y) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)x[i] + y[i]; x[i] = (uint64_t)c; c >>= 64; } return (uint64_t)c; };
  auto sub = [](uint64_t* x, const uint64_t* y) { u128 bw = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)x[i] - y[i] - (uint64_t)bw; x[i] = (uint64_t)d; bw = (d >> 64) & 1; } return (uint64_t)bw; };
  auto geq = [](const uint64_t* x, const uint64_t* y) { for (int i = 3; i >= 0; i--) { if (x[i] != y[i]) return x[i] > y[i]; } return true; };
  auto half = [&](uint64_t* x) { uint64_t top = 0; if (x[0] & 1) top = add(x, P64); shr1(x, top); };
  uint64_t u[4] = {a[0], a[1], a[2], a[3]}, v[4] = {P64[0], P64[1], P64[2], P64[3]}, x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
  memset(r, 0, 32);
  if (is_zero(u)) return;
  while (!is_one(u) && !is_one(v)) {
    while (!(u[0] & 1)) { shr1(u, 0); half(x1); }
    while (!(v[0] & 1)) { shr1(v, 0); half(x2); }
    if (geq(u, v)) { sub(u, v); if (sub(x1, x2)) add(x1, P64); }
    else { sub(v, u); if (sub(x2, x1)) add(x2, P64); }
  }
  memcpy(r, is_one(u) ? x1 : x2, 32);
}
}  // namespace fph

// count XYZZ points (128 little-endian bytes each: X, Y, ZZ, ZZZ, lazy residues < 2^256) -> canonical affine (64 bytes each);
// the identity (ZZ = 0 mod p) becomes 64 zero bytes.  One shared inversion (Montgomery's trick over the ZZZ values).
inline void xyzz_to_affine_host(const uint8_t* xyzz, size_t count, uint8_t* out64) {
  using namespace fph;
  uint64_t pre[8][4], zzz[8][4];
  bool id[8];
  uint64_t acc[4] = {1, 0, 0, 0};
  for (size_t i = 0; i < count && i < 8; i++) {
    uint64_t zz[4];
    memcpy(zz, xyzz + 128 * i + 64, 32); canon(zz);
    memcpy(zzz[i], xyzz + 128 * i + 96, 32); canon(zzz[i]);
    id[i] = is_zero(zz) || is_zero(zzz[i]);
    memcpy(pre[i], acc, 32);
    if (!id[i]) mul(acc, acc, zzz[i]);
  }
  uint64_t inv_all[4];
  inv(inv_all, acc);
  for (size_t k = count > 8 ? 8 : count; k-- > 0;) {
    uint8_t* o = out64 + 64 * k;
    if (id[k]) { memset(o, 0, 64); continue; }
    uint64_t zi3[4], zi2[4], X[4], Y[4], ZZ[4], t[4];
    mul(zi3, inv_all, pre[k]);                       // ZZZ_k^-1
    mul(inv_all, inv_all, zzz[k]);
    memcpy(X, xyzz + 128 * k, 32); canon(X);
    memcpy(Y, xyzz + 128 * k + 32, 32); canon(Y);
    memcpy(ZZ, xyzz + 128 * k + 64, 32); canon(ZZ);
    mul(t, ZZ, ZZ); mul(zi2, zi3, zi3); mul(zi2, zi2, t);      // ZZ^-1 = ZZ^2 * ZZZ^-2   (ZZ^3 = ZZZ^2)
    mul(X, X, zi2); mul(Y, Y, zi3);
    memcpy(o, X, 32); memcpy(o + 32, Y, 32);
  }
}
}  // namespace bp
