/*
 * oracle/ecc_oracle.c -- CPU restatement of the elliptic-curve work on the hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in python_bulletproofs_b200/ may link, load or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, as the checker / CPU baseline.
 *
 * What it restates (all citations relative to /root/reference/):
 *   - the group law the reference obtains from the third-party `fastecdsa` package
 *     (Point.__add__, Point.__mul__; call sites src/pippenger/group.py:31-32,
 *     src/utils/utils.py:43-44).  fastecdsa is unpinned (CI builds GitHub master,
 *     .travis.yml:12-14) and absent from this image; the secp256k1 group law on canonical
 *     affine coordinates is a public standard (SEC 2 v2.0 s.2.4.1) with unique outputs.
 *   - orc_msm_naive / orc_msm_bucket: RESULT semantics of Pippenger.multiexp
 *     (src/pippenger/pippenger.py:22-29,56-61): scalars taken mod q, empty input => identity,
 *     output = sum e_i * g_i as a canonical affine point.
 *   - orc_msm_subset: the reference's actual ALGORITHM (Pippenger's subset-table method,
 *     src/pippenger/pippenger.py:31-94) restated with bitmask-indexed tables, so the CPU
 *     baseline can time "the reference's algorithm in C" where it is still feasible.
 *   - orc_fold: one generator-folding step g'_i = x_lo*lo_i + x_hi*hi_i
 *     (src/innerproduct/inner_product_prover.py:107-108).
 *   - orc_scalar_mul_batch: n independent scalar multiplications, e.g. hsp[i] = y^-i * hs[i]
 *     (src/rangeproofs/rangeproof_prover.py:77, rangeproof_verifier.py:72).
 *
 * Parity pinning: tests/test_oracle.py checks this file against (a) published secp256k1
 * multiples of G, (b) Python big-int arithmetic, (c) tests/golden/ fixtures produced by the
 * UNMODIFIED reference run in the build container (oracle/gen_golden.py).
 *
 * Data format (same as include/bp_gpu.h): affine point = 64 bytes, x then y, each 32-byte
 * little-endian; identity = 64 zero bytes; scalar = 32-byte little-endian, any value < 2^256
 * (reduced mod q here).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } fe;           /* field element, canonical in [0,p) */
typedef struct { fe X, Y, Z; int inf; } jac;    /* Jacobian; inf!=0 => identity */

static const fe FE_P = {{0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL}};
static const uint64_t Q_LIMBS[4] = {0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL};
#define P_C 0x1000003D1ULL /* 2^256 - p */

static int fe_is_zero(const fe *a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static int fe_eq(const fe *a, const fe *b) { return memcmp(a, b, sizeof(fe)) == 0; }
static int fe_geq_p(const fe *a) {
  for (int i = 3; i >= 0; i--) { if (a->v[i] > FE_P.v[i]) return 1; if (a->v[i] < FE_P.v[i]) return 0; }
  return 1;
}
static void fe_sub_p(fe *a) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)a->v[i] - FE_P.v[i] - (uint64_t)b; a->v[i] = (uint64_t)t; b = (t >> 64) & 1; }
}
static void fe_add(fe *r, const fe *a, const fe *b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a->v[i] + b->v[i]; r->v[i] = (uint64_t)c; c >>= 64; }
  if (c) { /* 2^256 = P_C (mod p) */
    u128 t = (u128)r->v[0] + P_C; r->v[0] = (uint64_t)t; t >>= 64;
    for (int i = 1; i < 4 && t; i++) { t += r->v[i]; r->v[i] = (uint64_t)t; t >>= 64; }
  }
  if (fe_geq_p(r)) fe_sub_p(r);
}
static void fe_neg(fe *r, const fe *a) {
  if (fe_is_zero(a)) { *r = *a; return; }
  u128 b = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)FE_P.v[i] - a->v[i] - (uint64_t)b; r->v[i] = (uint64_t)t; b = (t >> 64) & 1; }
}
static void fe_sub(fe *r, const fe *a, const fe *b) { fe nb; fe_neg(&nb, b); fe_add(r, a, &nb); }
static void fe_mul(fe *r, const fe *a, const fe *b) {
  uint64_t t[8] = {0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a->v[i] * b->v[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
    t[i + 4] = (uint64_t)c;
  }
  /* fold: t = lo + hi * P_C   (hi*P_C < 2^(256+33)) */
  uint64_t s[5]; u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)t[4 + i] * P_C + t[i]; s[i] = (uint64_t)c; c >>= 64; }
  s[4] = (uint64_t)c;
  /* second fold of s[4] (< 2^34) */
  c = (u128)s[4] * P_C + s[0]; r->v[0] = (uint64_t)c; c >>= 64;
  for (int i = 1; i < 4; i++) { c += s[i]; r->v[i] = (uint64_t)c; c >>= 64; }
  if (c) { /* wrapped once more: remainder is tiny, adding P_C cannot carry out */
    u128 d = (u128)r->v[0] + P_C; r->v[0] = (uint64_t)d; d >>= 64;
    for (int i = 1; i < 4 && d; i++) { d += r->v[i]; r->v[i] = (uint64_t)d; d >>= 64; }
  }
  if (fe_geq_p(r)) fe_sub_p(r);
}
static void fe_sqr(fe *r, const fe *a) { fe_mul(r, a, a); }
static void fe_inv(fe *r, const fe *a) { /* a^(p-2), square-and-multiply MSB first */
  fe e = FE_P; e.v[0] -= 2;
  fe acc = {{1, 0, 0, 0}};
  for (int i = 255; i >= 0; i--) {
    fe_sqr(&acc, &acc);
    if ((e.v[i >> 6] >> (i & 63)) & 1) fe_mul(&acc, &acc, a);
  }
  *r = acc;
}

/* ---- byte codecs ---- */
static void fe_from_le(fe *r, const uint8_t *b) { memcpy(r->v, b, 32); }
static void fe_to_le(uint8_t *b, const fe *a) { memcpy(b, a->v, 32); }
static int bytes_all_zero(const uint8_t *b, size_t n) { for (size_t i = 0; i < n; i++) if (b[i]) return 0; return 1; }
static void jac_from_affine_bytes(jac *r, const uint8_t *b) {
  if (bytes_all_zero(b, 64)) { memset(r, 0, sizeof(*r)); r->inf = 1; return; }
  fe_from_le(&r->X, b); fe_from_le(&r->Y, b + 32);
  r->Z = (fe){{1, 0, 0, 0}}; r->inf = 0;
}
static void jac_to_affine_bytes(uint8_t *b, const jac *a) {
  if (a->inf) { memset(b, 0, 64); return; }
  fe zi, zi2, zi3, x, y;
  fe_inv(&zi, &a->Z); fe_sqr(&zi2, &zi); fe_mul(&zi3, &zi2, &zi);
  fe_mul(&x, &a->X, &zi2); fe_mul(&y, &a->Y, &zi3);
  fe_to_le(b, &x); fe_to_le(b + 32, &y);
}

/* ---- group law (a = 0, b = 7) ---- */
static void jac_dbl(jac *r, const jac *a) {
  if (a->inf || fe_is_zero(&a->Y)) { memset(r, 0, sizeof(*r)); r->inf = 1; return; }
  fe A, B, C, D, E, F, t, X3, Y3, Z3;
  fe_sqr(&A, &a->X); fe_sqr(&B, &a->Y); fe_sqr(&C, &B);
  fe_add(&t, &a->X, &B); fe_sqr(&t, &t); fe_sub(&t, &t, &A); fe_sub(&t, &t, &C); fe_add(&D, &t, &t);
  fe_add(&E, &A, &A); fe_add(&E, &E, &A);
  fe_sqr(&F, &E);
  fe_sub(&X3, &F, &D); fe_sub(&X3, &X3, &D);
  fe_sub(&t, &D, &X3); fe_mul(&Y3, &E, &t);
  fe_add(&C, &C, &C); fe_add(&C, &C, &C); fe_add(&C, &C, &C); fe_sub(&Y3, &Y3, &C);
  fe_mul(&Z3, &a->Y, &a->Z); fe_add(&Z3, &Z3, &Z3);
  r->X = X3; r->Y = Y3; r->Z = Z3; r->inf = 0;
}
static void jac_add(jac *r, const jac *a, const jac *b) {
  if (a->inf) { *r = *b; return; }
  if (b->inf) { *r = *a; return; }
  fe z1z1, z2z2, u1, u2, s1, s2, h, rr, t;
  fe_sqr(&z1z1, &a->Z); fe_sqr(&z2z2, &b->Z);
  fe_mul(&u1, &a->X, &z2z2); fe_mul(&u2, &b->X, &z1z1);
  fe_mul(&t, &b->Z, &z2z2); fe_mul(&s1, &a->Y, &t);
  fe_mul(&t, &a->Z, &z1z1); fe_mul(&s2, &b->Y, &t);
  fe_sub(&h, &u2, &u1); fe_sub(&rr, &s2, &s1);
  if (fe_is_zero(&h)) {
    if (fe_is_zero(&rr)) { jac_dbl(r, a); return; }
    memset(r, 0, sizeof(*r)); r->inf = 1; return;
  }
  fe h2, h3, v, X3, Y3, Z3;
  fe_sqr(&h2, &h); fe_mul(&h3, &h2, &h); fe_mul(&v, &u1, &h2);
  fe_sqr(&X3, &rr); fe_sub(&X3, &X3, &h3); fe_sub(&X3, &X3, &v); fe_sub(&X3, &X3, &v);
  fe_sub(&t, &v, &X3); fe_mul(&Y3, &rr, &t); fe_mul(&t, &s1, &h3); fe_sub(&Y3, &Y3, &t);
  fe_mul(&Z3, &a->Z, &b->Z); fe_mul(&Z3, &Z3, &h);
  r->X = X3; r->Y = Y3; r->Z = Z3; r->inf = 0;
}
static void jac_neg(jac *r, const jac *a) { *r = *a; if (!a->inf) fe_neg(&r->Y, &a->Y); }

/* ---- scalars ---- */
static void scalar_from_le_mod_q(uint64_t k[4], const uint8_t *b) {
  memcpy(k, b, 32);
  /* k < 2^256 < 2q, so one conditional subtraction reduces (pippenger.py:26) */
  int ge = 1;
  for (int i = 3; i >= 0; i--) { if (k[i] > Q_LIMBS[i]) { ge = 1; break; } if (k[i] < Q_LIMBS[i]) { ge = 0; break; } }
  if (ge) { u128 bw = 0; for (int i = 0; i < 4; i++) { u128 t = (u128)k[i] - Q_LIMBS[i] - (uint64_t)bw; k[i] = (uint64_t)t; bw = (t >> 64) & 1; } }
}
static int scalar_bit(const uint64_t k[4], int i) { return (int)((k[i >> 6] >> (i & 63)) & 1); }
static unsigned scalar_bits(const uint64_t k[4], int lo, int n) { /* n <= 24 bits starting at lo */
  unsigned r = 0;
  for (int i = 0; i < n; i++) { int b = lo + i; if (b < 256) r |= (unsigned)scalar_bit(k, b) << i; }
  return r;
}
static void jac_scalar_mul(jac *r, const jac *p, const uint64_t k[4]) {
  jac acc; memset(&acc, 0, sizeof(acc)); acc.inf = 1;
  for (int i = 255; i >= 0; i--) { jac_dbl(&acc, &acc); if (scalar_bit(k, i)) jac_add(&acc, &acc, p); }
  *r = acc;
}

/* ================= exported ================= */

int orc_point_add(const uint8_t *a64, const uint8_t *b64, uint8_t *out64) {
  jac a, b, r; jac_from_affine_bytes(&a, a64); jac_from_affine_bytes(&b, b64);
  jac_add(&r, &a, &b); jac_to_affine_bytes(out64, &r); return 0;
}
int orc_on_curve(const uint8_t *a64) {
  if (bytes_all_zero(a64, 64)) return 1;
  fe x, y, l, r, seven = {{7, 0, 0, 0}};
  fe_from_le(&x, a64); fe_from_le(&y, a64 + 32);
  if (fe_geq_p(&x) || fe_geq_p(&y)) return 0;
  fe_sqr(&l, &y); fe_sqr(&r, &x); fe_mul(&r, &r, &x); fe_add(&r, &r, &seven);
  return fe_eq(&l, &r);
}
int orc_scalar_mul_batch(const uint8_t *pts, const uint8_t *sc, size_t n, uint8_t *out) {
#pragma omp parallel for schedule(dynamic, 16)
  for (long i = 0; i < (long)n; i++) {
    jac p, r; uint64_t k[4];
    jac_from_affine_bytes(&p, pts + 64 * i); scalar_from_le_mod_q(k, sc + 32 * i);
    jac_scalar_mul(&r, &p, k); jac_to_affine_bytes(out + 64 * i, &r);
  }
  return 0;
}
/* out_i = x_lo * lo_i + x_hi * hi_i   (inner_product_prover.py:107-108) */
int orc_fold(const uint8_t *lo, const uint8_t *hi, size_t n, const uint8_t *x_lo, const uint8_t *x_hi, uint8_t *out) {
  uint64_t kl[4], kh[4]; scalar_from_le_mod_q(kl, x_lo); scalar_from_le_mod_q(kh, x_hi);
#pragma omp parallel for schedule(dynamic, 16)
  for (long i = 0; i < (long)n; i++) {
    jac a, b, ra, rb, r;
    jac_from_affine_bytes(&a, lo + 64 * i); jac_from_affine_bytes(&b, hi + 64 * i);
    jac_scalar_mul(&ra, &a, kl); jac_scalar_mul(&rb, &b, kh); jac_add(&r, &ra, &rb);
    jac_to_affine_bytes(out + 64 * i, &r);
  }
  return 0;
}
/* sum e_i * g_i, one double-and-add per term: the obviously-correct form. */
int orc_msm_naive(const uint8_t *pts, const uint8_t *sc, size_t n, uint8_t *out64) {
  jac acc; memset(&acc, 0, sizeof(acc)); acc.inf = 1;
  for (size_t i = 0; i < n; i++) {
    jac p, r; uint64_t k[4];
    jac_from_affine_bytes(&p, pts + 64 * i); scalar_from_le_mod_q(k, sc + 32 * i);
    jac_scalar_mul(&r, &p, k); jac_add(&acc, &acc, &r);
  }
  jac_to_affine_bytes(out64, &acc); return 0;
}
/* bucket method on one slice (unsigned c-bit windows); used for the multi-threaded CPU baseline. */
static void msm_bucket_slice(jac *res, const uint8_t *pts, const uint8_t *sc, size_t n, int c) {
  int W = (256 + c - 1) / c; size_t nb = ((size_t)1 << c) - 1;
  jac *bk = (jac *)malloc(sizeof(jac) * nb);
  jac *P = (jac *)malloc(sizeof(jac) * n);
  uint64_t (*K)[4] = malloc(sizeof(uint64_t[4]) * n);
  for (size_t i = 0; i < n; i++) { jac_from_affine_bytes(&P[i], pts + 64 * i); scalar_from_le_mod_q(K[i], sc + 32 * i); }
  jac total; memset(&total, 0, sizeof(total)); total.inf = 1;
  for (int w = W - 1; w >= 0; w--) {
    for (int d = 0; d < c; d++) jac_dbl(&total, &total);
    for (size_t b = 0; b < nb; b++) { memset(&bk[b], 0, sizeof(jac)); bk[b].inf = 1; }
    for (size_t i = 0; i < n; i++) { unsigned d = scalar_bits(K[i], w * c, c); if (d) jac_add(&bk[d - 1], &bk[d - 1], &P[i]); }
    jac run, sum; memset(&run, 0, sizeof(run)); run.inf = 1; sum = run;
    for (size_t b = nb; b-- > 0;) { jac_add(&run, &run, &bk[b]); jac_add(&sum, &sum, &run); }
    jac_add(&total, &total, &sum);
  }
  free(bk); free(P); free(K); *res = total;
}
int orc_msm_bucket(const uint8_t *pts, const uint8_t *sc, size_t n, uint8_t *out64, int nthreads) {
  if (n == 0) { memset(out64, 0, 64); return 0; }
  int T = nthreads > 0 ? nthreads : 1;
  if ((size_t)T > n) T = (int)n;
  size_t per = (n + T - 1) / T;
  int c = 1; { double best = 1e300; for (int cc = 1; cc <= 20; cc++) { double cost = ceil(256.0 / cc) * ((double)per + 2.0 * ((1u << cc) - 1) + cc); if (cost < best) { best = cost; c = cc; } } }
  jac *part = (jac *)malloc(sizeof(jac) * T);
#pragma omp parallel for num_threads(T) schedule(static, 1)
  for (int t = 0; t < T; t++) {
    size_t lo = per * t, hi = lo + per; if (hi > n) hi = n;
    if (lo >= hi) { memset(&part[t], 0, sizeof(jac)); part[t].inf = 1; continue; }
    msm_bucket_slice(&part[t], pts + 64 * lo, sc + 32 * lo, hi - lo, c);
  }
  jac acc; memset(&acc, 0, sizeof(acc)); acc.inf = 1;
  for (int t = 0; t < T; t++) jac_add(&acc, &acc, &part[t]);
  free(part); jac_to_affine_bytes(out64, &acc); return 0;
}
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- the reference's own algorithm (pippenger.py:31-94), restated ---- */
static uint64_t isqrt_u64(uint64_t x) { uint64_t r = (uint64_t)sqrt((double)x); while (r * r > x) r--; while ((r + 1) * (r + 1) <= x) r++; return r; }
int orc_msm_subset(const uint8_t *pts, const uint8_t *sc, size_t N, uint8_t *out64) {
  if (N == 0) { memset(out64, 0, 64); return 0; }                       /* pippenger.py:28-29 */
  const uint64_t lamb = 256;
  uint64_t s = isqrt_u64(lamb / N) + 1, t = isqrt_u64(lamb * N) + 1;     /* :33-34 */
  size_t M = N * s;
  jac *gs = (jac *)malloc(sizeof(jac) * M);                              /* gs_bin flattened, :35-40,52 */
  uint64_t (*K)[4] = malloc(sizeof(uint64_t[4]) * N);
  for (size_t i = 0; i < N; i++) {
    scalar_from_le_mod_q(K[i], sc + 32 * i);                             /* :26 */
    jac_from_affine_bytes(&gs[i * s], pts + 64 * i);
    for (uint64_t j = 1; j < s; j++) jac_dbl(&gs[i * s + j], &gs[i * s + j - 1]);
  }
  /* element (i,j) has t-bit row e[k] = bit (j + s*k) of es[i]            :41-49 */
  int b = (M > 1) ? (int)floor(log2((double)M) - log2(log2((double)M))) : 0;   /* :66 */
  if (M == 2) b = 1; /* log2(2)-log2(log2(2)) = 1 - 0 = 1 */
  if (b <= 0) b = 1;                                                     /* :67 */
  size_t ntab = (M + b - 1) / b;                                         /* :68 */
  size_t tsz = (size_t)1 << b;
  jac *T = (jac *)malloc(sizeof(jac) * ntab * tsz);                      /* all-subset tables :69-81 */
  for (size_t q = 0; q < ntab; q++) {
    jac *tab = T + q * tsz; size_t base = q * b; size_t cnt = (base + b <= M) ? (size_t)b : M - base;
    memset(&tab[0], 0, sizeof(jac)); tab[0].inf = 1;
    for (size_t m = 1; m < ((size_t)1 << cnt); m++) {
      int top = 63 - __builtin_clzll((unsigned long long)m);
      size_t rest = m & ~((size_t)1 << top);
      if (rest == 0) tab[m] = gs[base + top]; else jac_add(&tab[m], &tab[rest], &gs[base + top]);   /* :75-79 */
    }
  }
  jac *Gs = (jac *)malloc(sizeof(jac) * t);
  for (uint64_t k = 0; k < t; k++) {                                     /* :83-92 */
    jac tmp; memset(&tmp, 0, sizeof(tmp)); tmp.inf = 1;
    for (size_t q = 0; q < ntab; q++) {
      size_t base = q * b; size_t cnt = (base + b <= M) ? (size_t)b : M - base; size_t m = 0;
      for (size_t e = 0; e < cnt; e++) {
        size_t el = base + e; size_t i = el / s, j = el % s; uint64_t bit = j + s * k;
        if (bit < 256 && scalar_bit(K[i], (int)bit)) m |= (size_t)1 << e;
      }
      if (m) jac_add(&tmp, &tmp, &T[q * tsz + m]);
    }
    Gs[k] = tmp;
  }
  jac ans = Gs[t - 1];                                                   /* Horner :56-59 */
  for (int64_t k = (int64_t)t - 2; k >= 0; k--) { for (uint64_t d = 0; d < s; d++) jac_dbl(&ans, &ans); jac_add(&ans, &ans, &Gs[k]); }
  jac_to_affine_bytes(out64, &ans);
  free(gs); free(K); free(T); free(Gs); return 0;
}

/* field self-test hooks for tests/test_oracle.py */
int orc_fe_mul(const uint8_t *a, const uint8_t *b, uint8_t *r) { fe x, y, z; fe_from_le(&x, a); fe_from_le(&y, b); fe_mul(&z, &x, &y); fe_to_le(r, &z); return 0; }
int orc_fe_inv(const uint8_t *a, uint8_t *r) { fe x, z; fe_from_le(&x, a); fe_inv(&z, &x); fe_to_le(r, &z); return 0; }
int orc_point_neg(const uint8_t *a64, uint8_t *out64) { jac a, r; jac_from_affine_bytes(&a, a64); jac_neg(&r, &a); jac_to_affine_bytes(out64, &r); return 0; }
