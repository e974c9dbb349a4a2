// ipa_round.cuh -- one inner-product-argument round as ONE launch for proofs over a generator table (n <= 4096).
//
// Replaces: one iteration of FastNIProver2.prove's loop (/root/reference/src/innerproduct/inner_product_prover.py:84-110):
// the fold of a, b (and, in coefficient form, of g, h) with the previous challenge (:107-110), c_L / c_R (:96-97) and the two
// multi-exponentiations L, R (:98-99).
//
// A small proof is a latency chain (rounds x [scalars -> lookups -> addition tree -> host hash]).  Splitting a round into a
// scalar kernel and an MSM kernel puts the scalar work of all n terms on one SM (one block, because c_L / c_R and the in-place
// fold need a barrier) in front of every MSM: 39 us of the 125 us a round took at n = 1024.  Here every block of the table MSM
// derives the scalars of its own 32 terms from the state BEFORE the fold -- a', b', cg', ch' are recomputed where they are
// needed (<= 6 Montgomery products per thread) -- and the state is ping-ponged (in -> out) so that no block reads what another
// one writes.  The block that owns the u term computes c_L or c_R with all its threads while the others are in their lookups.
// Term layout of each MSM (as k_build_lr_sv): slots [0, n/2) g-part, [n/2, n) h-part, slot n = u; points [u | g (n) | h (n)].
#pragma once
#include "ipa.cuh"
#include "fixedbase.cuh"

namespace bp {

struct IpaState { Fq *a, *b, *cg, *ch; };      // a, b standard form (length m), cg, ch Montgomery form (length n)

BP_DI Fq ipa_fold_a(const IpaState& s, u32 j, u32 m, bool fold, const Fq& xm, const Fq& xim) {      // a'[j], j < m
  const Fq lo = ld_fq(s.a + j);
  return fold ? fq_add(fq_mont(lo, xm), fq_mont(ld_fq(s.a + m + j), xim)) : lo;
}
BP_DI Fq ipa_fold_b(const IpaState& s, u32 j, u32 m, bool fold, const Fq& xm, const Fq& xim) {      // b'[j], j < m
  const Fq lo = ld_fq(s.b + j);
  return fold ? fq_add(fq_mont(lo, xim), fq_mont(ld_fq(s.b + m + j), xm)) : lo;
}

__global__ void __launch_bounds__(256) k_ipa_round(const Affine* __restrict__ tab, IpaState in, IpaState out, u32 n, const IpaRound* __restrict__ rp,
                                                   XYZZ* __restrict__ blockpart, u32* __restrict__ ticket, XYZZ* __restrict__ out_xyzz) {
  __shared__ XYZZ sm[256];
  __shared__ u32 s_last;
  __shared__ IpaRound s_rp;
  __shared__ Fq s_red[8];
  __shared__ Fq s_c;
  if (threadIdx.x < sizeof(IpaRound) / 4) ((u32*)&s_rp)[threadIdx.x] = ((const volatile u32*)rp)[threadIdx.x];   // (rp may be mapped host memory)
  __syncthreads();
  const u32 side = blockIdx.y, nbx = gridDim.x;                  // side 0 = L, 1 = R
  const u32 m = s_rp.m, k = m >> 1;                              // m = vector length of this round (after the fold)
  const bool fold = s_rp.fold != 0;
  const Fq xm = s_rp.xm, xim = s_rp.xim;
  const u32 slot = blockIdx.x * 32 + (threadIdx.x >> 3), grp = threadIdx.x & 7;
  const bool owns_u = blockIdx.x * 32 <= n && n < blockIdx.x * 32 + 32;      // this block holds slot n
  if (owns_u) {
    // c_L = <a'_lo, b'_hi>, c_R = <a'_hi, b'_lo>                       inner_product_prover.py:96-97
    Fq acc = fq_zero();
    for (u32 i = threadIdx.x; i < k; i += blockDim.x) {
      const Fq av = ipa_fold_a(in, side == 0 ? i : i + k, m, fold, xm, xim);
      const Fq bv = ipa_fold_b(in, side == 0 ? i + k : i, m, fold, xm, xim);
      acc = fq_add(acc, fq_mont(av, bv));                        // carries a factor R^-1, removed below
    }
    for (int off = 16; off > 0; off >>= 1) {
      Fq x;
#pragma unroll
      for (int i = 0; i < 8; i++) x.v[i] = __shfl_down_sync(0xFFFFFFFFu, acc.v[i], off);
      acc = fq_add(acc, x);
    }
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      Fq t = s_red[0];
      for (int w = 1; w < 8; w++) t = fq_add(t, s_red[w]);
      s_c = fq_to_mont(t);                                       // (sum * R^-1) * R
    }
    __syncthreads();
  }
  XYZZ acc = xyzz_identity();
  if (slot <= n) {
    Fq sc; u32 row;
    if (slot == n) { sc = s_c; row = 0; }
    else {
      const bool gpart = slot < n / 2;
      const u32 s2 = gpart ? slot : slot - n / 2, blk = s2 / k, r = s2 % k;
      // L: g-part = upper halves (t = blk*m + k + r, a'[r]), h-part = lower halves (t = blk*m + r, b'[r + k]);  R: mirrored
      const bool upper = gpart ? side == 0 : side == 1;
      const u32 t = blk * m + (upper ? k : 0) + r;
      const u32 j = gpart ? (side == 0 ? r : r + k) : (side == 0 ? r + k : r);      // index into a' (g-part) or b' (h-part)
      const bool hi2 = (t % (2 * m)) >= m;                       // which factor the previous challenge puts on generator t
      Fq coef, val;
      if (gpart) {
        coef = ld_fq(in.cg + t); if (fold) coef = fq_mont(coef, hi2 ? xm : xim);
        val = ipa_fold_a(in, j, m, fold, xm, xim);
        if (grp == 0) { st_fq(out.cg + t, coef); if (blk == 0) st_fq(out.a + j, val); }
        row = 1 + t;
      } else {
        coef = ld_fq(in.ch + t); if (fold) coef = fq_mont(coef, hi2 ? xim : xm);
        val = ipa_fold_b(in, j, m, fold, xm, xim);
        if (grp == 0) { st_fq(out.ch + t, coef); if (blk == 0) st_fq(out.b + j, val); }
        row = 1 + n + t;
      }
      sc = fq_mont(val, coef);                                   // standard form, reduced
    }
    const u32 word = sc.v[grp];
#pragma unroll 1
    for (int jj = 0; jj < 4; jj++) {
      const u32 d = (word >> (8 * jj)) & 0xFFu;
      if (d) { Affine p = ld_affine(tab + fb_index(row, 4 * grp + jj, d)); xyzz_madd_ni(acc, p); }
    }
  }
  fb_block_tail(sm, &s_last, acc, side, nbx, blockpart, ticket, nullptr, out_xyzz);
}

}  // namespace bp
