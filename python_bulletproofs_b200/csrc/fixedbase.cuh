// fixedbase.cuh -- precomputed-table ("fixed-base") multi-scalar multiplication for generator sets that are used again
// and again: the g_i, h_i, g, h, u of a Bulletproofs instance.
//
// Replaces (for REPEATED generator sets only; the first use of a set always takes the bucket method of msm.cuh):
//   Pippenger.multiexp on the commitment / L,R / verifier call sites that always pass the same group elements
//   (/root/reference/src/utils/commitments.py:9-13, /root/reference/src/innerproduct/inner_product_prover.py:98-99,
//    /root/reference/src/rangeproofs/rangeproof_prover.py:47,57,78-86).
//
// Table: for every point P_g and every byte position w of a 256-bit scalar, the 255 affine points d * 2^(8w) * P_g,
// d = 1..255  (32 windows x 255 entries x 64 B = 522 KB per point; 130 generators of a 64-bit range proof = 68 MB, the
// 2049 of a 1024-wide inner-product argument = 1.07 GB -- small change in 180 GB of HBM3e).  An MSM is then
// sum_i sum_w T[g_i][w][byte_w(k_i)]: 32 mixed additions per term, NO doublings, no bucket reduction and no Horner
// chain, which is what bounds the latency of the small MSMs of a prover (0.58 ms -> ~0.1 ms at 65 terms).
// Scalars are taken mod q and used unsigned (no recoding, no carries between windows), so each (term, window) is
// independent work.  Results are canonical affine points, hence bit-identical to the bucket method.
#pragma once
#include "ec.cuh"
#include "fq.cuh"
#include "coop4.cuh"

namespace bp {

#define BP_FB_WINDOWS 32
#define BP_FB_ENTRIES 255
#define BP_FB_BUILD_GENS 256          // generators per build launch (bounds the XYZZ scratch: 256*32*255*128 B = 268 MB)

BP_DI size_t fb_index(u32 gi, u32 w, u32 d) { return ((size_t)gi * BP_FB_WINDOWS + w) * BP_FB_ENTRIES + (d - 1); }

BP_DI Fp ld_fp_coherent(const Fp* p) {       // plain loads: the build kernel re-reads what it wrote
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  Fp r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// One thread per (generator, window): B = 2^(8w) * P, then the chain B, 2B, ..., 255B in XYZZ (scratch), normalised to
// affine with ONE field inversion per thread (Montgomery's trick over the 255 ZZZ values; the running products are
// parked in the x half of the output slots until the backward pass overwrites them).
__global__ void __launch_bounds__(64) k_fb_build(const Affine* __restrict__ pts, u32 g0, u32 ng, XYZZ* __restrict__ scratch,
                                                 Affine* __restrict__ tab) {
  u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ng * BP_FB_WINDOWS) return;
  const u32 gi = g0 + t / BP_FB_WINDOWS, w = t % BP_FB_WINDOWS;
  Affine P = ld_affine(pts + gi);
  Affine* out = tab + fb_index(gi, w, 1);
  if (affine_is_identity(P)) {
    Affine z; z.x = fp_zero(); z.y = fp_zero();
    for (int d = 0; d < BP_FB_ENTRIES; d++) st_affine(out + d, z);
    return;
  }
  XYZZ B = xyzz_from_affine(P);
  for (u32 i = 0; i < 8 * w; i++) B = xyzz_dbl_ni(B);
  const Affine Ba = xyzz_to_affine(B);
  XYZZ* sc = scratch + (size_t)t * BP_FB_ENTRIES;
  XYZZ acc = xyzz_from_affine(Ba);
  Fp pre = fp_one();
  for (int d = 1; d <= BP_FB_ENTRIES; d++) {
    if (d > 1) xyzz_madd_ni(acc, Ba);                  // d = 2 is the doubling case of the complete formula
    st_xyzz(sc + d - 1, acc);
    pre = fp_mul(pre, acc.ZZZ);                        // d * B is never the identity (d < q, B of prime order)
    st_fp(&out[d - 1].x, pre);
  }
  Fp inv = fp_inv(pre);
  for (int d = BP_FB_ENTRIES; d >= 1; d--) {
    XYZZ e = ld_xyzz(sc + d - 1);
    Fp prev = d > 1 ? ld_fp_coherent(&out[d - 2].x) : fp_one();
    Fp zi3 = fp_mul(inv, prev);                        // ZZZ_d^-1
    inv = fp_mul(inv, e.ZZZ);
    Fp zi2 = fp_mul(fp_sqr(e.ZZ), fp_sqr(zi3));        // ZZ^-1 = ZZ^2 * ZZZ^-2
    Affine r; r.x = fp_canon(fp_mul(e.X, zi2)); r.y = fp_canon(fp_mul(e.Y, zi3));
    st_affine(out + d - 1, r);
  }
}

// Table MSM, step 1.  blockIdx.y = MSM m with terms [offsets[m], offsets[m+1]); a block of 256 threads takes 32
// consecutive terms: thread = (term, group of 4 byte-windows) does up to 4 mixed additions, then the block folds its
// 256 partial sums with 4-lane cooperative additions (3 serial + 6 tree levels) into blockpart[m * nbx + blockIdx.x].
// idx (optional) maps a term to its table row (without it term j of every MSM reads row j); scalars are 32-byte values reduced mod q here (pippenger.py:26).
// ticket / out_final (optional, both or neither): the block that finishes LAST for its MSM (atomic ticket) also does step 2 --
// it adds the nbx block sums and writes the canonical affine result to out_final[m] (or the XYZZ sum to out_final_xyzz[m]), which may be mapped host memory (the
// latency-bound IPA rounds: one launch and no copy-out per round instead of two launches and a copy).
BP_DI XYZZ ld_xyzz_cg(const XYZZ* p) {       // L2 loads: partial sums written by other blocks of the same launch
  XYZZ r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = __ldcg(q + i);
  Fp* f = &r.X;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    f[i].v[0] = w[2 * i].x; f[i].v[1] = w[2 * i].y; f[i].v[2] = w[2 * i].z; f[i].v[3] = w[2 * i].w;
    f[i].v[4] = w[2 * i + 1].x; f[i].v[5] = w[2 * i + 1].y; f[i].v[6] = w[2 * i + 1].z; f[i].v[7] = w[2 * i + 1].w;
  }
  return r;
}
// Tail of a table-MSM block (256 threads, `acc` = this thread's partial sum).  Phase 0: the block's 256 partial sums -> one (each
// quad adds its 4, then a tree over the 64 quads) -> blockpart[m * nbx + blockIdx.x].  Phase 1 (only the block that finishes last
// for its MSM, when a ticket is given): the nbx block sums -> one -> out_final[m] (canonical affine) or out_final_xyzz[m].
// Both phases run through ONE loop with ONE cooperative-addition site (code size: these kernels are latency chains).
BP_DI void fb_block_tail(XYZZ* sm, u32* s_last, const XYZZ& acc, u32 m, u32 nbx, XYZZ* __restrict__ blockpart, u32* __restrict__ ticket,
                         Affine* __restrict__ out_final, XYZZ* __restrict__ out_final_xyzz) {
  st_xyzz(&sm[threadIdx.x], acc);
  __syncthreads();
  const int lane = threadIdx.x & 31, role = lane & 3, base = lane & ~3;
  const u32 q = threadIdx.x >> 2;
#pragma unroll 1
  for (int phase = 0; phase < 2; phase++) {
    // serial steps (each quad adds further operands to its own sum), then a tree over `width` quads -- only as many levels as
    // there are live sums (3 block sums of a 64-wide proof: 2 levels, not 6)
    const u32 nserial = phase == 0 ? 3u : (nbx + 63) / 64 - 1;
    u32 width = 64;
    if (phase == 1) { width = 1; while (width < nbx && width < 64) width <<= 1; }
    u32 nlev = 0; while ((1u << nlev) < width) nlev++;
    XYZZ v = phase == 0 ? ld_xyzz(&sm[4 * q]) : (q < nbx ? ld_xyzz_cg(blockpart + (size_t)m * nbx + q) : xyzz_identity());
    if (nserial == 0) {                                          // nothing to add before the tree: publish the loaded sums
      __syncthreads();
      if (q < width && role == 0) st_xyzz(&sm[q], v);
      __syncthreads();
    }
#pragma unroll 1
    for (u32 st = 0; st < nserial + nlev; st++) {
      XYZZ x;
      if (st < nserial) {
        if (phase == 0) x = ld_xyzz(&sm[4 * q + st + 1]);
        else { const u32 i = (st + 1) * 64 + q; x = i < nbx ? ld_xyzz_cg(blockpart + (size_t)m * nbx + i) : xyzz_identity(); }
      } else {
        const u32 off = (width >> 1) >> (st - nserial);
        x = q < off ? ld_xyzz(&sm[q + off]) : xyzz_identity();
      }
      v = coop_add(v, x, role, base);
      if (st + 1 >= nserial) {                                   // publish: after the last serial step and after every tree level
        const u32 live = st + 1 == nserial ? width : (width >> 1) >> (st - nserial);
        __syncthreads();
        if (q < live && role == 0) st_xyzz(&sm[q], v);
        __syncthreads();
      }
    }
    if (phase == 0) {
      if (threadIdx.x == 0) st_xyzz(blockpart + (size_t)m * nbx + blockIdx.x, v);
      if (!ticket) return;
      if (threadIdx.x == 0) {
        __threadfence();                                         // the block sum is visible before the ticket is taken
        *s_last = atomicAdd(ticket + m, 1u) == nbx - 1 ? 1u : 0u;
      }
      __syncthreads();
      if (!*s_last) return;
      __threadfence();
    } else if (threadIdx.x == 0) {
      ticket[m] = 0;                                             // ready for the next launch
      if (out_final_xyzz) st_xyzz(out_final_xyzz + m, v);         // the caller finishes the conversion (IPA rounds: on the host)
      else st_affine(out_final + m, xyzz_to_affine(v, true));
      __threadfence_system();
    }
  }
}
__global__ void __launch_bounds__(256) k_fb_msm(const Affine* __restrict__ tab, const u32* __restrict__ idx, const Fq* __restrict__ sc,
                                                const u32* __restrict__ offsets, u32 single_n, XYZZ* __restrict__ blockpart,
                                                u32* __restrict__ ticket = nullptr, Affine* __restrict__ out_final = nullptr,
                                                XYZZ* __restrict__ out_final_xyzz = nullptr) {
  __shared__ XYZZ sm[256];
  __shared__ u32 s_last;
  const u32 m = blockIdx.y, nbx = gridDim.x;
  const u32 lo = offsets ? offsets[m] : 0u, hi = offsets ? offsets[m + 1] : single_n;
  const u32 t = lo + blockIdx.x * 32 + (threadIdx.x >> 3), grp = threadIdx.x & 7;
  XYZZ acc = xyzz_identity();
  if (t < hi) {
    const u32* kw = reinterpret_cast<const u32*>(sc + t);
    Fq k;
#pragma unroll
    for (int i = 0; i < 8; i++) k.v[i] = __ldg(kw + i);
    k = fq_reduce(k);
    const u32 word = k.v[grp];                                   // windows 4*grp .. 4*grp+3 are the bytes of limb grp
    const u32 gi = idx ? __ldg(idx + t) : t - lo;               // no map: term j of an MSM reads table row j
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
      const u32 d = (word >> (8 * j)) & 0xFFu;
      if (d) { Affine p = ld_affine(tab + fb_index(gi, 4 * grp + j, d)); xyzz_madd_ni(acc, p); }
    }
  }
  fb_block_tail(sm, &s_last, acc, m, nbx, blockpart, ticket, out_final, out_final_xyzz);
}

// Table MSM, step 2: one block per MSM adds its nbx block sums (quads stride over them, then a tree) and writes the
// canonical affine result (and/or the XYZZ value, for callers that add further terms).
__global__ void __launch_bounds__(256) k_fb_finish(const XYZZ* __restrict__ blockpart, u32 nbx, Affine* __restrict__ out,
                                                   XYZZ* __restrict__ out_xyzz) {
  __shared__ XYZZ sm[64];
  const u32 m = blockIdx.x;
  const int lane = threadIdx.x & 31, role = lane & 3, base = lane & ~3;
  const u32 q = threadIdx.x >> 2, nq = blockDim.x >> 2;
  XYZZ acc = xyzz_identity();
  for (u32 i0 = 0; i0 < nbx; i0 += nq) {
    const u32 i = i0 + q;
    XYZZ x = i < nbx ? ld_xyzz(blockpart + (size_t)m * nbx + i) : xyzz_identity();
    acc = coop_add(acc, x, role, base);
  }
  if (nq > 1) {
    if (role == 0) st_xyzz(&sm[q], acc);
    __syncthreads();
#pragma unroll 1
    for (u32 off = nq >> 1; off > 0; off >>= 1) {
      XYZZ x = (q < off) ? ld_xyzz(&sm[q + off]) : xyzz_identity();
      acc = coop_add(acc, x, role, base);
      __syncthreads();
      if (q < off && role == 0) st_xyzz(&sm[q], acc);
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    if (out_xyzz) st_xyzz(out_xyzz + m, acc);
    if (out) st_affine(out + m, xyzz_to_affine(acc, true));        // thread 0 of its own block
  }
}

// ---- 16-bit windows for the batch verifier's generator table ------------------------------------------------------------
// T16[g][w][d] = d * 2^(16w) * P_g, d = 1..65535 (16 windows, 67 MB per point: 8.9 GB for the 133 points of a 64-bit range
// proof -- HBM3e capacity spent to halve the lookups per term).  Built from the byte table: d = 256*hi + lo is
// T8[g][2w+1][hi] + T8[g][2w][lo], one affine addition per entry with the 255 inversions of a (g, w, hi) row shared by
// Montgomery's trick (running products parked in the x half of the output slots).  The two summands are distinct multiples
// of P below q, so neither the doubling nor the cancelling case of the addition can occur.
#define BP_FB16_WINDOWS 16
#define BP_FB16_ENTRIES 65535
BP_DI size_t fb_index16(u32 gi, u32 w, u32 d) { return ((size_t)gi * BP_FB16_WINDOWS + w) * BP_FB16_ENTRIES + (d - 1); }

__global__ void __launch_bounds__(128) k_fb_build16(const Affine* __restrict__ tab8, u32 n, Affine* __restrict__ tab16) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;          // (g, w, hi)
  if (t >= n * BP_FB16_WINDOWS * 256) return;
  const u32 hi = t & 255u, w = (t >> 8) % BP_FB16_WINDOWS, gi = (t >> 8) / BP_FB16_WINDOWS;
  const Affine* lo_row = tab8 + fb_index(gi, 2 * w, 1);         // lo_row[lo - 1] = lo * 2^(16w) * P
  Affine* out = tab16 + fb_index16(gi, w, 256 * hi + (hi ? 0 : 1));   // first entry of this row: d = 256*hi (hi > 0) or d = 1
  if (hi == 0) {
    for (u32 lo = 1; lo < 256; lo++) st_affine(out + lo - 1, ld_affine(lo_row + lo - 1));
    return;
  }
  const Affine A = ld_affine(tab8 + fb_index(gi, 2 * w + 1, hi));
  st_affine(out, A);                                             // lo = 0
  if (affine_is_identity(A)) {                                   // identity generator: every multiple is the identity
    for (u32 lo = 1; lo < 256; lo++) st_affine(out + lo, A);
    return;
  }
  Fp pre = fp_one();
  for (u32 lo = 1; lo < 256; lo++) {
    Fp dx = fp_sub(ld_fp(&lo_row[lo - 1].x), A.x);
    pre = fp_mul(pre, dx);
    st_fp(&out[lo].x, pre);
  }
  Fp inv = fp_inv(pre);
  for (u32 lo = 255; lo >= 1; lo--) {
    const Affine B = ld_affine(lo_row + lo - 1);
    Fp dx = fp_sub(B.x, A.x);
    Fp prev = lo > 1 ? ld_fp_coherent(&out[lo - 1].x) : fp_one();
    Fp dxi = fp_mul(inv, prev);                                  // 1 / (xB - xA)
    inv = fp_mul(inv, dx);
    Fp lam = fp_mul(fp_sub(B.y, A.y), dxi);
    Fp x3 = fp_sub(fp_sub(fp_sqr(lam), A.x), B.x);
    Fp y3 = fp_sub(fp_mul(lam, fp_sub(A.x, x3)), A.y);
    Affine r; r.x = fp_canon(x3); r.y = fp_canon(y3);
    st_affine(out + lo, r);
  }
}

// ---- throughput form for batches of MSMs whose terms mix table rows and other points (batch verifier) ----------------
// Step 1: one WARP per MSM, lane l owns byte-window l: for every term with a table row (idx < nfixed) one lookup and one
// mixed addition per lane; terms with idx >= nfixed belong to the caller's other pass and are skipped (warp-uniform).
// The 32 lane sums go to part[32 * m + l]; step 2 folds them.  Few registers here (occupancy), the cooperative adds
// live in the second kernel.
__global__ void __launch_bounds__(256) k_fb_lookup_warp(const Affine* __restrict__ tab, const u32* __restrict__ idx, const Fq* __restrict__ sc,
                                                        const u32* __restrict__ offsets, u32 nmsm, u32 nfixed, XYZZ* __restrict__ part) {
  const u32 m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= nmsm) return;
  const u32 lo = __ldg(offsets + m), hi = __ldg(offsets + m + 1);
  XYZZ acc = xyzz_identity();
  for (u32 t = lo; t < hi; t++) {
    const u32 gi = __ldg(idx + t);
    if (gi >= nfixed) continue;
    const u32* kw = reinterpret_cast<const u32*>(sc + t);
    Fq k;
#pragma unroll
    for (int i = 0; i < 8; i++) k.v[i] = __ldg(kw + i);
    k = fq_reduce(k);
    const u32 d = (k.v[lane >> 2] >> (8 * (lane & 3))) & 0xFFu;
    if (d) { Affine p = ld_affine(tab + fb_index(gi, lane, d)); xyzz_madd_ni(acc, p); }
  }
  st_xyzz(part + (size_t)32 * m + lane, acc);
}
// The same for batches of LARGE table MSMs (aggregated range proofs: 2048 generator terms in one equation): one warp per
// (MSM, slice of `slice` consecutive terms), lane = byte-window; an empty slice writes identities.  part[(m * nsl + s) * 32 + lane];
// k_fb_fold_warp then folds the 32 lane sums of every slice and k_fb_finish adds the nsl slice sums of every MSM.
__global__ void __launch_bounds__(256) k_fb_lookup_slice(const Affine* __restrict__ tab, const u32* __restrict__ idx, const Fq* __restrict__ sc,
                                                         const u32* __restrict__ offsets, u32 nmsm, u32 nsl, u32 slice, XYZZ* __restrict__ part) {
  const u32 w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= nmsm * nsl) return;
  const u32 m = w / nsl, s = w % nsl;
  const u32 lo = __ldg(offsets + m) + s * slice, end = __ldg(offsets + m + 1);
  const u32 hi = lo + slice < end ? lo + slice : end;
  XYZZ acc = xyzz_identity();
  for (u32 t = lo; t < hi; t++) {                     // (lo >= hi: empty slice)
    const u32 gi = __ldg(idx + t);
    const u32* kw = reinterpret_cast<const u32*>(sc + t);
    Fq k;
#pragma unroll
    for (int i = 0; i < 8; i++) k.v[i] = __ldg(kw + i);
    k = fq_reduce(k);
    const u32 d = (k.v[lane >> 2] >> (8 * (lane & 3))) & 0xFFu;
    if (d) { Affine p = ld_affine(tab + fb_index(gi, lane, d)); xyzz_madd_ni(acc, p); }
  }
  st_xyzz(part + (size_t)32 * w + lane, acc);
}
// Step 2: one warp per MSM = 8 quads: each quad adds 4 lane sums, then a 3-level tree; optionally adds `other[m]` (the
// caller's partial for the same MSM) and writes out[m].
__global__ void __launch_bounds__(128) k_fb_fold_warp(const XYZZ* __restrict__ part, const XYZZ* __restrict__ other, u32 nmsm,
                                                      XYZZ* __restrict__ out) {
  __shared__ XYZZ sm[4][8];
  const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  u32 m = blockIdx.x * 4 + warp;
  const bool live = m < nmsm;
  if (!live) m = nmsm - 1;                                   // whole warps stay in the shuffles; the duplicate does not store
  const int role = lane & 3, base = lane & ~3;
  const u32 q = lane >> 2;
  const XYZZ* src = part + (size_t)32 * m + 4 * q;
  XYZZ v = ld_xyzz(src);
#pragma unroll 1
  for (int j = 1; j < 4; j++) { XYZZ x = ld_xyzz(src + j); v = coop_add(v, x, role, base); }
  if (role == 0) st_xyzz(&sm[warp][q], v);
  __syncwarp();
#pragma unroll 1
  for (u32 off = 4; off > 0; off >>= 1) {
    XYZZ x = (q < off) ? ld_xyzz(&sm[warp][q + off]) : xyzz_identity();
    v = coop_add(v, x, role, base);
    __syncwarp();
    if (q < off && role == 0) st_xyzz(&sm[warp][q], v);
    __syncwarp();
  }
  if (other) { XYZZ x = ld_xyzz(other + m); v = coop_add(v, x, role, base); }
  if (live && lane == 0) st_xyzz(out + m, v);
}

}  // namespace bp
