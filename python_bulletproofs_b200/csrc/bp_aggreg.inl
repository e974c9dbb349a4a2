// bp_aggreg.inl -- batch verification of AGGREGATED range proofs (m values x n bits per proof) over one generator set
// (included at the end of bp_gpu.cu).
//
// Replaces, per proof: AggregRangeVerifier.verify (/root/reference/src/rangeproofs/rangeproof_aggreg_verifier.py:42-108) with
// Verifier1.verify / Verifier2.verify underneath (/root/reference/src/innerproduct/inner_product_verifier.py:36-58,104-147).
// With m = 1 it is RangeVerifier.verify, i.e. a second, generic route to the decisions of bp_rp_verify_batch.
//
// The single-value batch verifier (verify.cuh, svar.cuh) is specialised to N = n <= 128 generators per side: one block per proof
// expands the scalars, warps are dealt to equations, the generator table has 16-bit windows.  An aggregated proof has N = n*m up
// to 2048 generators per side, so this path is built from the generic pieces instead:
//   host   (OpenMP over proofs, 4 x 64-bit Montgomery limbs, rp_algebra.h): the three verify_transcript checks, y^-1 and x_j^-1,
//          delta(y, z), y^-i, z^(2+j) 2^i, the s vector, and EVERY term scalar of the four equations
//            E1  (t_hat - delta) g + taux h - sum_j z^(2+j) V_j - x T1 - x^2 T2                              == O
//            E2  A + x S - z Gsum + sum_i (z + zz_i y^-i) hs_i - mu h + (x1 t_hat) u - P_new                 == O
//            E3  x1 u - u_new                                                                                == O
//            E4  sum_i (a s_i) gs_i + sum_i (b s_i^-1 y^-i) hs_i + (a b) u_new - P_new - sum_j x_j^2 L_j - sum_j x_j^-2 R_j == O
//          (hsp_i = y^-i hs_i folded into the scalars; Gsum = sum gs_i a row of the generator table);
//   device the 3N + 6 generator terms of a proof as byte-table lookups (fixedbase.cuh: k_fb_lookup_slice -> k_fb_fold_warp -> k_fb_finish over 4 MSMs per proof;
//          the bucket method over the generator rows until the set has its table), the m + 8 + 2 log2(N) proof-specific terms
//          through the batched bucket MSM (msm.cuh), an on-curve check of every proof point, and the identity test of the four sums.
// Exact checks, no random linear combination: the decisions are the reference's.

namespace bp {

// XYZZ sums of the two parts (MSM index 4p + e) -> accept byte of proof p
__global__ void __launch_bounds__(128) k_rp_aggr_accept(const XYZZ* __restrict__ tabres, const XYZZ* __restrict__ varres, u32 nproofs,
                                                        const uint8_t* __restrict__ bad, uint8_t* __restrict__ accept) {
  const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nproofs) return;
  bool ok = bad[p] == 0;
  for (u32 e = 0; e < 4; e++) {
    XYZZ a = ld_xyzz(tabres + 4 * (size_t)p + e), b = ld_xyzz(varres + 4 * (size_t)p + e);
    xyzz_add_ni(a, b);
    ok = ok && xyzz_is_identity(a);
  }
  accept[p] = ok ? 1 : 0;
}

struct AggOffs { size_t V, A, S, T1, T2, Taux, Mu, That, Unew, Pnew, a, b, Xs, Ls, Rs, stride; u32 L; };
static AggOffs agg_offsets(size_t N, size_t m) {
  AggOffs o; o.L = 0; while (((size_t)1 << o.L) < N) o.L++;
  o.V = 0; o.A = 64 * m; o.S = o.A + 64; o.T1 = o.S + 64; o.T2 = o.T1 + 64; o.Taux = o.T2 + 64; o.Mu = o.Taux + 32; o.That = o.Mu + 32;
  o.Unew = o.That + 32; o.Pnew = o.Unew + 64; o.a = o.Pnew + 64; o.b = o.a + 32; o.Xs = o.b + 32; o.Ls = o.Xs + 32 * (size_t)o.L;
  o.Rs = o.Ls + 64 * (size_t)o.L; o.stride = o.Rs + 64 * (size_t)o.L;
  return o;
}

// The reference's three verify_transcript methods on one proof (rangeproof_aggreg_verifier.py:42-53 = rangeproof_verifier.py:42-53,
// inner_product_verifier.py:36-42 and :104-125): 1 = passed, 0 = "Proof invalid", 2 = the reference would raise (IndexError /
// ValueError / y = 0): the caller replays such a proof through the Python classes.  y, z, x, x1 come back reduced.
static uint8_t agg_host_verdict(const uint8_t* pr, const AggOffs& o, const uint8_t* transcripts, const uint64_t* tro, uint32_t start,
                                Fq* y, Fq* z, Fq* x, Fq* x1) {
  struct Slot { size_t first, second; };
  std::vector<Slot> sl;
  auto split_first = [&](const uint8_t* s, size_t n, size_t need) -> size_t {
    if (need > n + 1) need = n + 1;
    sl.resize(need);
    size_t cnt = 0, st = 0;
    while (cnt < need) {
      const uint8_t* amp = st <= n ? (const uint8_t*)memchr(s + st, '&', n - st) : nullptr;
      const size_t end = amp ? (size_t)(amp - s) : n;
      sl[cnt].first = st; sl[cnt].second = end - st; cnt++;
      if (!amp) break;
      st = end + 1;
    }
    return cnt;
  };
  *y = fq_one(); *z = fq_one(); *x = fq_one(); *x1 = fq_one();
  const uint8_t* t0 = transcripts + tro[0]; const size_t t0n = tro[1] - tro[0];
  if (split_first(t0, t0n, 8) < 8) return 2;
  if (!b64_point_eq(t0 + sl[1].first, sl[1].second, pr + o.A) || !b64_point_eq(t0 + sl[2].first, sl[2].second, pr + o.S)) return 0;
  if (!decimal_to_fq_fast(t0 + sl[3].first, sl[3].second, y) || !decimal_to_fq_fast(t0 + sl[4].first, sl[4].second, z)) return 2;
  if (!b64_point_eq(t0 + sl[5].first, sl[5].second, pr + o.T1) || !b64_point_eq(t0 + sl[6].first, sl[6].second, pr + o.T2)) return 0;
  if (!decimal_to_fq_fast(t0 + sl[7].first, sl[7].second, x)) return 2;
  if (fq_is_zero(*y)) return 2;                                      // y.inv() raises in the reference
  const uint8_t* t1 = transcripts + tro[1]; const size_t t1n = tro[2] - tro[1];
  if (split_first(t1, t1n, 2) < 2) return 2;
  *x1 = mod_hash_q(t1, sl[0].second + 1);                            // parts[0] + b"&"
  if (!decimal_slot_eq(t1 + sl[1].first, sl[1].second, *x1)) return 0;
  const uint8_t* t2 = transcripts + tro[2]; const size_t t2n = tro[3] - tro[2];
  const size_t need = (size_t)start + 3 * (size_t)o.L;
  if (o.L && split_first(t2, t2n, need) < need) return 2;
  RunningModHash rh;
  for (u32 j = 0; j < o.L; j++) {
    const Slot& sL = sl[start + 3 * j]; const Slot& sR = sl[start + 3 * j + 1]; const Slot& sX = sl[start + 3 * j + 2];
    if (!b64_point_eq(t2 + sL.first, sL.second, pr + o.Ls + 64 * j) || !b64_point_eq(t2 + sR.first, sR.second, pr + o.Rs + 64 * j)) return 0;
    Fq xj; fq_from_le(&xj, pr + o.Xs + 32 * j);
    const Fq want = rh.challenge(t2, sX.first);
    if (!fq_eq(xj, want) || !decimal_slot_eq(t2 + sX.first, sX.second, want)) return 0;
  }
  return 1;
}

}  // namespace bp

extern "C" {

size_t bp_rp_aggreg_proof_stride(size_t n, size_t m) { return agg_offsets(n * m, m).stride; }

int bp_rp_verify_aggreg_batch(const uint8_t* gs64, const uint8_t* hs64, const uint8_t g64[64], const uint8_t h64[64], const uint8_t u64_[64],
                              size_t n, size_t m, const uint8_t* proofs, size_t proof_stride, size_t nproofs, const uint8_t* transcripts,
                              const uint64_t* tr_off, const uint32_t* start_transcript, uint8_t* accept) {
  using namespace rpa;
  BP_NEED_INIT();
  const size_t N = n * m;
  if (n == 0 || m == 0 || (N & (N - 1)) || N > 2048) return fail("bp_rp_verify_aggreg_batch: n*m must be a power of two <= 2048");
  if (nproofs == 0) return 0;
  const AggOffs o = agg_offsets(N, m);
  const u32 L = o.L;
  if (proof_stride < o.stride) return fail("bp_rp_verify_aggreg_batch: proof_stride too small");
  const size_t F = 2 * N + 5;                                    // table rows [gs | hs | g | h | u | Gsum | Hsum]
  const u32 rG = (u32)(2 * N), rH = rG + 1, rU = rG + 2, rGsum = rG + 3;
  const size_t TT = 3 * N + 6, VT = m + 8 + 2 * (size_t)L;       // generator / proof-specific terms per proof
  // chunks: large enough to fill the device and to amortise the ~0.4 ms of dependent tail kernels per chunk (measured at 16 x 64 bits:
  // 64 / 128 / 256 / 512 proofs per chunk -> 30.9 / 36.5 / 42.5 / 44.9 k proofs/s), at least four per batch where that is possible so
  // that host and device overlap, at most 64 MB of term scalars each; double-buffered
  static const size_t ch_max = [] { const char* e = getenv("BP_AGG_CHUNK"); long v = e ? atol(e) : 512; return (size_t)(v >= 1 && v <= 4096 ? v : 512); }();
  size_t CH = (nproofs + 3) / 4;
  if (CH < 64) CH = 64;
  if (CH > ch_max) CH = ch_max;
  if (CH > nproofs) CH = nproofs;
  while (CH > 1 && CH * TT * 36 > ((size_t)64 << 20)) CH /= 2;
  const size_t b_tsc = CH * TT * 32, b_tidx = CH * TT * 4, b_off = (4 * CH + 1) * 4, b_vpt = CH * VT * 64, b_vsc = CH * VT * 32;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t stage_each = al(b_tsc) + al(b_tidx) + 2 * al(b_off) + al(b_vpt) + al(b_vsc) + 2 * al(CH);
  uint8_t* stage = g.pinned_stage(2 * stage_each);
  if (!stage) return fail("pinned staging allocation failed");
  struct HostSet { uint8_t* tsc; u32* tidx; u32* toff; u32* voff; uint8_t* vpt; uint8_t* vsc; uint8_t* ok; uint8_t* dev; } H[2];
  for (int b = 0; b < 2; b++) {
    uint8_t* q = stage + (size_t)b * stage_each;
    H[b].tsc = q; q += al(b_tsc); H[b].tidx = (u32*)q; q += al(b_tidx); H[b].toff = (u32*)q; q += al(b_off); H[b].voff = (u32*)q; q += al(b_off);
    H[b].vpt = q; q += al(b_vpt); H[b].vsc = q; q += al(b_vsc); H[b].ok = q; q += al(CH); H[b].dev = q;
  }
  Affine* rows = (Affine*)g.ws_pts.ensure((F + 2 * CH * VT) * sizeof(Affine));
  Fq* d_sc2 = (Fq*)g.ws_terms_sc.ensure(2 * CH * (TT + VT) * sizeof(Fq));
  u32* d_tidx2 = (u32*)g.ws_idx.ensure(2 * CH * TT * sizeof(u32));
  u32* d_off2 = (u32*)g.ws_off.ensure(4 * (4 * CH + 1) * sizeof(u32));
  XYZZ* d_res2 = (XYZZ*)g.ws_fb_var.ensure(2 * 2 * 4 * CH * sizeof(XYZZ));
  uint8_t* d_acc2 = (uint8_t*)g.ws_misc.ensure(4 * CH + 256);
  if (!rows || !d_sc2 || !d_tidx2 || !d_off2 || !d_res2 || !d_acc2) return fail("device allocation failed");
  if (g.ensure_stage_events()) return 1;
  // ---- generator rows, Gsum / Hsum, table (second sighting of the set onwards)
  BP_CUDA(cudaMemcpyAsync(rows, gs64, N * 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(rows + N, hs64, N * 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(rows + rG, g64, 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(rows + rH, h64, 64, cudaMemcpyHostToDevice, g.stream));
  BP_CUDA(cudaMemcpyAsync(rows + rU, u64_, 64, cudaMemcpyHostToDevice, g.stream));
  const Affine* tab = nullptr;
  bool have_sums = false;
  if (fb_enabled()) {
    const FbSrc src = {{gs64, hs64, g64, h64, u64_}, {N * 64, N * 64, 64, 64, 64}, 5};
    const uint64_t key = src.hash(0x72707631ull);
    auto it = fb.tabs.find(key ^ ((uint64_t)F * 0xD6E8FEB86659FD93ull));
    have_sums = it != fb.tabs.end() && it->second.n == F && src.equals(it->second.src);      // the table holds the Gsum / Hsum rows already
    if (!have_sums) {
      ++g.nlaunch, k_sum_points<<<1, 128, 0, g.stream>>>(rows, (u32)N, rows + rGsum);
      ++g.nlaunch, k_sum_points<<<1, 128, 0, g.stream>>>(rows + N, (u32)N, rows + rGsum + 1);
    }
    tab = fb_get(key, src, rows, F);
  }
  if (!tab && !have_sums) {      // bucket method over the rows: the sums are needed as points
    ++g.nlaunch, k_sum_points<<<1, 128, 0, g.stream>>>(rows, (u32)N, rows + rGsum);
    ++g.nlaunch, k_sum_points<<<1, 128, 0, g.stream>>>(rows + N, (u32)N, rows + rGsum + 1);
  }
  unsigned nthreads = std::thread::hardware_concurrency();
  if (nthreads == 0) nthreads = 1;
  if (const char* lws = getenv("LOCAL_WORLD_SIZE")) { long w = atol(lws); if (w > 1) nthreads = nthreads / (unsigned)w ? nthreads / (unsigned)w : 1; }
  if (nthreads > 64) nthreads = 64;
  const bool timing = getenv("BP_VERIFY_TIMING") != nullptr;
  const auto t_call0 = std::chrono::steady_clock::now();
  double host_ms = 0;
  H4 two_pow[64];                                                // 2^i mod q, i < n (n <= 64 in every use; larger n recomputed below)
  { H4 t = one(); for (size_t i = 0; i < 64; i++) { two_pow[i] = t; t = add(t, t); } }
  // finalise chunk `b` (wait for its accept bytes, merge with the host verdicts)
  auto finish = [&](int b, size_t lo, size_t cn) -> int {
    BP_CUDA(cudaEventSynchronize(g.stage_ev[b]));
    for (size_t pi = 0; pi < cn; pi++) accept[lo + pi] = H[b].ok[pi] != 1 ? H[b].ok[pi] : H[b].dev[pi];
    return 0;
  };
  BP_CUDA(cudaEventRecord(g.ev_copy_gate, g.stream));             // earlier calls may still be reading these workspaces
  BP_CUDA(cudaStreamWaitEvent(g.copy_stream, g.ev_copy_gate, 0));
  int cur = 0;
  size_t prev_lo = 0, prev_cn = 0;
  for (size_t lo = 0; lo < nproofs; lo += CH, cur ^= 1) {
    const size_t cn = nproofs - lo < CH ? nproofs - lo : CH;
    HostSet& hs_ = H[cur];
    const auto t_h0 = std::chrono::steady_clock::now();
    // ---- host: verdict + every term scalar of the chunk's proofs (buffer set `cur` was finalised two chunks ago)
    const long nt = (long)(nthreads > cn ? (unsigned)cn : nthreads);
#pragma omp parallel for schedule(dynamic, 1) num_threads((int)nt)
    for (long pi = 0; pi < (long)cn; pi++) {
      const size_t p = lo + (size_t)pi;
      const uint8_t* pr = proofs + p * proof_stride;
      Fq yq, zq, xq, x1q;
      const uint8_t verdict = agg_host_verdict(pr, o, transcripts, tr_off + 3 * p, start_transcript[p], &yq, &zq, &xq, &x1q);
      hs_.ok[pi] = verdict;
      uint8_t* tsc = hs_.tsc + (size_t)pi * TT * 32; u32* tidx = hs_.tidx + (size_t)pi * TT;
      uint8_t* vsc = hs_.vsc + (size_t)pi * VT * 32; uint8_t* vpt = hs_.vpt + (size_t)pi * VT * 64;
      if (verdict != 1) {                                        // nothing to evaluate: zero scalars, identity points
        memset(tsc, 0, TT * 32); memset(vsc, 0, VT * 32); memset(vpt, 0, VT * 64);
        for (size_t t = 0; t < TT; t++) tidx[t] = 0;
        continue;
      }
      H4 y, z, x, x1;
      memcpy(y.v, yq.v, 32); memcpy(z.v, zq.v, 32); memcpy(x.v, xq.v, 32); memcpy(x1.v, x1q.v, 32);
      const H4 that = reduce(ld(pr + o.That)), taux = reduce(ld(pr + o.Taux)), mu = reduce(ld(pr + o.Mu));
      const H4 pa = reduce(ld(pr + o.a)), pb = reduce(ld(pr + o.b));
      auto inv_h4 = [](const H4& v) { Fq t; memcpy(t.v, v.v, 32); Fq r = fq_is_zero(t) ? fq_zero() : fq_inv_host(t); H4 o4; memcpy(o4.v, r.v, 32); return o4; };
      const H4 zM = to_m(z), xM = to_m(x), yM = to_m(y), yinvM = to_m(inv_h4(y)), paM = to_m(pa), pbM = to_m(pb);
      // running powers in Montgomery form: yi = y^-i R (kept: E2 and E4 both need it), sum of y^i for delta
      std::vector<H4> yiM(N), s(N);
      {
        H4 acc = to_m(one()), ypow = acc, sumM = zero();
        for (size_t i = 0; i < N; i++) { yiM[i] = acc; acc = mont(acc, yinvM); sumM = add(sumM, ypow); ypow = mont(ypow, yM); }
        // delta = (z - z^2) sum y^i - sum_{j=1..m} z^(j+2) (2^n - 1)                        rangeproof_aggreg_verifier.py:72-79
        const H4 z2 = mul_sm(z, zM);
        H4 two_n = one();
        for (size_t i = 0; i < n; i++) two_n = add(two_n, two_n);
        const H4 tn1M = to_m(sub(two_n, one()));
        H4 delta = mul_sm(sub(z, z2), sumM), zj = mul_sm(z2, zM);
        for (size_t j = 1; j <= m; j++) { delta = sub(delta, mul_sm(zj, tn1M)); zj = mul_sm(zj, zM); }
        st(tsc, sub(that, delta)); tidx[0] = rG;                                          // E1: (t_hat - delta) g
      }
      st(tsc + 32, taux); tidx[1] = rH;                                                   //     taux h
      size_t t = 2;
      auto put = [&](u32 row, const H4& v) { st(tsc + 32 * t, v); tidx[t] = row; t++; };
      put(rGsum, sub(zero(), z));                                                         // E2: -z Gsum
      {
        H4 zjM = mont(zM, zM);                                                            // z^(2+j) R
        for (size_t j = 0; j < m; j++) {
          H4 tw = one();
          for (size_t i = 0; i < n; i++) {                                                //     (z + z^(2+j) 2^i y^-(jn+i)) hs
            const H4 zz = mul_sm(i < 64 ? two_pow[i] : tw, zjM);
            put((u32)(N + j * n + i), add(z, mul_sm(zz, yiM[j * n + i])));
            tw = add(tw, tw);
          }
          zjM = mont(zjM, zM);
        }
      }
      put(rH, sub(zero(), mu)); put(rU, mul_sm(x1, to_m(that)));
      put(rU, x1);                                                                        // E3
      // s vector (Verifier2.get_ss, inner_product_verifier.py:91-102): bit (L-1-j) of i picks x_j, else x_j^-1
      H4 xs2[16], xi2[16], xs2M[16];
      {
        H4 s0 = one();
        for (u32 j = 0; j < L; j++) {
          const H4 xj = reduce(ld(pr + o.Xs + 32 * j)), xji = inv_h4(xj), xjiM = to_m(xji);
          xs2[j] = mul_sm(xj, to_m(xj)); xi2[j] = mul_sm(xji, xjiM); xs2M[j] = to_m(xs2[j]);
          s0 = mul_sm(s0, xjiM);
        }
        s[0] = s0;
        for (size_t i = 1; i < N; i++) {
          u32 k = 0; while (((size_t)2 << k) <= i) k++;           // top set bit of i
          s[i] = mul_sm(s[i - ((size_t)1 << k)], xs2M[L - 1 - k]);   // x^-1 -> x at that position: times x^2
        }
      }
      for (size_t i = 0; i < N; i++) put((u32)i, mul_sm(s[i], paM));                      // E4: a s_i
      for (size_t i = 0; i < N; i++) put((u32)(N + i), mul_sm(mul_sm(s[N - 1 - i], pbM), yiM[i]));      //     b s_i^-1 y^-i
      // ---- proof-specific terms
      size_t v = 0;
      auto putv = [&](const uint8_t* pt, const H4& sc) { memcpy(vpt + 64 * v, pt, 64); st(vsc + 32 * v, sc); v++; };
      const H4 minus1 = sub(zero(), one());
      {
        H4 zp = mul_sm(z, zM);                                                            // E1: - z^(2+j) V_j - x T1 - x^2 T2
        for (size_t j = 0; j < m; j++) { putv(pr + o.V + 64 * j, sub(zero(), zp)); zp = mul_sm(zp, zM); }
        putv(pr + o.T1, sub(zero(), x)); putv(pr + o.T2, sub(zero(), mul_sm(x, xM)));
      }
      putv(pr + o.A, one()); putv(pr + o.S, x); putv(pr + o.Pnew, minus1);                // E2
      putv(pr + o.Unew, minus1);                                                          // E3
      putv(pr + o.Unew, mul_sm(pa, pbM)); putv(pr + o.Pnew, minus1);                      // E4
      for (u32 j = 0; j < L; j++) putv(pr + o.Ls + 64 * j, sub(zero(), xs2[j]));
      for (u32 j = 0; j < L; j++) putv(pr + o.Rs + 64 * j, sub(zero(), xi2[j]));
    }
    for (size_t pi = 0; pi < cn; pi++) {                         // offsets of the 4 MSMs of every proof (index 4p + e)
      u32* to = hs_.toff + 4 * pi; u32* vo = hs_.voff + 4 * pi;
      const u32 tb = (u32)(pi * TT), vb = (u32)(pi * VT);
      to[0] = tb; to[1] = tb + 2; to[2] = tb + 2 + (u32)N + 3; to[3] = to[2] + 1;
      vo[0] = vb; vo[1] = vb + (u32)m + 2; vo[2] = vo[1] + 3; vo[3] = vo[2] + 1;
    }
    hs_.toff[4 * cn] = (u32)(cn * TT); hs_.voff[4 * cn] = (u32)(cn * VT);
    const double h_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_h0).count();
    host_ms += h_ms;
    // ---- device (buffer set `cur`)
    Fq* d_tsc = d_sc2 + (size_t)cur * CH * (TT + VT); Fq* d_vsc = d_tsc + CH * TT;
    u32* d_tidx = d_tidx2 + (size_t)cur * CH * TT;
    u32* d_off = d_off2 + (size_t)cur * 2 * (4 * CH + 1); u32* d_voff = d_off + (4 * CH + 1);
    Affine* d_vpt = rows + F + (size_t)cur * CH * VT;
    XYZZ* d_tabres = d_res2 + (size_t)cur * 8 * CH; XYZZ* d_varres = d_tabres + 4 * CH;
    uint8_t* d_acc = d_acc2 + (size_t)cur * 2 * CH; uint8_t* d_bad = d_acc + CH;
    // uploads on the copy stream: they run under the previous chunk's kernels (buffer set `cur` was released two chunks ago)
    BP_CUDA(cudaMemcpyAsync(d_tsc, hs_.tsc, cn * TT * 32, cudaMemcpyHostToDevice, g.copy_stream));
    BP_CUDA(cudaMemcpyAsync(d_tidx, hs_.tidx, cn * TT * 4, cudaMemcpyHostToDevice, g.copy_stream));
    BP_CUDA(cudaMemcpyAsync(d_off, hs_.toff, (4 * cn + 1) * 4, cudaMemcpyHostToDevice, g.copy_stream));
    BP_CUDA(cudaMemcpyAsync(d_voff, hs_.voff, (4 * cn + 1) * 4, cudaMemcpyHostToDevice, g.copy_stream));
    BP_CUDA(cudaMemcpyAsync(d_vpt, hs_.vpt, cn * VT * 64, cudaMemcpyHostToDevice, g.copy_stream));
    BP_CUDA(cudaMemcpyAsync(d_vsc, hs_.vsc, cn * VT * 32, cudaMemcpyHostToDevice, g.copy_stream));
    BP_CUDA(cudaEventRecord(g.ev_half[cur], g.copy_stream));
    BP_CUDA(cudaStreamWaitEvent(g.stream, g.ev_half[cur], 0));
    BP_CUDA(cudaMemsetAsync(d_bad, 0, cn, g.stream));
    ++g.nlaunch, k_rp_check_points<<<(unsigned)((cn * VT + 127) / 128), 128, 0, g.stream>>>(d_vpt, (u32)VT, (u32)cn, d_bad);
    if (tab ? fb_msm_run_slices(tab, d_tidx, d_tsc, d_off, (u32)(4 * cn), 2 * N, d_tabres)
            : msm_run(rows, d_tidx, d_tsc, (u32)(cn * TT), d_off, (u32)(4 * cn), TT / 4, nullptr, d_tabres)) return 1;
    if (msm_run(d_vpt, nullptr, d_vsc, (u32)(cn * VT), d_voff, (u32)(4 * cn), (VT + 3) / 4, nullptr, d_varres)) return 1;
    ++g.nlaunch, k_rp_aggr_accept<<<(unsigned)((cn + 127) / 128), 128, 0, g.stream>>>(d_tabres, d_varres, (u32)cn, d_bad, d_acc);
    BP_CUDA(cudaGetLastError());
    BP_CUDA(cudaMemcpyAsync(hs_.dev, d_acc, cn, cudaMemcpyDeviceToHost, g.stream));
    BP_CUDA(cudaEventRecord(g.stage_ev[cur], g.stream));
    if (timing) fprintf(stderr, "aggreg chunk of %zu (N = %zu, table %s): host scalars %.3f ms (%u threads)\n", cn, N, tab ? "yes" : "no", h_ms, nthreads);
    if (prev_cn && finish(cur ^ 1, prev_lo, prev_cn)) return 1;  // the chunk before this one (its buffer set is the next to be refilled)
    prev_lo = lo; prev_cn = cn;
  }
  if (prev_cn && finish(cur ^ 1, prev_lo, prev_cn)) return 1;
  if (timing) fprintf(stderr, "aggreg batch of %zu: wall %.3f ms, host scalars %.3f ms in total\n", nproofs,
                      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call0).count(), host_ms);
  return 0;
}

}  // extern "C"
