// bp_fixed.inl -- host side of the fixed-base tables (fixedbase.cuh): a cache keyed by a hash of the generator bytes.
// Policy (bp_fb_set_mode): 0 = never, 1 = build a table the SECOND time the same point set is seen (one-off MSMs never
// pay the ~2 ms build), 2 = build at first sight.  Tables are evicted least-recently-used past a byte budget.
// (included inside namespace bp by bp_gpu.cu)

// `src` = the generator bytes the table was built from: a cache hit is confirmed by comparing them (a 64-bit hash alone
// would let two different generator sets share a table)
struct FbEntry { Affine* tab; size_t n; size_t bytes; unsigned long long stamp; std::vector<uint8_t> src; const Affine* parent = nullptr; };
struct FbSrc {                           // up to 5 byte ranges that make up a generator set, in hashing order
  const uint8_t* p[5]; size_t len[5]; int cnt;
  uint64_t hash(uint64_t seed) const;
  bool equals(const std::vector<uint8_t>& v) const {
    size_t off = 0;
    for (int i = 0; i < cnt; i++) { if (off + len[i] > v.size() || memcmp(v.data() + off, p[i], len[i]) != 0) return false; off += len[i]; }
    return off == v.size();
  }
  std::vector<uint8_t> copy() const {
    std::vector<uint8_t> v;
    for (int i = 0; i < cnt; i++) v.insert(v.end(), p[i], p[i] + len[i]);
    return v;
  }
};
struct FbCache {
  std::map<uint64_t, FbEntry> tabs;
  std::map<uint64_t, unsigned> seen, seen16;
  std::map<uint64_t, FbEntry> tabs16;       // 16-bit-window tables (batch verifier), built from the byte tables
  size_t bytes = 0, cap = (size_t)24 << 30;
  size_t max_points = 4200;        // largest point set that gets a table (2049 of a 1024-wide IPA, 2 * 2048 + 1 commitments)
  int mode = 1;
  unsigned long long clock = 0, hits = 0, builds = 0;
  DevBuf scratch, blockpart, tickets;
};
static FbCache fb;

// a forced window width (bp_msm_set_window) is an explicit request for the bucket method
static bool fb_enabled() { return fb.mode != 0 && g.force_c == 0; }

static uint64_t fb_hash(uint64_t h, const uint8_t* p, size_t nbytes) {
  // 64-bit multiply-xorshift over 8-byte words: only the cache INDEX -- a hit is confirmed against the stored generator bytes
  size_t i = 0;
  for (; i + 8 <= nbytes; i += 8) {
    uint64_t w; memcpy(&w, p + i, 8);
    h = (h ^ w) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
  }
  for (; i < nbytes; i++) { h = (h ^ p[i]) * 0x100000001B3ull; }
  return h;
}

uint64_t FbSrc::hash(uint64_t seed) const { uint64_t h = seed; for (int i = 0; i < cnt; i++) h = fb_hash(h, p[i], len[i]); return h; }

static void fb_evict_for(size_t need) {
  while (!fb.tabs.empty() && fb.bytes + need > fb.cap) {
    auto victim = fb.tabs.begin();
    for (auto it = fb.tabs.begin(); it != fb.tabs.end(); ++it) if (it->second.stamp < victim->second.stamp) victim = it;
    cudaStreamSynchronize(g.stream);
    for (auto it = fb.tabs16.begin(); it != fb.tabs16.end();) {      // a 16-bit table is only valid next to the byte table it was built from
      if (it->second.parent == victim->second.tab) { cudaFree(it->second.tab); fb.bytes -= it->second.bytes; it = fb.tabs16.erase(it); }
      else ++it;
    }
    cudaFree(victim->second.tab);
    fb.bytes -= victim->second.bytes;
    fb.tabs.erase(victim);
    alloc_generation()++;            // captured graphs may hold the freed pointer
  }
}

// Table for the point set `key` (n points at d_pts, readable in g.stream order), or nullptr when the set has no table
// (mode, size, first sighting, or out of memory: the caller then takes the bucket method).
static const Affine* fb_get(uint64_t key, const FbSrc& src, const Affine* d_pts, size_t n) {
  if (fb.mode == 0 || n == 0 || n > fb.max_points) return nullptr;
  key ^= (uint64_t)n * 0xD6E8FEB86659FD93ull;
  auto it = fb.tabs.find(key);
  if (it != fb.tabs.end()) {
    if (it->second.n != n || !src.equals(it->second.src)) return nullptr;      // hash collision: this set keeps the bucket method
    it->second.stamp = ++fb.clock; fb.hits++;
    return it->second.tab;
  }
  if (fb.mode == 1) {
    if (fb.seen.size() > 8192) fb.seen.clear();
    if (++fb.seen[key] < 2) return nullptr;
  }
  const size_t bytes = n * (size_t)BP_FB_WINDOWS * BP_FB_ENTRIES * sizeof(Affine);
  if (bytes > fb.cap) return nullptr;
  fb_evict_for(bytes);
  const size_t ngmax = n < BP_FB_BUILD_GENS ? n : BP_FB_BUILD_GENS;
  XYZZ* scratch = (XYZZ*)fb.scratch.ensure(ngmax * BP_FB_WINDOWS * BP_FB_ENTRIES * sizeof(XYZZ));
  if (!scratch) return nullptr;
  Affine* tab = nullptr;
  if (cudaMalloc((void**)&tab, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  for (size_t g0 = 0; g0 < n; g0 += BP_FB_BUILD_GENS) {
    const u32 ng = (u32)(n - g0 < BP_FB_BUILD_GENS ? n - g0 : BP_FB_BUILD_GENS);
    ++g.nlaunch, k_fb_build<<<(ng * BP_FB_WINDOWS + 63) / 64, 64, 0, g.stream>>>(d_pts, (u32)g0, ng, scratch, tab);
  }
  if (cudaGetLastError() != cudaSuccess) { cudaFree(tab); return nullptr; }
  fb.tabs[key] = FbEntry{tab, n, bytes, ++fb.clock, src.copy()};
  fb.bytes += bytes;
  fb.builds++;
  return tab;
}

// 16-bit-window table of the same point set (n <= 512 points: 67 MB each), derived from its byte table `tab8`; built one
// sighting after the byte table in mode 1.  nullptr when absent (the caller keeps using the byte table).
static const Affine* fb_get16(uint64_t key, const Affine* tab8, size_t n) {
  if (!tab8 || n == 0 || n > 512) return nullptr;
  key ^= (uint64_t)n * 0xD6E8FEB86659FD93ull;
  auto it = fb.tabs16.find(key);
  if (it != fb.tabs16.end()) {
    // confirmed through its parent: `tab8` was matched byte for byte against the caller's generators (fb_get), and this entry
    // was built from exactly that table; anything else under the same key is a collision and keeps the byte table
    if (it->second.n != n || it->second.parent != tab8) return nullptr;
    it->second.stamp = ++fb.clock;
    return it->second.tab;
  }
  if (fb.mode == 1) {
    if (fb.seen16.size() > 8192) fb.seen16.clear();
    if (++fb.seen16[key] < 2) return nullptr;
  }
  const size_t bytes = n * (size_t)BP_FB16_WINDOWS * BP_FB16_ENTRIES * sizeof(Affine);
  if (fb.bytes + bytes > fb.cap) return nullptr;                  // never evict byte tables for this
  Affine* tab = nullptr;
  if (cudaMalloc((void**)&tab, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  const size_t threads = n * BP_FB16_WINDOWS * 256;
  ++g.nlaunch, k_fb_build16<<<(unsigned)((threads + 127) / 128), 128, 0, g.stream>>>(tab8, (u32)n, tab);
  if (cudaGetLastError() != cudaSuccess) { cudaFree(tab); return nullptr; }
  fb.tabs16[key] = FbEntry{tab, n, bytes, ++fb.clock, {}, tab8};
  fb.bytes += bytes;
  fb.builds++;
  return tab;
}

// nmsm table MSMs on g.stream: terms of MSM m are [offsets[m], offsets[m+1]) (device array), or [0, single_n) when
// offsets == nullptr (then nmsm must be 1); max_terms bounds the longest MSM.
static int fb_msm_run(const Affine* tab, const u32* d_idx, const Fq* d_sc, const u32* d_offsets, u32 nmsm, size_t max_terms,
                      u32 single_n, Affine* out_affine, XYZZ* out_xyzz) {
  const u32 nbx = (u32)((max_terms + 31) / 32);
  if (nbx == 0 || nmsm == 0) return 0;
  if (nmsm > 65535) return fail("fb_msm_run: too many MSMs in one launch");
  XYZZ* part = (XYZZ*)fb.blockpart.ensure((size_t)nmsm * nbx * sizeof(XYZZ));
  if (!part) return fail("workspace allocation failed");
  ++g.nlaunch, k_fb_msm<<<dim3(nbx, nmsm), 256, 0, g.stream>>>(tab, d_idx, d_sc, d_offsets, single_n, part);
  u32 nq = 8;
  while (nq < nbx && nq < 64) nq <<= 1;
  ++g.nlaunch, k_fb_finish<<<nmsm, 4 * nq, 0, g.stream>>>(part, nbx, out_affine, out_xyzz);
  BP_CUDA(cudaGetLastError());
  return 0;
}

// Throughput form for a batch of table MSMs of very different lengths (aggregated proofs: 2, N + 3, 1 and 2N terms per proof):
// warps over (MSM, 64-term slice) -> lane sums -> slice sums -> MSM sums.  The block-tree form above (k_fb_msm) is a latency
// design: 180 registers, one block per SM, a 9-level cooperative tree per block of 32 terms -- 8.2 ms for 128 aggregated proofs
// against ~2 ms here.
static int fb_msm_run_slices(const Affine* tab, const u32* d_idx, const Fq* d_sc, const u32* d_offsets, u32 nmsm, size_t max_terms, XYZZ* out_xyzz) {
  const u32 slice = 64, nsl = (u32)((max_terms + slice - 1) / slice);
  if (nsl == 0 || nmsm == 0) return 0;
  const size_t nw = (size_t)nmsm * nsl;
  XYZZ* part = (XYZZ*)fb.blockpart.ensure(nw * 33 * sizeof(XYZZ));
  if (!part) return fail("workspace allocation failed");
  XYZZ* slsum = part + nw * 32;
  ++g.nlaunch, k_fb_lookup_slice<<<(unsigned)((nw + 7) / 8), 256, 0, g.stream>>>(tab, d_idx, d_sc, d_offsets, nmsm, nsl, slice, part);
  ++g.nlaunch, k_fb_fold_warp<<<(unsigned)((nw + 3) / 4), 128, 0, g.stream>>>(part, nullptr, (u32)nw, slsum);
  u32 nq = 8;
  while (nq < nsl && nq < 64) nq <<= 1;
  ++g.nlaunch, k_fb_finish<<<nmsm, 4 * nq, 0, g.stream>>>(slsum, nsl, nullptr, out_xyzz);
  BP_CUDA(cudaGetLastError());
  return 0;
}

// The same for a caller that wants the results in HOST memory right away (commitments and statements of a prover: a handful of
// MSMs, each a latency chain): ONE launch -- the last block of every MSM reduces the block sums (ticket) and writes the XYZZ sum
// into mapped pinned memory -- and the host finishes the affine conversion (fp_host.h: ~2 us against a ~45 us lone-thread
// inversion on the device plus a second launch and a copy).
static int fb_msm_run_host(const Affine* tab, const u32* d_idx, const Fq* d_sc, const u32* d_offsets, u32 nmsm, size_t max_terms,
                           u32 single_n, uint8_t* out64) {
  const u32 nbx = (u32)((max_terms + 31) / 32);
  if (nbx == 0 || nmsm == 0) return 0;
  XYZZ* part = (XYZZ*)fb.blockpart.ensure((size_t)nmsm * nbx * sizeof(XYZZ));
  const bool fresh = fb.tickets.p == nullptr || fb.tickets.cap < nmsm * sizeof(u32);
  u32* ticket = (u32*)fb.tickets.ensure(nmsm * sizeof(u32));
  uint8_t* hx = g.pinned2((size_t)nmsm * 128);
  if (!part || !ticket || !hx) return fail("workspace allocation failed");
  if (fresh) BP_CUDA(cudaMemsetAsync(ticket, 0, fb.tickets.cap, g.stream));      // every launch leaves its tickets at zero
  ++g.nlaunch, k_fb_msm<<<dim3(nbx, nmsm), 256, 0, g.stream>>>(tab, d_idx, d_sc, d_offsets, single_n, part, ticket, nullptr, (XYZZ*)hx);
  BP_CUDA(cudaGetLastError());
  BP_CUDA(cudaStreamSynchronize(g.stream));
  for (u32 i = 0; i < nmsm; i += 8) xyzz_to_affine_host(hx + 128 * (size_t)i, nmsm - i < 8 ? nmsm - i : 8, out64 + 64 * (size_t)i);
  return 0;
}

static void fb_release_all() {
  for (auto& kv : fb.tabs) cudaFree(kv.second.tab);
  for (auto& kv : fb.tabs16) cudaFree(kv.second.tab);
  fb.tabs16.clear(); fb.seen16.clear();
  fb.tabs.clear(); fb.seen.clear(); fb.bytes = 0;
  fb.scratch.release(); fb.blockpart.release(); fb.tickets.release();
}

