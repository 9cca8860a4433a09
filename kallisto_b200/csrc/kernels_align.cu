// K1: per-fragment pseudoalignment on the device.
//
//   match_kernel    one thread per fragment (read pair or single read).  Restates
//                   KmerIndex::match (src/KmerIndex.cpp:1698-1940: k-mer iteration, skip-ahead to the
//                   end of the EC block, middle probe, one-step back-off) on the flat 32-byte-slot
//                   table, then the pair combination of MinCollector::intersectKmers /
//                   intersectECs (src/MinCollector.cpp:160-218, 425-496) reduced to its net effect:
//                   the intersection of the distinct non-empty EC sets hit by the two mates.
//                   Fragments whose hits fall in a single EC set, or whose tuple of EC sets has been
//                   seen before (memo tables), are finished here; the others are queued.
//   resolve_kernel  one warp per queued fragment: warp-cooperative sorted-list intersection
//                   (lanes own elements of the smallest set and binary-search the others), the strand
//                   filter of doStrandSpecificity (src/ProcessReads.cpp:61-124), content-addressed
//                   dictionary insert (ecmapinv semantics) and memo publication.
//
// Per-fragment result = a set handle; per-handle counters (count, first fragment index) replace
// MasterProcessor::update + MinCollector::increaseCount (src/ProcessReads.cpp:424-483,
// src/MinCollector.cpp:251-269): EC ids are assigned afterwards in order of first occurrence,
// which is what the reference produces with -t 1.
#include "kb_device.cuh"
#include "kernels.hpp"

#ifndef KB_MATCH_MIN_BLOCKS
#define KB_MATCH_MIN_BLOCKS 4
#endif
#ifndef KB_LOOKAHEAD
// Experiment (off): inside a run of misses fetch the presence-filter bits of the next KB_LOOKAHEAD k-mers together.
// Measured slower than one k-mer per iteration (2 / 4 / 8: 1.52 / 1.57 / 1.93 ms against 1.44 ms per 2 M pairs,
// profiles/match_lookahead_r02.log): the kernel is bound by instruction issue at 12 of 32 lanes active, and the
// extra hashes cost more than the saved iterations.
#define KB_LOOKAHEAD 0
#endif

namespace kb {

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int32_t ld_relaxed_s32(const int32_t* p) {
  int32_t v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// One 256-bit read-only load (LDG.E.256 on sm_100a): a whole k-mer slot, half a packed read, or a
// block of memo entries per instruction.  The address must be 32-byte aligned.
__device__ __forceinline__ void ld256_nc(const void* p, uint32_t (&w)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}
// The k-mer table probe: one random 32-byte sector out of a multi-GB table, never reused.
#ifndef KB_PROBE_LD
#define KB_PROBE_LD 0
#endif
__device__ __forceinline__ void ld256_probe(const void* p, uint32_t (&w)[8]) {
#if KB_PROBE_LD == 0
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#elif KB_PROBE_LD == 1
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#elif KB_PROBE_LD == 2
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#elif KB_PROBE_LD == 3
  asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%8]; ld.global.nc.v4.b32 {%4,%5,%6,%7}, [%8+16];"
#elif KB_PROBE_LD == 4
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#elif KB_PROBE_LD == 5
  asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#elif KB_PROBE_LD == 6
  asm volatile("ld.global.cg.v4.b32 {%0,%1,%2,%3}, [%8]; ld.global.cg.v4.b32 {%4,%5,%6,%7}, [%8+16];"
#endif
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}
// Per-lane view of the read being matched: 2-bit bases in shared memory as 32-bit words (base i in
// bits 30-2*(i&15) of word i>>4), word w of lane t at [w * stride + t] (bank-conflict free).
// Bases other than A/C/G/T are rare, so the lane keeps only a flag in a register; a read that has
// one consults the invalid-base masks of its packed form in global memory (bit i&31 of word i>>5;
// positions past the end of the read are marked invalid there as well).
struct ReadView {
  const uint32_t* bw;
  const uint32_t* gmask;
  int n_mask;
  int stride;
  int len;
  int k;
  bool has_invalid;

  __device__ __forceinline__ uint64_t kmer(int p) const {
    const int w = p >> 4, s = (p & 15) * 2;
    uint64_t x = (uint64_t)bw[w * stride] << 32;
    if (s + 2 * k > 32) x |= bw[(w + 1) * stride];
    x <<= s;
    if (s + 2 * k > 64) x |= (uint64_t)(bw[(w + 2) * stride] >> (32 - s));   // s > 0 here: 2k <= 62
    return x >> (64 - 2 * k);
  }
  // first start position >= p whose k-window holds only A/C/G/T, or -1
  // (KmerIterator::operator++ / operator+=, ext/bifrost/src/KmerIterator.cpp:6-63)
  __device__ __forceinline__ int next_valid(int p) const {
    if (!has_invalid) return p <= len - k ? p : -1;
    const uint64_t wmask = (1ULL << k) - 1;
    while (p <= len - k) {
      const int w = p >> 5, s = p & 31;
      const uint64_t lo = __ldg(gmask + w);
      const uint64_t hi = (w + 1 < n_mask) ? __ldg(gmask + w + 1) : 0xFFFFFFFFu;
      const uint64_t x = (((hi << 32) | lo) >> s) & wmask;      // s + k <= 63
      if (x == 0) return p;
      p += 64 - __clzll((long long)x);
    }
    return -1;
  }
};

__device__ __forceinline__ uint64_t tuple_hash(const uint32_t* w, int n, int stride) {
  uint64_t h = 0x243F6A8885A308D3ULL ^ (uint64_t)n;
  for (int i = 0; i < n; ++i) h = kb_mix64(h ^ ((uint64_t)w[i * stride] + 0x9E3779B97F4A7C15ULL * (i + 1)));
  return h;
}

// Memo lookups used by the resolve kernel (one lane).  Return KB_H_NOTREADY on a miss.
__device__ __forceinline__ int32_t memo2_lookup(const DevDict& dd, uint32_t e0, uint32_t e1) {
  const unsigned long long key = ((unsigned long long)e0 << 32) | e1;
  uint64_t s = kb_mix64(key) & dd.m2_mask;
  for (;;) {
    const unsigned long long kk = __ldcg(&dd.m2[s].key);
    if (kk == key) return ld_relaxed_s32(&dd.m2[s].val);
    if (kk == ~0ULL) return KB_H_NOTREADY;
    s = (s + 1) & dd.m2_mask;
  }
}
__device__ __forceinline__ bool tuple_equal(const DevDict& dd, uint32_t toff, const uint32_t* w, int n, int stride) {
  const uint32_t* t = dd.tpool + toff;
  bool eq = __ldcg(t) == (uint32_t)n;
  for (int i = 0; eq && i < n; ++i) eq = __ldcg(t + 1 + i) == w[i * stride];
  return eq;
}
__device__ __forceinline__ int32_t memon_lookup(const DevDict& dd, const uint32_t* w, int n, int stride) {
  const uint64_t th = tuple_hash(w, n, stride);
  const uint32_t tag = (uint32_t)(th >> 32);
  uint64_t s = th & dd.mn_mask;
  for (;;) {
    const unsigned long long word = ld_acquire_u64(&dd.mn_key[s]);
    if (word == ~0ULL) return KB_H_NOTREADY;
    if ((uint32_t)(word >> 32) == tag && tuple_equal(dd, (uint32_t)word, w, n, stride)) return ld_relaxed_s32(&dd.mn_val[s]);
    s = (s + 1) & dd.mn_mask;
  }
}

// Location of read `ridx` of a batch (mates interleaved in one buffer, or one buffer per mate).
__device__ __forceinline__ void read_span(const BatchArgs& ba, uint32_t ridx, const uint8_t*& base, uint64_t& off, int& len) {
  base = ba.bases;
  const uint32_t* offs = ba.off;
  uint32_t i = ridx;
  if (ba.bases2) {
    i = ridx >> 1;
    if (ridx & 1) { base = ba.bases2; offs = ba.off2; }
  }
  if (offs) {
    const uint32_t o0 = offs[i], o1 = offs[i + 1];
    off = o0;
    len = (int)(o1 - o0);
  } else {
    off = (uint64_t)i * ba.fixed_len;
    len = (int)ba.fixed_len;
  }
  const bool second = ba.paired && (ridx & 1);
  uint32_t st = second ? ba.start2 : ba.start;
  if (ba.notag && ba.notag[ba.paired ? (ridx >> 1) : ridx]) st = second ? ba.alt_start2 : ba.alt_start;
  if (st) {
    off += st;
    len -= (int)st;
    if (len < 0) len = 0;
  }
  if (ba.skip && ba.skip[ba.paired ? (ridx >> 1) : ridx]) len = 0;
}

// Lane states of match_kernel.
enum : int {
  S_MAIN = 0, S_JUMP = 1, S_MIDDLE = 2, S_BACKOFF = 3,   // k-mer table lookups (KmerIndex::match control flow)
  S_FIN = 4,                                             // fragment finished, waiting for the next service round
  S_EMPTY = 5                                            // no fragment assigned
};

}  // namespace

// ---------------------------------------------------------------------------------------------
// pack_kernel: ASCII reads -> 2-bit bases + invalid-base masks, one thread per (read, 32-base word).
// Streaming and convergent: nine aligned 32-bit loads per thread (whatever the byte offset of the read; never past
// the word that holds the read's last base), then 4 bases at a time with SWAR arithmetic
//   idx   = (c >> 1) & 3                      A 0, C 1, T 2, G 3
//   code  = idx ^ (idx >> 1)                  == Kmer::set_kmer (Kmer.cpp:92-107) on A/C/G/T
//   valid = (c & 0xDF) == "ACTG"[idx]         == isDNA(c & 0xDF)  (KmerIterator.cpp:12-14); the four expected letters
//                                             of a group come from ONE byte permute of the constant "ACTG" with the
//                                             indices as selector, so a group costs 13 instructions; the per-base mask
//                                             is only assembled for words that hold a non-ACGT letter or the read's end
// Packed read = nb 64-bit base words, then nb 32-bit invalid masks, padded (with "invalid") to a
// multiple of 32 bytes.  Bases past the end of the read are packed as A and marked invalid.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_kernel(BatchArgs ba, uint32_t n_reads, uint32_t* out) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;     // launch_pseudoalign: n_reads * nb < 2^32
  const uint32_t nb = ba.nb;
  if (gid >= n_reads * nb) return;
  const uint32_t r = gid / nb, w = gid - r * nb;
  uint64_t off;
  int len;
  const uint8_t* src_bases;
  read_span(ba, r, src_bases, off, len);
  const int base = (int)w * 32;
  uint32_t hi = 0, lo = 0, inv32 = ~0u;
  if (base < len) {
    const int n = min(32, len - base);
    const uint64_t a = off + (uint64_t)base;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(src_bases + (a & ~3ull));
    const int sh = (int)(a & 3) * 8;
    const int last = ((int)(a & 3) + n - 1) >> 2;       // index of the aligned word holding base n-1
    uint32_t wd[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) wd[q] = __ldg(wp + min(q, last));
    uint32_t d[8], pr[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t four = __funnelshift_r(wd[q], wd[q + 1], sh);
      uint32_t x = (four >> 1) & 0x03030303u;
      const uint32_t y = (x | (x >> 4)) & 0x00330033u;                  // index nibbles of bytes 0,1 and of bytes 2,3
      d[q] = (four & 0xDFDFDFDFu) ^ __byte_perm(0x47544341u, 0u, y | (y >> 8));   // 0 where the letter is the expected one
      x ^= (x >> 1) & 0x01010101u;
      pr[q] = x * 0x40100401u;                                          // byte 3 = b0<<6 | b1<<4 | b2<<2 | b3
    }
    hi = __byte_perm(__byte_perm(pr[3], pr[2], 0x0073u), __byte_perm(pr[1], pr[0], 0x0073u), 0x5410u);
    lo = __byte_perm(__byte_perm(pr[7], pr[6], 0x0073u), __byte_perm(pr[5], pr[4], 0x0073u), 0x5410u);
    const bool full = (n == 32);
    uint32_t any = full ? (d[0] | d[1] | d[2] | d[3] | d[4] | d[5] | d[6] | d[7]) : 1u;
    inv32 = 0;
    if (any) {
      const int ng = (n + 3) >> 2;    // a warp whose only lanes here are read ends (4 bases of a 100-base read) leaves after one group
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q >= ng) break;
        uint32_t t = (d[q] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
        t = (t | d[q]) & 0x80808080u;                                   // bit 7 of a byte = letter differs
        inv32 |= (((t >> 7) * 0x01020408u) >> 24) << (4 * q);           // bit j = base j invalid
      }
    }
    if (!full) {                      // n in [1, 31]: bases n.. are not part of the read
      inv32 |= ~0u << n;
      if (n > 16) lo &= ~0u << (64 - 2 * n);
      else { lo = 0; if (n < 16) hi &= ~0u << (32 - 2 * n); }
    }
  }
  uint32_t* dst = out + (size_t)r * ba.pstride;
  reinterpret_cast<uint2*>(dst)[w] = make_uint2(lo, hi);
  dst[2 * nb + w] = inv32;
  for (uint32_t j = 3 * nb + w; j < ba.pstride; j += nb) dst[j] = ~0u;   // padding reads as "invalid"
}

// ---------------------------------------------------------------------------------------------
// dlist_scan_kernel: the D-list rule of KmerIndex::match (src/KmerIndex.cpp:1818-1826, 1928-1939).  The reference
// appends a hit on the dummy unitig -- whose equivalence class is the single off-list target -- to a read's hit list
// when any of the read's k-mers is a distinguishing flanking k-mer; the intersection with the on-list targets
// (ProcessReads.cpp:1072) is then empty.  Net effect with default flags: a fragment holding a D-list k-mer in either
// mate is not pseudoaligned.  One thread per (read, window of 32 k-mer start positions) over the packed 2-bit reads;
// a hit marks the fragment in the skip array match_kernel honours.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dlist_scan_kernel(DevIndex ix, BatchArgs ba, uint32_t n_reads) {
  const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nb = ba.nb;
  if (gid >= (uint64_t)n_reads * nb) return;
  const uint32_t r = (uint32_t)(gid / nb), w = (uint32_t)(gid % nb);
  const uint32_t frag = ba.paired ? (r >> 1) : r;
  uint64_t off;
  int len;
  const uint8_t* unused;
  read_span(ba, r, unused, off, len);
  const int k = ix.k;
  const int p0 = (int)w * 32;
  if (len < k || p0 > len - k) return;
  const uint32_t* pk = ba.packed + (size_t)r * ba.pstride;
  const unsigned long long* bw = reinterpret_cast<const unsigned long long*>(pk);
  const unsigned long long w0 = bw[w], w1 = (w + 1 < nb) ? bw[w + 1] : 0ull;
  const unsigned long long inv = (unsigned long long)pk[2 * nb + w] | ((w + 1 < nb) ? ((unsigned long long)pk[2 * nb + w + 1] << 32) : 0xFFFFFFFF00000000ull);
  const unsigned long long wmask = (1ull << k) - 1;
  for (int j = 0; j < 32; ++j) {
    const int p = p0 + j;
    if (p > len - k) break;
    if ((inv >> j) & wmask) continue;                       // a base other than A/C/G/T in the window
    unsigned long long x = w0 << (2 * j);
    if (j) x |= w1 >> (64 - 2 * j);
    const uint64_t fwd = x >> (64 - 2 * k);
    const uint64_t rc = kb_revcomp(fwd, k);
    const uint64_t canon = fwd < rc ? fwd : rc;
    uint64_t h = kb_mix64(canon) & ix.dfk_mask;
    for (;;) {
      const unsigned long long key = __ldg(ix.dfk + h);
      if (key == canon) { ba.skip_w[frag] = 1; return; }
      if (key == KB_EMPTY_KEY) break;
      h = (h + 1) & ix.dfk_mask;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// match_kernel: persistent warps, 32 independent fragment state machines per warp.
//
// KmerIndex::match (src/KmerIndex.cpp:1698-1940, default flags, empty D-list) is a chain of
// dependent k-mer lookups whose length varies from 2 (clean read inside one EC block) to >100
// (unmappable read: every k-mer is probed).  As straight-line per-thread code a warp pays the
// maximum over its lanes, and lanes sitting at different call sites serialise.  Here each lane
// keeps an explicit state and every iteration of the warp's loop performs exactly ONE lookup per
// active lane through a single convergent site: canonical k-mer + hash, one LDG.E.256 of the
// 32-byte slot, then the reference's control flow as a state transition
//   MAIN     the k-mer at p.  Miss: next valid k-mer.  Hit: record it, distance to the end of its EC
//            block (1780-1788); if >= 2 go to JUMP.
//   JUMP     the jump target (1793-1827).  Absent or same (unitig, EC set): accepted, scanning resumes
//            after the target.  Otherwise MIDDLE (dist > 4) or BACKOFF.
//   MIDDLE   the middle k-mer (1831-1873).
//   BACKOFF  the k-mer after p, once (1876-1925: the outer nextPos is never updated, so the back-off
//            loop runs a single iteration), then MAIN.
// (linear-probing collisions cost one more iteration in the same state).  Lanes whose fragment is
// finished wait until `refill_min` of them can be served together: the rare, expensive steps --
// pair combination, memo lookup, accounting, loading the next packed reads -- then run convergent
// over many lanes instead of once per lane.  The lookups executed are exactly the reference's,
// in the same order per read.  `partial` (single-end early exit) does not change the result and
// is not modelled; the pushes of the anchor hit at synthetic positions (1820, 1824) add no new EC
// set and are dropped.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, KB_MATCH_MIN_BLOCKS) match_kernel(DevIndex ix, DevDict dd, BatchArgs ba) {
  extern __shared__ uint32_t smem[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const unsigned lane = tid & 31;
  const int nb = (int)ba.nb, nw = 2 * nb;
  const int k = ix.k;
  // shared memory per lane (32-bit words, word w of lane t at [w * nt + t]): the tuple of EC-set
  // handles, 3 words of first-mate information, the 2-bit bases of both mates.  When a fragment is
  // finalised the handle tuple may grow by up to 6 words over the two regions that follow it; their
  // contents are in registers by then.
  uint32_t* elist = smem + tid;                                  // [KB_MAX_E]
  uint32_t* msave = elist + (size_t)KB_MAX_E * nt;               // [3]: block, offset | strand << 31, read position
  uint32_t* s_bw = msave + (size_t)3 * nt;                       // [mate][nw]
  uint32_t* spill = ba.spill + ((size_t)blockIdx.x * nt + tid) * KB_SPILL;   // handles beyond KB_MAX_E (global memory)
  static_assert(KB_MAX_E + 6 <= KB_MAX_E + 3 + 4, "tuple extension must fit in the words that follow the handle list");

  // contiguous chunk of fragments owned by this warp
  const uint32_t n_warps = (gridDim.x * nt) >> 5;
  const uint32_t gw = (blockIdx.x * nt + tid) >> 5;
  const uint32_t chunk = (ba.n_frag + n_warps - 1) / n_warps;
  uint32_t next = min(ba.n_frag, gw * chunk);
  const uint32_t end = min(ba.n_frag, next + chunk);
  const int n_mates = ba.paired ? 2 : 1;
  const int n_chunks = (int)(ba.pstride >> 3);   // 32-byte pieces per packed read

  int st = S_EMPTY;
  uint32_t frag = 0;
  int mate = 0;
  int p = -1, p2 = -1, p3 = -1, np = 0, dist = 0;   // np = the reference's nextPos
  int len1 = 0;
  uint32_t hu = 0, he = 0, h2u = 0, h2e = 0;
  int n_e = 0;
  bool overflow = false, need_prep = false;
#if KB_LOOKAHEAD > 1
  bool in_run = false;      // the last MAIN lookup of this lane missed: the scan is inside a run of misses
#endif
  // first hit of the mate being matched (findFirstMappingKmer / mapPair) and hit flags of both mates
  bool v_cur = false, s_cur = false, v_first = false, s_first = false, f_strand = false;
  uint32_t f_blk = 0, f_dist = 0;
  int f_pos = 0;
  unsigned inv_flags = 0;    // bit m: mate m holds a base other than A/C/G/T
  uint64_t canon = 0, slot = 0;
  bool is_canon = false;
  uint32_t n_probes = 0, n_visits = 0, n_memo = 0;   // per-lane totals: touched once per fragment (they live in local memory)
  uint32_t pv = 0;          // hot-loop counter of the current fragment: lookups in the low half, slot visits in the high half
  ReadView rv;
  rv.stride = nt;
  rv.k = k;
  rv.len = 0;
  rv.bw = s_bw;
  rv.gmask = nullptr;
  rv.n_mask = nb;
  rv.has_invalid = false;
  // point the view at mate `mt` of the lane's fragment
  auto set_mate = [&](int mt, int len_mt) {
    rv.bw = s_bw + (size_t)mt * nw * nt;
    rv.len = len_mt;
    rv.has_invalid = ((inv_flags >> mt) & 1u) != 0;
    rv.gmask = ba.packed + (size_t)(ba.paired ? 2 * frag + mt : frag) * ba.pstride + nw;
  };

  for (;;) {
    // ------------------------------------------------------------------ service round
    const unsigned fin = __ballot_sync(0xFFFFFFFFu, st == S_FIN);
    const unsigned idle = fin | __ballot_sync(0xFFFFFFFFu, st == S_EMPTY);
    const bool work_left = next < end;
    if (idle == 0xFFFFFFFFu && fin == 0 && !work_left) break;
    if (idle == 0xFFFFFFFFu || (work_left && __popc(idle) >= ba.refill_min)) {
      if (st == S_FIN) {
        n_probes += pv & 0xFFFFu;     // a fragment executes at most a few hundred lookups
        n_visits += pv >> 16;
        pv = 0;
        // ---- MinCollector::intersectKmers, net effect (MinCollector.cpp:160-218) ----
        bool v0 = v_cur, s0 = s_cur, v1 = false, s1 = false;
        if (mate == 1) { v0 = v_first; s0 = s_first; v1 = v_cur; s1 = s_cur; }
        // first hit of the first mate (written at the mate switch), before the tuple may grow over it
        const uint32_t m_blk = msave[0], m_ds = msave[nt], m_pos = msave[2 * nt];
        bool mapped = v0 || v1;
        if ((v0 && !s0) || (v1 && !s1)) mapped = false;
        if (mapped && n_e == 0) mapped = false;
        int32_t handle = KB_H_UNMAPPED;
        if (mapped) {
          // single-end reads / pairs with one mate mapped, known mean fragment length: the transcripts
          // whose ends the fragment would overhang are filtered per fragment (ProcessReads.cpp:1095-1136)
          const bool want_fp = ba.fp_fl >= 0 && (!ba.paired || !v0 || !v1);
          // words that follow the set handles in the tuple: (block, orientation) of each mate's first hit for the
          // strand filter; block, orientation, read position and unitig offset of the mapped mate's first hit for
          // the position filter
          const bool stranded = ba.strand_mode != 0;
          uint32_t xs0 = 0xFFFFFFFFu, xs1 = 0xFFFFFFFFu;
          if (stranded) {
            if (mate == 1) {
              xs0 = v_first ? (m_blk * 2u + (m_ds >> 31)) : 0xFFFFFFFFu;
              xs1 = v_cur ? (f_blk * 2u + (f_strand ? 1u : 0u)) : 0xFFFFFFFFu;
            } else {
              xs0 = v_cur ? (f_blk * 2u + (f_strand ? 1u : 0u)) : 0xFFFFFFFFu;
            }
            // a fragment without the UMI tag is not strand-filtered (doStrandSpecificityIfPossible = false,
            // ProcessReads.cpp:1526): "no first hit" for both mates makes resolve_kernel skip the filter
            if (ba.notag && ba.notag[frag]) xs0 = xs1 = 0xFFFFFFFFu;
          }
          const bool use_first = (mate == 1) && !v_cur;     // second mate empty: the first mate's hit
          const uint32_t xf0 = use_first ? m_blk : f_blk;
          const uint32_t xf1 = use_first ? (m_ds >> 31) : (f_strand ? 1u : 0u);
          const uint32_t xf2 = use_first ? m_pos : (uint32_t)f_pos;
          const uint32_t xf3 = use_first ? (m_ds & 0x7FFFFFFFu) : f_dist;
          const int nx = (stranded ? 2 : 0) + (want_fp ? 4 : 0);
          // appends the extra words to a tuple stored with stride `st` starting at index `at`
          auto put_extras = [&](uint32_t* dst, int at, int st) {
            if (stranded) { dst[at * st] = xs0; dst[(at + 1) * st] = xs1; at += 2; }
            if (want_fp) { dst[at * st] = xf0; dst[(at + 1) * st] = xf1; dst[(at + 2) * st] = xf2; dst[(at + 3) * st] = xf3; }
          };
          if (overflow) {
            atomicOr(dd.error, KB_DEVERR_E_OVERFLOW);
          } else if (n_e > KB_MAX_E) {
            // ---- rare: more distinct EC sets than the shared-memory tuple holds (reads crossing many short EC
            //      blocks).  The tail of the list lives in this lane's spill area in global memory; the tuple is
            //      sorted through an accessor and handed to the resolve kernel through the wide queue, unmemoised.
            auto E = [&](int i) -> uint32_t { return i < KB_MAX_E ? elist[i * nt] : spill[i - KB_MAX_E]; };
            auto S = [&](int i, uint32_t x) { if (i < KB_MAX_E) elist[i * nt] = x; else spill[i - KB_MAX_E] = x; };
            for (int i = 1; i < n_e; ++i) {
              const uint32_t x = E(i);
              int j = i - 1;
              while (j >= 0 && E(j) > x) { S(j + 1, E(j)); --j; }
              S(j + 1, x);
            }
            const uint32_t q = atomicAdd(ba.qbig_count, 1u);
            if (q >= ba.qbig_cap) {
              atomicOr(dd.error, KB_DEVERR_E_OVERFLOW);
            } else {
              uint32_t* e = ba.qbig_entries + (size_t)q * KB_QBIG_STRIDE;
              e[0] = frag;
              e[1] = (uint32_t)(n_e + nx) | (want_fp ? 0x80000000u : 0u) | 0x40000000u;   // bit 30: never memoised
              for (int i = 0; i < n_e; ++i) e[2 + i] = E(i);
              put_extras(e + 2, n_e, 1);
              handle = KB_H_PENDING;
            }
          } else {
            for (int i = 1; i < n_e; ++i) {   // sort the distinct set handles (<= 16 entries)
              const uint32_t x = elist[i * nt];
              int j = i - 1;
              while (j >= 0 && elist[j * nt] > x) { elist[(j + 1) * nt] = elist[j * nt]; --j; }
              elist[(j + 1) * nt] = x;
            }
            if (ba.strand_mode == 0 && n_e == 1 && !want_fp) {
              handle = (int32_t)elist[0];            // a single EC set: its handle is stored in the slot
            } else {
              const int n = n_e + nx;
              int32_t r;
              if (ba.strand_mode == 0 && n_e == 2 && !want_fp) {
                r = memo2_lookup(dd, elist[0], elist[nt]);
              } else {
                put_extras(elist, n_e, nt);
                // position-filtered fragments depend on the read itself: never memoised
                r = want_fp ? KB_H_NOTREADY : memon_lookup(dd, elist, n, nt);
              }
              if (r == KB_H_NOTREADY) {
                handle = KB_H_PENDING;
                const uint32_t q = atomicAdd(ba.q_count, 1u);
                uint32_t* e = ba.q_entries + (size_t)q * KB_Q_STRIDE;
                e[0] = frag;
                e[1] = (uint32_t)n | (want_fp ? 0x80000000u : 0u);
                for (int i = 0; i < n; ++i) e[2 + i] = elist[i * nt];
              } else {
                handle = r;
                ++n_memo;
              }
            }
          }
        }
        ba.handle_out[frag] = handle;
        if (ba.tl_out) {
          // KmerIndex::mapPair (KmerIndex.cpp:1622-1693): the first k-mer found by a linear scan is
          // the first hit of match(); same unitig, same EC set, opposite strands, same block end --
          // i.e. the same EC block of the index (blocks tile their unitig and carry one EC set).
          uint16_t tl = 0;
          if (ba.paired && mate == 1 && v_first && v_cur) {
            const uint32_t a_dist = m_ds & 0x7FFFFFFFu;
            const bool a_strand = (m_ds >> 31) != 0;
            const int a_pos = (int)m_pos;
            const int q1 = a_strand ? (int)a_dist - a_pos : (int)a_dist + k + a_pos;
            const int q2 = f_strand ? (int)f_dist - f_pos : (int)f_dist + k + f_pos;
            if (m_blk == f_blk && (a_strand != f_strand)) {
              const int d = q1 > q2 ? q1 - q2 : q2 - q1;
              if (d > 0 && d < 1000) tl = (uint16_t)d;
            }
          }
          if (ba.notag && !ba.notag[frag]) tl = 0;     // tag runs: only fragments without the tag are sampled (getFragLenIfPaired, :1525)
          ba.tl_out[frag] = tl;
        }
        // per-handle accounting, aggregated over the lanes finalised in this round
        const unsigned grp = __match_any_sync(fin, handle);
        const uint32_t fmin = __reduce_min_sync(grp, frag);
        if (handle >= 0 && lane == (unsigned)(__ffs(grp) - 1)) {
          atomicAdd(&dd.count[handle], (uint32_t)__popc(grp));
          atomicMin(&dd.first[handle], (unsigned long long)(ba.frag_base + fmin));
        }
        st = S_EMPTY;
      }
      __syncwarp();
      // ---- refill: the idle lanes take the next fragments of the warp's chunk and copy their packed
      //      reads (pack_kernel output) into shared memory with 256-bit loads
      {
        const bool is_idle = (idle >> lane) & 1u;
        const uint32_t rank = __popc(idle & ((1u << lane) - 1));
        const uint32_t avail = end - next;
        if (is_idle && rank < avail) {
          const uint32_t fidx = next + rank;
          int l0 = 0;
          unsigned inv = 0;
          for (int mt = 0; mt < n_mates; ++mt) {
            const uint32_t ridx = ba.paired ? 2 * fidx + mt : fidx;
            const uint8_t* unused_base;
            uint64_t unused_off;
            int len;
            read_span(ba, ridx, unused_base, unused_off, len);
            if (len > nb * 32) len = nb * 32;   // cannot happen: the host sizes nb from the longest read
            if (mt == 0) l0 = len; else len1 = len;
            const uint32_t* src = ba.packed + (size_t)ridx * ba.pstride;
            uint32_t* dbw = s_bw + (size_t)mt * nw * nt;
            uint32_t bad = 0;
            for (int c = 0; c < n_chunks; ++c) {
              uint32_t v[8];
              ld256_nc(src + c * 8, v);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                // packed stream: nb 64-bit base words (little-endian halves: the high half holds the
                // first 16 bases), then nb 32-bit invalid masks, then padding
                const int g = c * 8 + i;
                if (g < nw) {
                  dbw[(g ^ 1) * nt] = v[i];
                } else if (g < nw + nb) {
                  const int first = (g - nw) * 32;           // mask of bases [first, first + 32)
                  const uint32_t in_read = len >= first + 32 ? 0xFFFFFFFFu : (len > first ? ((1u << (len - first)) - 1u) : 0u);
                  bad |= v[i] & in_read;
                }
              }
            }
            if (bad) inv |= 1u << mt;
          }
          frag = fidx;
          mate = 0;
          n_e = 0;
          overflow = false;
          v_cur = s_cur = v_first = s_first = false;
          inv_flags = inv;
          set_mate(0, l0);
          p = rv.next_valid(0);
          st = S_MAIN;
          need_prep = true;
          if (p < 0) {   // no k-mer in the first mate
            st = S_FIN;
            if (n_mates == 2) {
              mate = 1;
              set_mate(1, len1);
              p = rv.next_valid(0);
              if (p >= 0) st = S_MAIN;
            }
          }
        }
        const uint32_t n_idle = __popc(idle);
        next += n_idle < avail ? n_idle : avail;
      }
      continue;
    }
    // ------------------------------------------------------------------ one lookup per active lane
    if (st <= S_BACKOFF) {
      bool absent = false;
      if (need_prep) {
        bool known_present = false;
#if KB_LOOKAHEAD > 1
        if (st == S_MAIN && in_run && ix.filter && !rv.has_invalid) {
          // Look-ahead of the linear scan (MAIN: a miss just moves on to the next k-mer, KmerIndex.cpp:1750-1753).  69 % of
          // all lookups are such misses and they come in runs (31 k-mers around every sequencing error, whole unmappable
          // reads), so once a scan has missed, the presence filter is consulted for the next KB_LOOKAHEAD k-mers at once --
          // independent L2 loads -- and the scan jumps to the first one that may be in the table.  The k-mers passed over are exactly the ones the
          // reference looks up and misses; they are counted as probes.
          const int last = rv.len - k;
          const int m = min(KB_LOOKAHEAD, last - p + 1);
          uint32_t fw[KB_LOOKAHEAD], fi[KB_LOOKAHEAD];
#pragma unroll
          for (int i = 0; i < KB_LOOKAHEAD; ++i) {
            fw[i] = 0;
            fi[i] = 0;
            if (i < m) {
              const uint64_t f0 = rv.kmer(p + i);
              const uint64_t r0 = kb_revcomp(f0, k);
              const uint64_t h0 = kb_mix64(f0 < r0 ? f0 : r0);
              fi[i] = (uint32_t)(h0 >> 32) & ix.filter_mask;
              fw[i] = __ldg(ix.filter + (fi[i] >> 5));
            }
          }
          int adv = m - 1;                       // all absent: stand on the last one, it is a miss
#pragma unroll
          for (int i = KB_LOOKAHEAD - 1; i >= 0; --i)
            if (i < m && ((fw[i] >> (fi[i] & 31)) & 1u)) { adv = i; known_present = true; }
          // (the loop runs downwards, so adv ends up as the FIRST position whose bit is set)
          p += adv;
          pv += (uint32_t)adv;
          absent = !known_present;
        }
#endif
        const int pq = (st == S_JUMP) ? p2 : ((st == S_MIDDLE) ? p3 : p);
        const uint64_t fwd = rv.kmer(pq);
        const uint64_t rc = kb_revcomp(fwd, k);
        is_canon = fwd < rc;
        canon = is_canon ? fwd : rc;
        const uint64_t hsh = kb_mix64(canon);
        slot = hsh & ix.mask;
        need_prep = false;
        ++pv;
        // presence filter (L2 resident): a clear bit means the k-mer is not in the index -- no HBM sector is touched
        if (ix.filter && !known_present && !absent) {
          const uint32_t fidx = (uint32_t)(hsh >> 32) & ix.filter_mask;
          absent = ((__ldg(ix.filter + (fidx >> 5)) >> (fidx & 31)) & 1u) == 0;
        }
      }
      uint32_t v[8];
      if (!absent) {
        ld256_probe(ix.slots + slot, v);
        pv += 0x10000u;
      } else {
        v[0] = v[1] = 0xFFFFFFFFu;     // reads as an empty slot: a miss
      }
      const uint64_t key = (uint64_t)v[0] | ((uint64_t)v[1] << 32);
      if (key != canon && key != KB_EMPTY_KEY) {
        slot = (slot + 1) & ix.mask;              // linear probing: one more iteration
      } else {
        const bool f = key == canon;
        // hit fields: v[2] unitig, v[3] blk, v[4] ec (set handle), v[5] dist|flag, v[6] lb, v[7] ub
        const uint32_t r_unitig = v[2], r_ec = v[4];
        const bool r_strand = (is_canon == ((v[5] >> 31) != 0));
        const int l = rv.len;
        bool push = false, end_mate = false, to_backoff = false;
        int nv_from = -1;      // >= 0: continue with p = next_valid(nv_from) in MAIN (or BACKOFF)
        if (st == S_MAIN) {
#if KB_LOOKAHEAD > 1
          in_run = !f;
#endif
          if (!f) {
            nv_from = p + 1;
          } else {
            push = true;
            const int r_dist = (int)(v[5] & 0x7FFFFFFFu);
            const int off = r_dist - (int)v[6], blen = (int)(v[7] - v[6]);
            dist = r_strand ? (blen - 1 - off) : off;                       // 1780-1788
            if (dist >= 2) {
              np = (p + dist >= l - k) ? (l - k) : (p + dist);              // 1793-1798
              p2 = rv.next_valid(np);                                       // kit2 += nextPos-pos (adv 0: p itself)
              if (p2 < 0) {
                end_mate = true;                                            // 1882-1886
              } else {
                hu = r_unitig; he = r_ec;
                st = S_JUMP;
                need_prep = true;
              }
            } else {
              nv_from = p + 1;
            }
          }
        } else if (st == S_JUMP) {
          const bool found2 = !f || (hu == r_unitig && he == r_ec);         // 1807-1815
          const int found2pos = !f ? p : p + dist;
          if (found2) {
            if (found2pos >= l - k) end_mate = true;                        // "fake position", break (1819-1822)
            else nv_from = p2 + 1;                                          // kit = kit2; ++kit
          } else {
            h2u = r_unitig; h2e = r_ec;
            if (dist > 4) {
              const int middlePos = (p + np) / 2;
              p3 = rv.next_valid(middlePos);                                // kit3 += middlePos-pos
              if (p3 >= 0) { st = S_MIDDLE; need_prep = true; }
              else to_backoff = true;
            } else {
              to_backoff = true;
            }
          }
        } else if (st == S_MIDDLE) {
          const bool foundMiddle = f && ((hu == r_unitig && he == r_ec) || (h2u == r_unitig && h2e == r_ec));
          if (foundMiddle) {
            push = true;
            if (np >= l - k) end_mate = true;                               // 1867
            else nv_from = p2 + 1;                                          // kit = kit2; ++kit
          } else {
            to_backoff = true;
          }
        } else {   // S_BACKOFF: the single probe of the back-off loop
          push = f;
          nv_from = p + 1;
        }
        if (push) {
          if (!v_cur) {
            v_cur = true;
            f_blk = v[3]; f_dist = v[5] & 0x7FFFFFFFu;
            f_pos = p;   // only a MAIN hit can be the first hit of a read
            f_strand = r_strand;
          }
          if (r_ec != ba.empty_ec) {               // "Don't intersect empty EC", MinCollector.cpp:468-469
            s_cur = true;
            bool dup = false;
            const int n_sh = n_e < KB_MAX_E ? n_e : KB_MAX_E;
            for (int i = 0; i < n_sh; ++i) dup |= (elist[i * nt] == r_ec);
            for (int i = KB_MAX_E; i < n_e; ++i) dup |= (spill[i - KB_MAX_E] == r_ec);     // rare: spilled tail
            if (!dup) {
              if (n_e < KB_MAX_E) { elist[n_e * nt] = r_ec; ++n_e; }
              else if (n_e < KB_MAX_E + KB_SPILL) { spill[n_e - KB_MAX_E] = r_ec; ++n_e; }
              else overflow = true;
            }
          }
        }
        if (to_backoff) nv_from = p + 1;           // ++kit; backOff = true
        if (nv_from >= 0) {
          p = rv.next_valid(nv_from);
          if (p < 0) end_mate = true;
          else { st = to_backoff ? S_BACKOFF : S_MAIN; need_prep = true; }
        }
        if (end_mate) {
          st = S_FIN;
          if (mate + 1 < n_mates) {
            // keep the first mate's flags and first hit, move on to the second mate
            v_first = v_cur; s_first = s_cur;
            msave[0] = f_blk; msave[nt] = f_dist | (f_strand ? 0x80000000u : 0u); msave[2 * nt] = (uint32_t)f_pos;
            v_cur = s_cur = false;
            mate = 1;
            set_mate(1, len1);
            p = rv.next_valid(0);
            if (p >= 0) { st = S_MAIN; need_prep = true; }
          }
        }
      }
    }
  }
  // statistics: probes and slot visits
  n_probes += pv & 0xFFFFu;
  n_visits += pv >> 16;
  for (int o = 16; o > 0; o >>= 1) {
    n_probes += __shfl_xor_sync(0xFFFFFFFFu, n_probes, o);
    n_visits += __shfl_xor_sync(0xFFFFFFFFu, n_visits, o);
    n_memo += __shfl_xor_sync(0xFFFFFFFFu, n_memo, o);
  }
  if (lane == 0 && n_probes) {
    atomicAdd(&dd.stats[0], (unsigned long long)n_probes);
    atomicAdd(&dd.stats[3], (unsigned long long)n_visits);
    if (n_memo) atomicAdd(&dd.stats[2], (unsigned long long)n_memo);
  }
}

namespace {

__device__ __forceinline__ bool bsearch_contains(const uint32_t* s, uint32_t n, uint32_t v, uint32_t* rank) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    const uint32_t x = __ldcg(s + mid);
    if (x < v) lo = mid + 1; else hi = mid;
  }
  if (rank) *rank = lo;
  return lo < n && __ldcg(s + lo) == v;
}

// Lane groups: G consecutive lanes of a warp (G = 32, 16, 8 or 4) work on one item; the groups of a warp are
// independent of each other (every collective below names only the group's lanes).
template <int G>
__device__ __forceinline__ unsigned group_mask(unsigned lane_in_warp) {
  return G == 32 ? 0xFFFFFFFFu : (((1u << (G & 31)) - 1u) << (lane_in_warp & ~(unsigned)(G - 1)));
}

// Group-cooperative lookup-or-insert of a sorted transcript-id list in the content-addressed set
// dictionary (ecmapinv semantics: equal sets share one handle).  `src` may be shared or global memory
// readable by all lanes of the group; `lane` is the lane's index inside its group, `gmask` the group's lanes.
// Returns the handle, or KB_H_UNMAPPED after flagging an error.
template <int G = 32>
__device__ __forceinline__ int32_t dict_insert_warp(const DevDict& dd, const uint32_t* src, uint32_t nres, unsigned lane,
                                                    unsigned gmask = 0xFFFFFFFFu) {
  uint64_t sum = 0;
  for (uint32_t i = lane; i < nres; i += G) sum += kb_mix64((uint64_t)src[i] + 0x9E3779B97F4A7C15ULL);
  for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(gmask, sum, o);
  const uint64_t hsh = kb_mix64(sum ^ nres);
  const unsigned long long tag = hsh >> 56;
  uint64_t s = hsh & dd.dmask;
  unsigned long long my_word = ~0ULL;   // pool space is allocated lazily
  uint64_t visited = 0;
  for (;;) {
    unsigned long long word = 0;
    if (lane == 0) word = ld_acquire_u64(&dd.dslots[s]);
    word = __shfl_sync(gmask, word, 0, G);
    if (word == ~0ULL) {
      if (my_word == ~0ULL) {
        unsigned long long off = 0;
        if (lane == 0) off = atomicAdd(dd.pool_top, (unsigned long long)nres);
        off = __shfl_sync(gmask, off, 0, G);
        if (off + nres > dd.pool_cap || off + nres > 0xFFFFFFFFULL) {
          if (lane == 0) atomicOr(dd.error, KB_DEVERR_POOL_FULL);
          return KB_H_UNMAPPED;
        }
        for (uint32_t i = lane; i < nres; i += G) dd.pool[off + i] = src[i];
        __threadfence();
        __syncwarp(gmask);
        my_word = off | ((unsigned long long)nres << 32) | (tag << 56);
      }
      unsigned long long old = 0;
      if (lane == 0) old = atomicCAS(&dd.dslots[s], ~0ULL, my_word);
      old = __shfl_sync(gmask, old, 0, G);
      if (old == ~0ULL) return (int32_t)s;
      word = old;   // somebody else took the slot: compare against theirs
    }
    if ((word >> 56) == tag && ((word >> 32) & 0xFFFFFFu) == nres) {
      const uint32_t* S = dd.pool + (uint32_t)word;
      bool eq = true;
      for (uint32_t i = lane; i < nres; i += G) eq = eq && (__ldcg(S + i) == src[i]);
      if (__all_sync(gmask, eq)) return (int32_t)s;
    }
    s = (s + 1) & dd.dmask;
    if (++visited > dd.dmask) {
      if (lane == 0) atomicOr(dd.error, KB_DEVERR_DICT_FULL);
      return KB_H_UNMAPPED;
    }
  }
}

}  // namespace

// One group of G lanes per queued fragment (ra.n_warps counts groups).  The kernel is a chain of dependent memory
// accesses per fragment (queue entry -> memo -> set descriptors -> set elements -> binary searches -> dictionary ->
// memo), so what sets its speed is the number of fragments in flight: the sets are short (2.5 ids on average), 8
// lanes hold them, and a warp then carries four fragments instead of one.
template <int G>
__global__ void __launch_bounds__(128, 8) resolve_kernel(DevIndex ix, DevDict dd, BatchArgs ba, ResolveArgs ra) {
  const unsigned lane = threadIdx.x & (G - 1);                          // lane inside its group
  const unsigned gmask = group_mask<G>(threadIdx.x & 31);
  const unsigned gshift = (threadIdx.x & 31) & ~(unsigned)(G - 1);      // the group's first lane of the warp
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / G;    // group index
  if (warp >= ra.n_warps) return;
  uint32_t* scratch = ra.scratch + (size_t)warp * ra.scratch_stride;
  const uint32_t* pool = dd.pool;

  // pass 0: the regular queue; pass 1: the wide queue (fragments with more than KB_MAX_E distinct EC sets)
  for (int pass = 0; pass < 2; ++pass) {
  const uint32_t nq = pass == 0 ? *ba.q_count : min(*ba.qbig_count, ba.qbig_cap);
  const uint32_t* entries = pass == 0 ? ba.q_entries : ba.qbig_entries;
  const size_t stride = pass == 0 ? (size_t)KB_Q_STRIDE : (size_t)KB_QBIG_STRIDE;
  if (warp == 0 && lane == 0 && nq) atomicAdd(&dd.stats[1], (unsigned long long)nq);   // fragments finished here
  for (uint32_t q = warp; q < nq; q += ra.n_warps) {
    const uint32_t* e = entries + (size_t)q * stride;
    const uint32_t f = e[0];
    const bool has_fp = (e[1] >> 31) != 0;
    const bool no_memo = has_fp || ((e[1] >> 30) & 1u) != 0;
    const int n = (int)(e[1] & 0xFFFFu);
    const uint32_t* w = e + 2;
    const bool stranded = ba.strand_mode != 0;
    const int n_e = n - (stranded ? 2 : 0) - (has_fp ? 4 : 0);
    const bool use_m2 = (!stranded && !no_memo && n_e == 2);

    // 1. has somebody else resolved the same tuple in the meantime?
    int32_t handle = KB_H_NOTREADY;
    if (lane == 0 && !no_memo) handle = use_m2 ? memo2_lookup(dd, w[0], w[1]) : memon_lookup(dd, w, n, 1);
    handle = __shfl_sync(gmask, handle, 0, G);

    if (handle == KB_H_NOTREADY) {
      // 2. intersection of the n_e sets: lanes own elements of the smallest one
      // the tuple holds set handles: dslots[h] = offset | len << 32 | tag << 56
      int sm = 0;
      uint32_t sm_len = (uint32_t)((dd.dslots[w[0]] >> 32) & 0xFFFFFFu);
      for (int j = 1; j < n_e; ++j) {
        const uint32_t len = (uint32_t)((dd.dslots[w[j]] >> 32) & 0xFFFFFFu);
        if (len < sm_len) { sm_len = len; sm = j; }
      }
      const uint32_t* A = pool + (uint32_t)dd.dslots[w[sm]];
      uint32_t nres = 0;
      for (uint32_t base = 0; base < sm_len; base += G) {
        const uint32_t i = base + lane;
        bool alive = i < sm_len;
        const uint32_t a = alive ? __ldcg(A + i) : 0;
        for (int j = 0; j < n_e; ++j) {
          if (j == sm) continue;
          const unsigned long long bw_ = dd.dslots[w[j]];
          const uint32_t* B = pool + (uint32_t)bw_;
          const uint32_t blen = (uint32_t)((bw_ >> 32) & 0xFFFFFFu);
          if (alive) alive = bsearch_contains(B, blen, a, nullptr);
        }
        const unsigned bal = (__ballot_sync(gmask, alive) >> gshift);
        if (alive) scratch[nres + __popc(bal & ((1u << lane) - 1))] = a;
        nres += __popc(bal);
      }
      __syncwarp(gmask);
      // 2b. fragment-position filter (ProcessReads.cpp:1095-1136 with KmerIndex::findPosition,
      //     KmerIndex.cpp:2188-2292): keep the transcripts the fragment fits into
      if (has_fp && nres > 0) {
        const uint32_t* fw = w + n - 4;
        const uint32_t blk = fw[0];
        const bool csense = fw[1] != 0;
        const long long pp = (long long)fw[2], udist = (long long)fw[3];
        const long long usize = (long long)ix.blk_usize[blk], kk = (long long)ix.k, fl = (long long)ba.fp_fl;
        const unsigned long long bword = dd.dslots[ix.blk_ec[blk]];
        const uint32_t* B = pool + (uint32_t)bword;
        const uint32_t blen = (uint32_t)((bword >> 32) & 0xFFFFFFu);
        const uint4* info = ix.fp_info + ix.blk_strand_off[blk];
        uint32_t n_v = 0;
        for (uint32_t base = 0; base < nres; base += G) {
          const uint32_t i = base + lane;
          bool keep = false;
          const uint32_t tr = i < nres ? scratch[i] : 0;
          if (i < nres) {
            uint32_t rank = 0;
            if (bsearch_contains(B, blen, tr, &rank)) {
              const uint4 c = info[rank];
              const long long trpos = (long long)(c.x & 0x7FFFFFFFu);
              const bool trsense = (c.x >> 31) == 0;
              long long x;
              bool s;
              if (trsense) {
                if (csense) { x = trpos - pp + udist + 1 - (long long)c.y; s = true; }           // case I
                else { x = trpos + pp + kk + udist - (long long)c.z; s = false; }                // case III
              } else {
                if (csense) { x = trpos - udist + usize - (long long)c.w + pp; s = false; }      // case IV
                else { x = trpos + usize - udist - (long long)c.w - kk + 1 - pp; s = true; }     // case II
              }
              const int xi = (int)x;
              keep = (s && xi + (int)fl <= (int)ix.target_len[tr]) || (!s && xi - (int)fl >= 0);
            }
          }
          const unsigned bv = (__ballot_sync(gmask, keep) >> gshift);
          if (keep) scratch[ra.scratch_stride / 2 + n_v + __popc(bv & ((1u << lane) - 1))] = tr;
          n_v += __popc(bv);
        }
        __syncwarp(gmask);
        if (n_v < nres) {
          for (uint32_t i = lane; i < n_v; i += G) scratch[i] = scratch[ra.scratch_stride / 2 + i];
          nres = n_v;
        }
        __syncwarp(gmask);
      }
      // 3. doStrandSpecificity (ProcessReads.cpp:61-124), first mate then second mate
      if (stranded) {
        for (int mate = 0; mate < 2 && nres > 0; ++mate) {
          const uint32_t sw = w[n_e + mate];
          if (sw == 0xFFFFFFFFu) continue;            // v empty for this mate
          const uint32_t blk = sw >> 1;
          const bool um_strand = (sw & 1) != 0;
          const bool want = (mate == 0) ? (ba.strand_mode == 1) : (ba.strand_mode == 2);
          // EC set of the first-hit block: recover its id from the tuple?  Not possible in general
          // (empty sets are not in the tuple), so the block's set is looked up via blk_ec.
          const unsigned long long bword = dd.dslots[ix.blk_ec[blk]];   // blk_ec holds set handles
          const uint32_t* B = pool + (uint32_t)bword;
          const uint32_t blen = (uint32_t)((bword >> 32) & 0xFFFFFFu);
          const uint8_t* sb = ix.strand + ix.blk_strand_off[blk];
          // u &= ec ; vtmp = strand-compatible subset
          uint32_t n_u = 0, n_v = 0;
          // two passes over scratch, compacting in place: first u &= ec (keeping a flag per kept
          // element in the top of the scratch area is avoided by recomputing the predicate)
          for (uint32_t base = 0; base < nres; base += G) {
            const uint32_t i = base + lane;
            bool in_u = i < nres;
            const uint32_t a = in_u ? scratch[i] : 0;
            uint32_t rank = 0;
            if (in_u) in_u = bsearch_contains(B, blen, a, &rank);
            bool in_v = false;
            if (in_u) {
              const uint8_t sense = sb[rank];
              in_v = ((um_strand == (sense != 0)) == want) || sense == 2;
            }
            const unsigned bu = (__ballot_sync(gmask, in_u) >> gshift);
            const unsigned bv = (__ballot_sync(gmask, in_v) >> gshift);
            __syncwarp(gmask);
            // u goes to the front of scratch (in place: n_u <= base), v to the second half
            if (in_u) scratch[n_u + __popc(bu & ((1u << lane) - 1))] = a;
            if (in_v) scratch[ra.scratch_stride / 2 + n_v + __popc(bv & ((1u << lane) - 1))] = a;
            n_u += __popc(bu);
            n_v += __popc(bv);
            __syncwarp(gmask);
          }
          if (n_v < n_u) {
            for (uint32_t i = lane; i < n_v; i += G) scratch[i] = scratch[ra.scratch_stride / 2 + i];
            nres = n_v;
          } else {
            nres = n_u;
          }
          __syncwarp(gmask);
        }
      }
      // 4. set -> handle through the content-addressed dictionary
      handle = nres == 0 ? KB_H_UNMAPPED : dict_insert_warp<G>(dd, scratch, nres, lane, gmask);
      // 5. publish tuple -> handle (not for position-filtered fragments: the result depends on the read)
      if (lane == 0 && !no_memo) {
        if (use_m2) {
          const unsigned long long key = ((unsigned long long)w[0] << 32) | w[1];
          uint64_t s = kb_mix64(key) & dd.m2_mask;
          uint64_t visited = 0;
          for (;;) {
            const unsigned long long old = atomicCAS(&dd.m2[s].key, ~0ULL, key);
            if (old == ~0ULL || old == key) { atomicExch(&dd.m2[s].val, handle); break; }
            s = (s + 1) & dd.m2_mask;
            if (++visited > dd.m2_mask) { atomicOr(dd.error, KB_DEVERR_MEMO_FULL); break; }
          }
        } else {
          const uint64_t th = tuple_hash(w, n, 1);
          const uint32_t tag = (uint32_t)(th >> 32);
          const unsigned long long toff = atomicAdd(dd.tpool_top, (unsigned long long)(n + 1));
          if (toff + n + 1 > dd.tpool_cap) {
            atomicOr(dd.error, KB_DEVERR_TPOOL_FULL);
          } else {
            dd.tpool[toff] = (uint32_t)n;
            for (int i = 0; i < n; ++i) dd.tpool[toff + 1 + i] = w[i];
            __threadfence();
            const unsigned long long word = ((unsigned long long)tag << 32) | toff;
            uint64_t s = th & dd.mn_mask;
            uint64_t visited = 0;
            for (;;) {
              unsigned long long old = atomicCAS(&dd.mn_key[s], ~0ULL, word);
              bool mine = (old == ~0ULL);
              if (!mine && (uint32_t)(old >> 32) == tag) {
                const uint32_t* t = dd.tpool + (uint32_t)old;
                bool eq = __ldcg(t) == (uint32_t)n;
                for (int i = 0; eq && i < n; ++i) eq = __ldcg(t + 1 + i) == w[i];
                mine = eq;
              }
              if (mine) { atomicExch(&dd.mn_val[s], handle); break; }
              s = (s + 1) & dd.mn_mask;
              if (++visited > dd.mn_mask) { atomicOr(dd.error, KB_DEVERR_MEMO_FULL); break; }
            }
          }
        }
      }
    }
    // 6. account for this fragment
    if (lane == 0) {
      ba.handle_out[f] = handle;
      if (handle >= 0) {
        atomicAdd(&dd.count[handle], 1u);
        atomicMin(&dd.first[handle], (unsigned long long)(ba.frag_base + f));
      }
    }
    __syncwarp(gmask);
  }
  }
}

// A fragment contributes to the fragment-length distribution only if its EC has a single
// transcript (ProcessReads.cpp:1174).
__global__ void fld_finalize_kernel(DevDict dd, BatchArgs ba) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= ba.n_frag) return;
  const int32_t h = ba.handle_out[f];
  if (h < 0) { ba.tl_out[f] = 0; return; }
  const unsigned long long word = dd.dslots[h];
  if (((word >> 32) & 0xFFFFFFu) != 1) ba.tl_out[f] = 0;
}

__global__ void collect_used_kernel(DevDict dd, uint32_t* used, uint32_t* n_used) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h <= dd.dmask; h += stride) {
    if (dd.count[h] > 0) used[atomicAdd(n_used, 1u)] = (uint32_t)h;
  }
}

void launch_pseudoalign(const DevIndex& ix, const DevDict& dd, const BatchArgs& ba, const ResolveArgs& ra,
                        int tpb, cudaStream_t st, cudaEvent_t* ev) {
  if (ba.n_frag == 0) return;
  cudaMemsetAsync(ba.q_count, 0, sizeof(uint32_t), st);
  cudaMemsetAsync(ba.qbig_count, 0, sizeof(uint32_t), st);
  // persistent grid: as many blocks as fit on the device at once
  const size_t smem = (size_t)tpb * 4 * ((size_t)KB_MAX_E + 3 + 4 * ba.nb);   // per lane: handle tuple, first-mate words, 2 x 2nb base words
  const int sms = device_sm_count();
  cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, match_kernel, tpb, smem);
  if (per_sm < 1) per_sm = 1;
  unsigned blocks = (unsigned)(sms * per_sm);
  const unsigned need = (ba.n_frag + tpb - 1) / tpb;   // never more lanes than fragments
  if (blocks > need) blocks = need;
  if (ev) cudaEventRecord(ev[0], st);
  {
    const uint32_t n_reads = ba.paired ? 2 * ba.n_frag : ba.n_frag;
    const uint64_t total = (uint64_t)n_reads * ba.nb;     // < 2^32: engine.cu bounds a batch by 2^31 bases
    pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ba, n_reads, const_cast<uint32_t*>(ba.packed));
    if (ix.dfk && ba.skip_w) dlist_scan_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ix, ba, n_reads);
  }
  if (ev) cudaEventRecord(ev[1], st);
  match_kernel<<<blocks, tpb, smem, st>>>(ix, dd, ba);
  if (ev) cudaEventRecord(ev[2], st);
  switch (ra.group) {      // lanes per fragment (engine.cu: KB_RESOLVE_G, default 32: profiles/resolve_group_sweep_r02.jsonl)
    case 4: resolve_kernel<4><<<(ra.n_warps * 4 + 127) / 128, 128, 0, st>>>(ix, dd, ba, ra); break;
    case 8: resolve_kernel<8><<<(ra.n_warps * 8 + 127) / 128, 128, 0, st>>>(ix, dd, ba, ra); break;
    case 16: resolve_kernel<16><<<(ra.n_warps * 16 + 127) / 128, 128, 0, st>>>(ix, dd, ba, ra); break;
    default: resolve_kernel<32><<<(ra.n_warps * 32 + 127) / 128, 128, 0, st>>>(ix, dd, ba, ra); break;
  }
  if (ev) cudaEventRecord(ev[3], st);
}

// Multi-GPU merge: equivalence classes exported by another rank (CSR of transcript ids, counts, first
// fragment index) are folded into this rank's dictionary -- the content-keyed reduction that replaces
// a dense all-reduce, since EC ids are discovered independently on every rank.  One warp per set.
__global__ void __launch_bounds__(128) import_sets_kernel(DevDict dd, uint32_t n_sets, const uint32_t* off, const uint32_t* tids,
                                                         const uint32_t* counts, const unsigned long long* first,
                                                         unsigned long long first_offset) {
  const unsigned lane = threadIdx.x & 31;
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t s = w; s < n_sets; s += nw) {
    const uint32_t o0 = off[s], n = off[s + 1] - o0;
    if (n == 0) continue;
    const int32_t h = dict_insert_warp(dd, tids + o0, n, lane);
    if (lane == 0 && h >= 0) {
      atomicAdd(&dd.count[h], counts[s]);
      atomicMin(&dd.first[h], first[s] + first_offset);
    }
    __syncwarp();
  }
}

struct ImportSegs {
  int n;
  uint32_t prefix[KB_IMPORT_SEGS + 1];
  ImportSeg seg[KB_IMPORT_SEGS];
};
__global__ void __launch_bounds__(128) import_segments_kernel(DevDict dd, ImportSegs a) {
  const unsigned lane = threadIdx.x & 31;
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
  const uint32_t total = a.prefix[a.n];
  for (uint32_t g = w; g < total; g += nw) {
    int k = 0;
    while (g >= a.prefix[k + 1]) ++k;
    const ImportSeg& sg = a.seg[k];
    const uint32_t s = g - a.prefix[k];
    const uint32_t o0 = sg.off[s], n = sg.off[s + 1] - o0;
    if (n == 0) continue;
    const int32_t h = dict_insert_warp(dd, sg.tids + o0, n, lane);
    if (lane == 0 && h >= 0) {
      atomicAdd(&dd.count[h], sg.counts[s]);
      atomicMin(&dd.first[h], sg.first[s]);
    }
    __syncwarp();
  }
}

int device_sm_count() {
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (sms[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = v > 0 ? v : 1;
  }
  return sms[dev];
}

void launch_import_segments(const DevDict& dd, const ImportSeg* segs, int n_segs, cudaStream_t st) {
  if (n_segs <= 0) return;
  ImportSegs a;
  a.n = n_segs;
  a.prefix[0] = 0;
  for (int i = 0; i < n_segs; ++i) { a.seg[i] = segs[i]; a.prefix[i + 1] = a.prefix[i] + segs[i].n_sets; }
  if (a.prefix[n_segs] == 0) return;
  const unsigned warps_needed = a.prefix[n_segs];
  unsigned blocks = (unsigned)device_sm_count() * 16;
  if (blocks > (warps_needed + 3) / 4) blocks = (warps_needed + 3) / 4;
  import_segments_kernel<<<blocks, 128, 0, st>>>(dd, a);
}

void launch_import_sets(const DevDict& dd, uint32_t n_sets, const uint32_t* off, const uint32_t* tids, const uint32_t* counts,
                        const unsigned long long* first, unsigned long long first_offset, cudaStream_t st) {
  if (n_sets == 0) return;
  import_sets_kernel<<<device_sm_count() * 8, 128, 0, st>>>(dd, n_sets, off, tids, counts, first, first_offset);
}

void launch_fld_finalize(const DevDict& dd, const BatchArgs& ba, cudaStream_t st) {
  if (ba.n_frag == 0 || !ba.tl_out) return;
  fld_finalize_kernel<<<(ba.n_frag + 255) / 256, 256, 0, st>>>(dd, ba);
}

void launch_collect_used(const DevDict& dd, uint32_t* used, uint32_t* n_used, cudaStream_t st) {
  cudaMemsetAsync(n_used, 0, sizeof(uint32_t), st);
  collect_used_kernel<<<device_sm_count() * 8, 256, 0, st>>>(dd, used, n_used);
}

}  // namespace kb
