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

namespace kb {

namespace {

struct Hit {
  uint32_t unitig, blk, ec, dist, lb, ub;
  bool strand;
};

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

// One lookup in the k-mer table: dbg.find(km) + get_mc_contig + ec[dist] of the reference.
// `strand` = the k-mer as it appears in the read equals the unitig-forward k-mer
// (CompactedDBG.tcc:1049-1107).
__device__ __forceinline__ bool probe(const DevIndex& ix, uint64_t fwd, Hit& h, uint32_t& n_visits) {
  const uint64_t rc = kb_revcomp(fwd, ix.k);
  const bool is_canon = fwd < rc;
  const uint64_t canon = is_canon ? fwd : rc;
  uint64_t s = kb_mix64(canon) & ix.mask;
  for (;;) {
    const uint4* sp = reinterpret_cast<const uint4*>(ix.slots + s);
    const uint4 a = __ldg(sp);
    ++n_visits;
    const uint64_t key = (uint64_t)a.x | ((uint64_t)a.y << 32);
    if (key == canon) {
      const uint4 b = __ldg(sp + 1);
      h.unitig = a.z;
      h.blk = a.w;
      h.ec = b.x;
      h.dist = b.y & 0x7FFFFFFFu;
      const bool fic = (b.y >> 31) != 0;
      h.strand = (is_canon == fic);
      h.lb = b.z;
      h.ub = b.w;
      return true;
    }
    if (key == KB_EMPTY_KEY) return false;
    s = (s + 1) & ix.mask;
  }
}

// Per-thread view of the read currently being matched: 2-bit bases and an invalid-base mask in
// shared memory, word w of thread t at [w * blockDim.x + t] (bank-conflict free).
struct ReadView {
  uint64_t* bw;   // base words: base i in bits 62-2*(i&31) of word i>>5
  uint64_t* iv;   // invalid mask: bit (i&63) of word i>>6
  int stride;
  int len;
  int k;

  __device__ __forceinline__ uint64_t kmer(int p) const {
    const int w = p >> 5, s = (p & 31) * 2;
    const uint64_t hi = bw[w * stride];
    uint64_t x = hi << s;
    if (s) x |= bw[(w + 1) * stride] >> (64 - s);
    return x >> (64 - 2 * k);
  }
  // first start position >= p whose k-window holds only A/C/G/T, or -1
  // (KmerIterator::operator++ / operator+=, ext/bifrost/src/KmerIterator.cpp:6-63)
  __device__ __forceinline__ int next_valid(int p) const {
    const uint64_t wmask = (1ULL << k) - 1;
    while (p <= len - k) {
      const int w = p >> 6, s = p & 63;
      uint64_t x = iv[w * stride] >> s;
      if (s) x |= iv[(w + 1) * stride] << (64 - s);
      x &= wmask;
      if (x == 0) return p;
      p += 64 - __clzll((long long)x);
    }
    return -1;
  }
};

__device__ __forceinline__ void load_read(const BatchArgs& ba, uint32_t read_idx, ReadView& rv) {
  uint64_t off;
  int len;
  if (ba.off) {
    off = ba.off[read_idx];
    len = (int)(ba.off[read_idx + 1] - ba.off[read_idx]);
  } else {
    off = (uint64_t)read_idx * ba.fixed_len;
    len = (int)ba.fixed_len;
  }
  const int nbw = (int)ba.bwords, niw = (int)ba.iwords;   // host guarantees niw == nbw/2 + 1
  const int maxlen = (nbw - 1) * 32;
  if (len > maxlen) len = maxlen;   // cannot happen: the host sizes bwords from the longest read
  rv.len = len;
  const uint8_t* s = ba.bases + off;
  for (int w2 = 0; w2 < niw; ++w2) {
    uint64_t inv64 = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int w = 2 * w2 + half;
      uint32_t inv32 = ~0u;
      if (w < nbw) {
        uint64_t bwv = 0;
        const int base = w * 32;
        if (base < len) {
          const int n = min(32, len - base);
          inv32 = 0;
          for (int j = 0; j < n; ++j) {
            const uint32_t c = __ldg(s + base + j);
            const uint32_t x = (c & 4) >> 1;
            const uint32_t code = x + ((x ^ (c & 2)) >> 1);          // Kmer::set_kmer
            const uint32_t u = c & 0xDF;                               // KmerIterator: mask lowercase bit
            const bool ok = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');   // isDNA, Common.hpp:45-50
            bwv |= (uint64_t)code << (62 - 2 * j);
            inv32 |= (ok ? 0u : 1u) << j;
          }
          if (n < 32) inv32 |= ~0u << n;
        }
        rv.bw[w * rv.stride] = bwv;
      }
      inv64 |= (uint64_t)inv32 << (32 * half);
    }
    rv.iv[w2 * rv.stride] = inv64;
  }
}

struct MateInfo {
  bool v_nonempty;     // match() returned at least one hit
  bool s_nonempty;     // at least one hit with a non-empty EC set
  // first hit (smallest read position): findFirstMappingKmer / mapPair
  uint32_t f_unitig, f_blk, f_ec, f_dist, f_ub;
  int f_pos;
  bool f_strand;
};

struct FragState {
  uint32_t* elist;   // shared memory, strided
  int stride;
  int n_e;
  bool overflow;
  uint32_t empty_ec;
};

__device__ __forceinline__ void push_hit(FragState& fs, MateInfo& mi, const Hit& h, int pos) {
  if (!mi.v_nonempty) {
    mi.v_nonempty = true;
    mi.f_unitig = h.unitig; mi.f_blk = h.blk; mi.f_ec = h.ec; mi.f_dist = h.dist; mi.f_ub = h.ub;
    mi.f_pos = pos; mi.f_strand = h.strand;
  }
  if (h.ec == fs.empty_ec) return;   // "Don't intersect empty EC", MinCollector.cpp:468-469
  mi.s_nonempty = true;
  for (int i = 0; i < fs.n_e; ++i)
    if (fs.elist[i * fs.stride] == h.ec) return;
  if (fs.n_e == KB_MAX_E) { fs.overflow = true; return; }
  fs.elist[fs.n_e * fs.stride] = h.ec;
  ++fs.n_e;
}

__device__ __forceinline__ bool same_ue(const Hit& a, const Hit& b) {
  // um.isSameReferenceUnitig(um2) && ec[um.dist] == ec[um2.dist]   (KmerIndex.cpp:1810-1811)
  return a.unitig == b.unitig && a.ec == b.ec;
}

// KmerIndex::match for one read, default flags (no shade/union/no_jump/cfc, empty D-list).
// `partial` only short-circuits reads whose running intersection empties; the final result is
// the same either way, so it is not modelled.
__device__ void match_read(const DevIndex& ix, const ReadView& rv, FragState& fs, MateInfo& mi,
                           uint32_t& n_probes, uint32_t& n_visits) {
  const int k = rv.k, l = rv.len;
  int p = rv.next_valid(0);
  while (p >= 0) {
    Hit h;
    ++n_probes;
    if (probe(ix, rv.kmer(p), h, n_visits)) {
      push_hit(fs, mi, h, p);
      const int off = (int)(h.dist - h.lb), blen = (int)(h.ub - h.lb);
      const int dist = h.strand ? (blen - 1 - off) : off;                    // 1780-1788
      if (dist >= 2) {
        const int nextPos = (p + dist >= l - k) ? (l - k) : (p + dist);      // 1793-1798
        const int adv = nextPos - p;
        const int p2 = adv == 0 ? p : rv.next_valid(p + adv);                // kit2 += nextPos-pos
        if (p2 < 0) break;                                                   // 1882-1886
        Hit h2;
        ++n_probes;
        const bool f2 = probe(ix, rv.kmer(p2), h2, n_visits);
        bool found2 = false;
        int found2pos = p + dist;
        if (!f2) { found2 = true; found2pos = p; }
        else if (same_ue(h, h2)) { found2 = true; found2pos = p + dist; }
        if (found2) {
          if (found2pos >= l - k) { push_hit(fs, mi, h, l - k); break; }
          push_hit(fs, mi, h, found2pos);
          p = p2;
        } else {
          bool foundMiddle = false;
          if (dist > 4) {
            const int middlePos = (p + nextPos) / 2;
            const int adv3 = middlePos - p;
            const int p3 = adv3 == 0 ? p : rv.next_valid(p + adv3);
            if (p3 >= 0) {
              Hit h3;
              ++n_probes;
              if (probe(ix, rv.kmer(p3), h3, n_visits)) {
                int found3pos = p + dist;
                if (same_ue(h, h3)) { foundMiddle = true; found3pos = middlePos; }
                else if (same_ue(h2, h3)) { foundMiddle = true; found3pos = p + dist; }
                if (foundMiddle) push_hit(fs, mi, h3, found3pos);
              }
              if (foundMiddle) {
                if (nextPos >= l - k) break;
                p = p2;
              }
            }
          }
          if (!foundMiddle) {
            p = rv.next_valid(p + 1);          // ++kit; backOff: exactly one probe (outer nextPos == 0)
            if (p < 0) break;
            Hit h4;
            ++n_probes;
            if (probe(ix, rv.kmer(p), h4, n_visits)) push_hit(fs, mi, h4, p);
          }
        }
      }
    }
    p = rv.next_valid(p + 1);
  }
}

__device__ __forceinline__ uint64_t tuple_hash(const uint32_t* w, int n, int stride) {
  uint64_t h = 0x243F6A8885A308D3ULL ^ (uint64_t)n;
  for (int i = 0; i < n; ++i) h = kb_mix64(h ^ ((uint64_t)w[i * stride] + 0x9E3779B97F4A7C15ULL * (i + 1)));
  return h;
}

// Memo lookups.  Return KB_H_NOTREADY on a miss.
__device__ __forceinline__ int32_t memo2_lookup(const DevDict& dd, uint32_t e0, uint32_t e1) {
  const unsigned long long key = ((unsigned long long)e0 << 32) | e1;
  uint64_t s = kb_mix64(key) & dd.m2_mask;
  for (;;) {
    const unsigned long long kk = __ldcg(&dd.m2_key[s]);
    if (kk == key) return ld_relaxed_s32(&dd.m2_val[s]);
    if (kk == ~0ULL) return KB_H_NOTREADY;
    s = (s + 1) & dd.m2_mask;
  }
}
__device__ __forceinline__ int32_t memon_lookup(const DevDict& dd, const uint32_t* w, int n, int stride) {
  const uint64_t th = tuple_hash(w, n, stride);
  const uint32_t tag = (uint32_t)(th >> 32);
  uint64_t s = th & dd.mn_mask;
  for (;;) {
    const unsigned long long word = ld_acquire_u64(&dd.mn_key[s]);
    if (word == ~0ULL) return KB_H_NOTREADY;
    if ((uint32_t)(word >> 32) == tag) {
      const uint32_t* t = dd.tpool + (uint32_t)word;
      bool eq = __ldcg(t) == (uint32_t)n;
      for (int i = 0; eq && i < n; ++i) eq = __ldcg(t + 1 + i) == w[i * stride];
      if (eq) return ld_relaxed_s32(&dd.mn_val[s]);
    }
    s = (s + 1) & dd.mn_mask;
  }
}

// Warp-aggregated per-handle accounting (valid for any subset of participating lanes that calls
// it convergently with the full mask).
__device__ __forceinline__ void account(const DevDict& dd, int32_t handle, uint64_t frag, unsigned lane) {
  const unsigned grp = __match_any_sync(0xFFFFFFFFu, handle);
  if (handle >= 0) {
    const unsigned leader = __ffs(grp) - 1;
    if (lane == leader) {
      atomicAdd(&dd.count[handle], (uint32_t)__popc(grp));
      atomicMin(&dd.first[handle], (unsigned long long)frag);   // lowest lane = lowest fragment index
    }
  }
}

}  // namespace

__global__ void __launch_bounds__(256) match_kernel(DevIndex ix, DevDict dd, BatchArgs ba) {
  extern __shared__ uint64_t smem[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const unsigned lane = tid & 31;
  const uint32_t f = blockIdx.x * nt + tid;
  const bool active = f < ba.n_frag;

  ReadView rv;
  rv.bw = smem + tid;
  rv.iv = smem + (size_t)ba.bwords * nt + tid;
  rv.stride = nt;
  rv.k = ix.k;
  rv.len = 0;
  FragState fs;
  fs.elist = reinterpret_cast<uint32_t*>(smem + (size_t)(ba.bwords + ba.iwords) * nt) + tid;
  fs.stride = nt;
  fs.n_e = 0;
  fs.overflow = false;
  fs.empty_ec = ba.empty_ec;

  MateInfo m[2];
  m[0].v_nonempty = m[0].s_nonempty = false;
  m[1].v_nonempty = m[1].s_nonempty = false;
  uint32_t n_probes = 0, n_visits = 0, n_memo = 0;

  int32_t handle = KB_H_UNMAPPED;
  bool queued = false;
  if (active) {
    const int nm = ba.paired ? 2 : 1;
    for (int mate = 0; mate < nm; ++mate) {
      load_read(ba, ba.paired ? 2 * f + mate : f, rv);
      match_read(ix, rv, fs, m[mate], n_probes, n_visits);
    }
    // ---- MinCollector::intersectKmers, net effect (MinCollector.cpp:160-218) ----
    bool mapped = m[0].v_nonempty || m[1].v_nonempty;
    if ((m[0].v_nonempty && !m[0].s_nonempty) || (m[1].v_nonempty && !m[1].s_nonempty)) mapped = false;
    if (mapped && fs.n_e == 0) mapped = false;
    if (mapped) {
      // sort the distinct EC-set ids (insertion sort, <= 16 entries)
      for (int i = 1; i < fs.n_e; ++i) {
        const uint32_t v = fs.elist[i * nt];
        int j = i - 1;
        while (j >= 0 && fs.elist[j * nt] > v) { fs.elist[(j + 1) * nt] = fs.elist[j * nt]; --j; }
        fs.elist[(j + 1) * nt] = v;
      }
      uint32_t sw0 = 0, sw1 = 0;
      if (fs.overflow) {
        atomicOr(dd.error, KB_DEVERR_E_OVERFLOW);
        handle = KB_H_UNMAPPED;
      } else if (ba.strand_mode == 0 && fs.n_e == 1) {
        handle = ix.ec_handle[fs.elist[0]];
      } else {
        int32_t r;
        if (ba.strand_mode == 0 && fs.n_e == 2) {
          r = memo2_lookup(dd, fs.elist[0], fs.elist[nt]);
        } else {
          int n = fs.n_e;
          if (ba.strand_mode != 0) {
            // the strand filter depends on the first hit of each mate: (block, orientation)
            sw0 = m[0].v_nonempty ? (m[0].f_blk * 2u + (m[0].f_strand ? 1u : 0u)) : 0xFFFFFFFFu;
            sw1 = m[1].v_nonempty ? (m[1].f_blk * 2u + (m[1].f_strand ? 1u : 0u)) : 0xFFFFFFFFu;
            fs.elist[n * nt] = sw0;          // elist has KB_MAX_E + 2 words per thread
            fs.elist[(n + 1) * nt] = sw1;
            n += 2;
          }
          r = memon_lookup(dd, fs.elist, n, nt);
        }
        if (r == KB_H_NOTREADY) {
          queued = true;
          handle = KB_H_PENDING;
        } else {
          handle = r;
          ++n_memo;
        }
      }
    }
    if (queued) {
      const uint32_t q = atomicAdd(ba.q_count, 1u);
      uint32_t* e = ba.q_entries + (size_t)q * KB_Q_STRIDE;
      const int n = fs.n_e + (ba.strand_mode != 0 ? 2 : 0);
      e[0] = f;
      e[1] = (uint32_t)n;
      for (int i = 0; i < n; ++i) e[2 + i] = fs.elist[i * nt];
    }
    ba.handle_out[f] = handle;
    if (ba.tl_out) {
      // KmerIndex::mapPair (KmerIndex.cpp:1622-1693): the first k-mer found by a linear scan is the
      // first hit of match(); same unitig, same EC set, opposite strands, same block end.
      uint16_t tl = 0;
      if (ba.paired && m[0].v_nonempty && m[1].v_nonempty) {
        const int k = ix.k;
        const int p1 = m[0].f_strand ? (int)m[0].f_dist - m[0].f_pos : (int)m[0].f_dist + k + m[0].f_pos;
        const int p2 = m[1].f_strand ? (int)m[1].f_dist - m[1].f_pos : (int)m[1].f_dist + k + m[1].f_pos;
        if (m[0].f_unitig == m[1].f_unitig && m[0].f_ec == m[1].f_ec && (m[0].f_strand != m[1].f_strand) &&
            m[0].f_ub == m[1].f_ub) {
          const int d = p1 > p2 ? p1 - p2 : p2 - p1;
          if (d > 0 && d < 1000) tl = (uint16_t)d;
        }
      }
      ba.tl_out[f] = tl;
    }
  }
  __syncwarp();
  account(dd, (active && handle >= 0) ? handle : (int32_t)(-100 - (int)lane), ba.frag_base + f, lane);
  // statistics: probes and slot visits
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

}  // namespace

// One warp per queued fragment.
__global__ void __launch_bounds__(128) resolve_kernel(DevIndex ix, DevDict dd, BatchArgs ba, ResolveArgs ra) {
  const unsigned lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= ra.n_warps) return;
  const uint32_t nq = *ba.q_count;
  uint32_t* scratch = ra.scratch + (size_t)warp * ra.scratch_stride;
  const uint32_t* pool = dd.pool;

  for (uint32_t q = warp; q < nq; q += ra.n_warps) {
    const uint32_t* e = ba.q_entries + (size_t)q * KB_Q_STRIDE;
    const uint32_t f = e[0];
    const int n = (int)e[1];
    const uint32_t* w = e + 2;
    const bool stranded = ba.strand_mode != 0;
    const int n_e = stranded ? n - 2 : n;
    const bool use_m2 = (!stranded && n_e == 2);

    // 1. has somebody else resolved the same tuple in the meantime?
    int32_t handle = KB_H_NOTREADY;
    if (lane == 0) handle = use_m2 ? memo2_lookup(dd, w[0], w[1]) : memon_lookup(dd, w, n, 1);
    handle = __shfl_sync(0xFFFFFFFFu, handle, 0);

    if (handle == KB_H_NOTREADY) {
      // 2. intersection of the n_e sets: lanes own elements of the smallest one
      int sm = 0;
      uint32_t sm_len = ix.ec_off[w[0] + 1] - ix.ec_off[w[0]];
      for (int j = 1; j < n_e; ++j) {
        const uint32_t len = ix.ec_off[w[j] + 1] - ix.ec_off[w[j]];
        if (len < sm_len) { sm_len = len; sm = j; }
      }
      const uint32_t* A = pool + ix.ec_off[w[sm]];
      uint32_t nres = 0;
      for (uint32_t base = 0; base < sm_len; base += 32) {
        const uint32_t i = base + lane;
        bool alive = i < sm_len;
        const uint32_t a = alive ? __ldcg(A + i) : 0;
        for (int j = 0; j < n_e; ++j) {
          if (j == sm) continue;
          const uint32_t* B = pool + ix.ec_off[w[j]];
          const uint32_t blen = ix.ec_off[w[j] + 1] - ix.ec_off[w[j]];
          if (alive) alive = bsearch_contains(B, blen, a, nullptr);
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, alive);
        if (alive) scratch[nres + __popc(bal & ((1u << lane) - 1))] = a;
        nres += __popc(bal);
      }
      __syncwarp();
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
          const uint32_t be = ix.blk_ec[blk];
          const uint32_t* B = pool + ix.ec_off[be];
          const uint32_t blen = ix.ec_off[be + 1] - ix.ec_off[be];
          const uint8_t* sb = ix.strand + ix.blk_strand_off[blk];
          // u &= ec ; vtmp = strand-compatible subset
          uint32_t n_u = 0, n_v = 0;
          // two passes over scratch, compacting in place: first u &= ec (keeping a flag per kept
          // element in the top of the scratch area is avoided by recomputing the predicate)
          for (uint32_t base = 0; base < nres; base += 32) {
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
            const unsigned bu = __ballot_sync(0xFFFFFFFFu, in_u);
            const unsigned bv = __ballot_sync(0xFFFFFFFFu, in_v);
            __syncwarp();
            // u goes to the front of scratch (in place: n_u <= base), v to the second half
            if (in_u) scratch[n_u + __popc(bu & ((1u << lane) - 1))] = a;
            if (in_v) scratch[ra.scratch_stride / 2 + n_v + __popc(bv & ((1u << lane) - 1))] = a;
            n_u += __popc(bu);
            n_v += __popc(bv);
            __syncwarp();
          }
          if (n_v < n_u) {
            for (uint32_t i = lane; i < n_v; i += 32) scratch[i] = scratch[ra.scratch_stride / 2 + i];
            nres = n_v;
          } else {
            nres = n_u;
          }
          __syncwarp();
        }
      }
      // 4. set -> handle through the content-addressed dictionary
      if (nres == 0) {
        handle = KB_H_UNMAPPED;
      } else {
        uint64_t sum = 0;
        for (uint32_t i = lane; i < nres; i += 32) sum += kb_mix64((uint64_t)scratch[i] + 0x9E3779B97F4A7C15ULL);
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        const uint64_t hsh = kb_mix64(sum ^ nres);
        const unsigned long long tag = hsh >> 56;
        uint64_t s = hsh & dd.dmask;
        unsigned long long my_word = ~0ULL;   // allocated lazily
        uint64_t visited = 0;
        for (;;) {
          unsigned long long word = 0;
          if (lane == 0) word = ld_acquire_u64(&dd.dslots[s]);
          word = __shfl_sync(0xFFFFFFFFu, word, 0);
          if (word == ~0ULL) {
            if (my_word == ~0ULL) {
              unsigned long long off = 0;
              if (lane == 0) off = atomicAdd(dd.pool_top, (unsigned long long)nres);
              off = __shfl_sync(0xFFFFFFFFu, off, 0);
              if (off + nres > dd.pool_cap || off + nres > 0xFFFFFFFFULL) {
                if (lane == 0) atomicOr(dd.error, KB_DEVERR_POOL_FULL);
                handle = KB_H_UNMAPPED;
                break;
              }
              for (uint32_t i = lane; i < nres; i += 32) dd.pool[off + i] = scratch[i];
              __threadfence();
              __syncwarp();
              my_word = off | ((unsigned long long)nres << 32) | (tag << 56);
            }
            unsigned long long old = 0;
            if (lane == 0) old = atomicCAS(&dd.dslots[s], ~0ULL, my_word);
            old = __shfl_sync(0xFFFFFFFFu, old, 0);
            if (old == ~0ULL) { handle = (int32_t)s; break; }
            word = old;   // somebody else took the slot: compare against theirs
          }
          if ((word >> 56) == tag && ((word >> 32) & 0xFFFFFFu) == nres) {
            const uint32_t* S = pool + (uint32_t)word;
            bool eq = true;
            for (uint32_t i = lane; i < nres; i += 32) eq = eq && (__ldcg(S + i) == scratch[i]);
            if (__all_sync(0xFFFFFFFFu, eq)) { handle = (int32_t)s; break; }
          }
          s = (s + 1) & dd.dmask;
          if (++visited > dd.dmask) {
            if (lane == 0) atomicOr(dd.error, KB_DEVERR_DICT_FULL);
            handle = KB_H_UNMAPPED;
            break;
          }
        }
      }
      // 5. publish tuple -> handle
      if (lane == 0) {
        if (use_m2) {
          const unsigned long long key = ((unsigned long long)w[0] << 32) | w[1];
          uint64_t s = kb_mix64(key) & dd.m2_mask;
          uint64_t visited = 0;
          for (;;) {
            const unsigned long long old = atomicCAS(&dd.m2_key[s], ~0ULL, key);
            if (old == ~0ULL || old == key) { atomicExch(&dd.m2_val[s], handle); break; }
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
      atomicAdd(&dd.stats[1], 1ULL);
    }
    __syncwarp();
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
  const size_t smem = (size_t)tpb * ((size_t)(ba.bwords + ba.iwords) * 8 + (KB_MAX_E + 2) * 4);
  const unsigned blocks = (ba.n_frag + tpb - 1) / tpb;
  if (ev) cudaEventRecord(ev[0], st);
  match_kernel<<<blocks, tpb, smem, st>>>(ix, dd, ba);
  if (ev) cudaEventRecord(ev[1], st);
  resolve_kernel<<<(ra.n_warps * 32 + 127) / 128, 128, 0, st>>>(ix, dd, ba, ra);
  if (ev) cudaEventRecord(ev[2], st);
}

void launch_fld_finalize(const DevDict& dd, const BatchArgs& ba, cudaStream_t st) {
  if (ba.n_frag == 0 || !ba.tl_out) return;
  fld_finalize_kernel<<<(ba.n_frag + 255) / 256, 256, 0, st>>>(dd, ba);
}

void launch_collect_used(const DevDict& dd, uint32_t* used, uint32_t* n_used, cudaStream_t st) {
  cudaMemsetAsync(n_used, 0, sizeof(uint32_t), st);
  collect_used_kernel<<<148 * 8, 256, 0, st>>>(dd, used, n_used);
}

}  // namespace kb
