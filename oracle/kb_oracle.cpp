// ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing in the product (kallisto_b200/, include/) may
// include, link, load or execute this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg use it, and only as the checker.
//
// Plain single-threaded CPU restatement of the kallisto `quant`/`bus` hot path, written to follow
// the reference's control flow literally (vectors of hits, sort, Roaring-style set operations on
// sorted vectors, EM loops in ecmapinv iteration order), each function citing the reference
// lines it restates.  Parity status: PINNED -- tests/test_oracle_vs_reference.py checks this file
// against the unmodified reference binary built by oracle/Makefile (oracle/_ref/kallisto) on the
// reference's bundled test data and func_tests inputs (abundance.tsv md5s of
// func_tests/runtests.sh:265-304, EC multisets and fragment-length histograms from
// `kallisto bus -x bulk`), and against the golden vectors committed under tests/golden/.
//
// Build: g++ -O2 -std=c++17 -shared -fPIC oracle/kb_oracle.cpp -o oracle/liboracle.so
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

typedef std::vector<uint32_t> TidSet;   // sorted, unique: stands in for a Roaring bitmap

// ---------------------------------------------------------------------------------------------
// Index (format v13): KmerIndex::load, src/KmerIndex.cpp:1330-1559
// ---------------------------------------------------------------------------------------------
struct Block {
  uint32_t lb, ub;
  TidSet tids;
  std::vector<uint8_t> sense;               // SparseVector::operator[], SparseVector.tcc:363-389
  std::vector<std::vector<uint32_t>> pos;   // per transcript: pos | antisense<<31
};
struct Unitig {
  std::string seq;            // forward sequence (short / abundant: the stored canonical k-mer)
  int kind;                   // 0 long, 1 short, 2 abundant
  std::vector<Block> blocks;  // BlockArray
};
struct KmerLoc {
  uint32_t unitig;
  uint32_t dist;
  bool fwd_is_rep;
};
struct OIndex {
  int k = 0;
  std::vector<Unitig> unitigs;
  std::unordered_map<uint64_t, KmerLoc> kmers;   // canonical k-mer -> location
  std::vector<uint32_t> target_len;
  std::vector<std::string> target_name;
  TidSet onlist;
  // D-list (distinguishing flanking k-mers, src/KmerIndex.cpp:1385-1403): canonical k-mers; the first one is the dummy,
  // the only one that is part of the graph (its unitig carries the single off-list target id)
  std::unordered_set<uint64_t> d_list;
  uint64_t dummy_dfk = 0;
};

struct Reader {
  const uint8_t* p;
  const uint8_t* e;
  template <class T> T get() {
    if ((size_t)(e - p) < sizeof(T)) throw std::runtime_error("oracle: truncated index");
    T v;
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  const uint8_t* take(size_t n) {
    if ((size_t)(e - p) < n) throw std::runtime_error("oracle: truncated index");
    const uint8_t* r = p;
    p += n;
    return r;
  }
};

int code_of(char c) {   // Kmer::set_kmer, ext/bifrost/src/Kmer.cpp:92-107
  const size_t x = (c & 4) >> 1;
  return (int)(x + ((x ^ (c & 2)) >> 1));
}
uint64_t pack(const char* s, int k) {
  uint64_t v = 0;
  for (int i = 0; i < k; ++i) v = (v << 2) | (uint64_t)code_of(s[i]);
  return v;
}
uint64_t twin(uint64_t v, int k) {   // Kmer::twin
  uint64_t r = 0;
  for (int i = 0; i < k; ++i) {
    r = (r << 2) | (3 - (v & 3));
    v >>= 2;
  }
  return r;
}

void read_roaring_portable(Reader& r, TidSet& out) {   // roaring.c:10405-10450
  const uint32_t cookie = r.get<uint32_t>();
  uint32_t n;
  std::vector<uint8_t> runflag;
  bool hasrun = false;
  if ((cookie & 0xFFFF) == 12347) {
    hasrun = true;
    n = (cookie >> 16) + 1;
    const uint8_t* b = r.take((n + 7) / 8);
    runflag.assign(b, b + (n + 7) / 8);
  } else if (cookie == 12346) {
    n = r.get<uint32_t>();
  } else {
    throw std::runtime_error("oracle: bad roaring cookie");
  }
  std::vector<uint16_t> keys(n), cards(n);
  for (uint32_t i = 0; i < n; ++i) {
    keys[i] = r.get<uint16_t>();
    cards[i] = r.get<uint16_t>();
  }
  if (!hasrun || n >= 4) r.take(4 * (size_t)n);
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t hi = (uint32_t)keys[i] << 16;
    const uint32_t card = (uint32_t)cards[i] + 1;
    if (hasrun && ((runflag[i / 8] >> (i % 8)) & 1)) {
      const uint16_t nr = r.get<uint16_t>();
      for (uint16_t j = 0; j < nr; ++j) {
        const uint32_t s = r.get<uint16_t>();
        const uint32_t l = r.get<uint16_t>();
        for (uint32_t v = s; v <= s + l; ++v) out.push_back(hi | v);
      }
    } else if (card > 4096) {
      for (uint32_t w = 0; w < 1024; ++w) {
        const uint64_t x = r.get<uint64_t>();
        for (int b = 0; b < 64; ++b)
          if ((x >> b) & 1) out.push_back(hi | (w * 64 + b));
      }
    } else {
      for (uint32_t j = 0; j < card; ++j) out.push_back(hi | r.get<uint16_t>());
    }
  }
}

void read_roaring_native(const uint8_t* p, size_t n, TidSet& out) {   // roaring.c:8554-8568
  Reader r{p, p + n};
  const uint8_t tag = r.get<uint8_t>();
  if (tag == 1) {
    const uint32_t card = r.get<uint32_t>();
    for (uint32_t i = 0; i < card; ++i) out.push_back(r.get<uint32_t>());
  } else if (tag == 2) {
    read_roaring_portable(r, out);
  } else {
    throw std::runtime_error("oracle: bad roaring tag");
  }
}

OIndex* load_index(const char* path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("oracle: cannot open index");
  std::vector<uint8_t> buf((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  Reader r{buf.data(), buf.data() + buf.size()};
  OIndex* ix = new OIndex();
  if (r.get<uint64_t>() != 13) throw std::runtime_error("oracle: index version != 13");
  uint64_t dbg_bytes = r.get<uint64_t>() & (~0ULL >> 1);
  {
    Reader g{r.p, r.p + dbg_bytes};
    r.take(dbg_bytes);
    if ((g.get<uint64_t>() >> 32) != 0x7e215f3fULL) throw std::runtime_error("oracle: bad graph header");
    ix->k = g.get<int32_t>();
    g.get<int32_t>();
    const uint64_t nl = g.get<uint64_t>();
    for (uint64_t i = 0; i < nl; ++i) {   // CompressedSequence::read, CompressedSequence.cpp:283-309
      const uint64_t len = g.get<uint64_t>();
      const uint8_t* d = g.take((len + 3) / 4);
      Unitig u;
      u.kind = 0;
      u.seq.resize(len);
      for (uint64_t j = 0; j < len; ++j) u.seq[j] = "ACGT"[(d[j >> 2] >> ((j & 3) << 1)) & 3];   // getChar :311-314
      ix->unitigs.push_back(std::move(u));
    }
    for (int kind = 1; kind <= 2; ++kind) {   // km_unitigs, then h_kmers_ccov (IO.tcc:1697-1727)
      const uint64_t n = g.get<uint64_t>();
      for (uint64_t i = 0; i < n; ++i) {
        const uint64_t w = g.get<uint64_t>() >> (64 - 2 * ix->k);
        Unitig u;
        u.kind = kind;
        u.seq.resize(ix->k);
        for (int j = 0; j < ix->k; ++j) u.seq[j] = "ACGT"[(w >> (2 * (ix->k - 1 - j))) & 3];
        ix->unitigs.push_back(std::move(u));
      }
    }
  }
  r.take(r.get<uint64_t>());   // BBHash MPHF: Bifrost-internal, not needed by a flat dictionary
  const uint64_t dlist_n = r.get<uint64_t>();
  r.get<uint64_t>();
  for (uint64_t i = 0; i < dlist_n; ++i) {
    const uint64_t w = r.get<uint64_t>() >> (64 - 2 * ix->k);     // Kmer: left-aligned 2-bit words (Kmer.cpp:92-107)
    ix->d_list.insert(w);
    if (i == 0) ix->dummy_dfk = w;
  }
  const int k = ix->k;
  for (uint32_t u = 0; u < ix->unitigs.size(); ++u) {
    const std::string& s = ix->unitigs[u].seq;
    for (uint32_t d = 0; d + k <= s.size(); ++d) {
      const uint64_t f = pack(s.data() + d, k), t = twin(f, k);
      ix->kmers[f < t ? f : t] = KmerLoc{u, d, f < t};
    }
  }
  const uint64_t n_nodes = r.get<uint64_t>();
  for (uint64_t i = 0; i < n_nodes; ++i) {
    const char* head = (const char*)r.take(k);
    const uint32_t sz = r.get<uint32_t>();
    Reader nr{r.p, r.p + sz};
    r.take(sz);
    const uint64_t f = pack(head, k), t = twin(f, k);
    auto it = ix->kmers.find(f < t ? f : t);
    if (it == ix->kmers.end()) throw std::runtime_error("oracle: node head not in graph");
    Unitig& un = ix->unitigs[it->second.unitig];
    nr.get<uint32_t>();   // Node::id
    const uint8_t flag = nr.get<uint8_t>();   // BlockArray::deserialize, BlockArray.hpp:441-471
    const uint64_t nb = flag == 0 ? 0 : (flag == 1 ? 1 : nr.get<uint64_t>());
    for (uint64_t b = 0; b < nb; ++b) {
      Block blk;
      blk.lb = nr.get<uint32_t>();
      blk.ub = nr.get<uint32_t>();
      const uint64_t rb = nr.get<uint64_t>();   // SparseVector::deserialize, SparseVector.tcc:424-507
      read_roaring_native(nr.take(rb), rb, blk.tids);
      const uint64_t vs = nr.get<uint64_t>();
      for (uint64_t j = 0; j < vs; ++j) {
        const uint64_t pb = nr.get<uint64_t>();
        TidSet ps;
        read_roaring_native(nr.take(pb), pb, ps);
        const uint32_t mn = ps.front(), mx = ps.back();
        const bool smin = (mn & 0x7FFFFFFF) == mn, smax = (mx & 0x7FFFFFFF) == mx;
        blk.sense.push_back(smin != smax ? 2 : (smin ? 1 : 0));
        blk.pos.push_back(ps);
      }
      un.blocks.push_back(std::move(blk));
    }
  }
  int32_t nt = r.get<int32_t>();
  nt -= (int32_t)dlist_n;     // the stored count includes one pseudo-target per D-list k-mer (src/KmerIndex.cpp:1470-1472)
  for (int32_t i = 0; i < nt; ++i) ix->target_len.push_back((uint32_t)r.get<int32_t>());
  for (int32_t i = 0; i < nt; ++i) {
    const uint64_t n = r.get<uint64_t>();
    const char* s = (const char*)r.take(n);
    ix->target_name.push_back(std::string(s, strnlen(s, n)));
  }
  const uint64_t ob = r.get<uint64_t>();
  Reader orr{r.p, r.p + ob};
  read_roaring_portable(orr, ix->onlist);
  return ix;
}

// ---------------------------------------------------------------------------------------------
// dbg.find + Node accessors
// ---------------------------------------------------------------------------------------------
struct Um {           // const_UnitigMap<Node> fields used on this path
  bool isEmpty = true;
  uint32_t unitig = 0;
  uint32_t dist = 0;
  bool strand = false;
  int block = -1;     // index of the EC block holding dist (BlockArray::operator[] / get_block_at)
};

Um find(const OIndex& ix, uint64_t km) {   // CompactedDBG::find, CompactedDBG.tcc:999-1119
  const uint64_t tw = twin(km, ix.k);
  const bool is_rep = km < tw;
  auto it = ix.kmers.find(is_rep ? km : tw);
  Um um;
  if (it == ix.kmers.end()) return um;
  um.isEmpty = false;
  um.unitig = it->second.unitig;
  um.dist = it->second.dist;
  um.strand = (is_rep == it->second.fwd_is_rep);
  const auto& blocks = ix.unitigs[um.unitig].blocks;
  // upper_bound on lb, then step back (BlockArray.hpp:306-322)
  int b = -1;
  for (size_t i = 0; i < blocks.size(); ++i)
    if (blocks[i].lb <= um.dist) b = (int)i;
  um.block = b;
  return um;
}
const Block& blk(const OIndex& ix, const Um& um) { return ix.unitigs[um.unitig].blocks[um.block]; }
bool same_unitig_ec(const OIndex& ix, const Um& a, const Um& b) {
  return a.unitig == b.unitig && blk(ix, a).tids == blk(ix, b).tids;   // isSameReferenceUnitig && ec == ec
}

// KmerIterator (ext/bifrost/src/KmerIterator.cpp:6-63) as a literal state machine over the
// NUL-terminated string.
struct KIt {
  const char* str = nullptr;
  bool invalid = true;
  int pos_s = 0, pos_e = 0;
  uint64_t km = 0;
  int p = 0;
  int k = 0;
  static bool isDNA(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }   // Common.hpp:45-50 after & 0xDF
  void inc() {
    if (invalid) return;
    while (str[pos_e] != '\0') {
      const char c = str[pos_e] & 0xDF;
      if (isDNA(c)) {
        if (pos_s + k - 1 == pos_e) {
          km = pack(str + pos_s, k);
          p = pos_s;
          ++pos_s;
          ++pos_e;
          return;
        }
      } else {
        pos_s = pos_e + 1;
      }
      ++pos_e;
    }
    invalid = true;
  }
  void add(int len) {
    if (invalid) return;
    if (len == 1) inc();
    else if (len > 1) {
      const int next_pos_e = pos_e + len - 1;
      while (pos_e < next_pos_e && str[pos_e] != '\0') ++pos_e;
      if (str[pos_e] != '\0') {
        pos_s = pos_e - k + 1;
        pos_e = pos_s;
        inc();
      } else {
        invalid = true;
      }
    }
  }
  static KIt begin(const char* s, int k) {
    KIt it;
    it.str = s;
    it.invalid = false;
    it.k = k;
    it.inc();
    return it;
  }
};

typedef std::vector<std::pair<Um, int>> HitVec;

// diagnostics: lookups of match() by kind -- 0 main hit, 1 main miss, 2 jump landing exactly where the
// read would be if it followed the unitig, 3 jump on an absent k-mer, 4 other jumps, 5 middle, 6 back-off;
// [7] = largest number of distinct non-empty EC sets hit by one fragment
static uint64_t g_kind[8];

// KmerIndex::match, src/KmerIndex.cpp:1698-1940 (default flags; D-list: 1818-1826 and 1928-1939)
void match(const OIndex& ix, const char* s, int l, HitVec& v, bool partial, uint64_t* n_find) {
  const int k = ix.k;
  auto rep = [&](uint64_t km) { const uint64_t t = twin(km, k); return km < t ? km : t; };
  auto dlist_tail = [&]() {   // lines 1928-1939
    if (ix.d_list.empty() || !(v.size() > 0 || !partial)) return;
    const Um um_dummy = find(ix, ix.dummy_dfk);
    for (KIt kd = KIt::begin(s, k); !kd.invalid; kd.inc())
      if (ix.d_list.count(rep(kd.km))) { v.push_back({um_dummy, kd.p}); break; }
  };
  KIt kit = KIt::begin(s, k);
  bool backOff = false;
  int nextPosOuter = 0;   // the outer `nextPos` (line 1748) is never updated: the inner one shadows it
  TidSet rtmp;
  auto and_partial = [&](const Um& um) -> bool {   // lines 1760-1772 / 1853-1864 / 1903-1914
    const TidSet& r2 = blk(ix, um).tids;
    if (rtmp.empty()) {
      if (!r2.empty()) rtmp = r2;
    } else {
      if (!r2.empty()) {
        TidSet t;
        std::set_intersection(rtmp.begin(), rtmp.end(), r2.begin(), r2.end(), std::back_inserter(t));
        rtmp.swap(t);
      }
      if (rtmp.empty()) return false;
    }
    return true;
  };
  for (; !kit.invalid; kit.inc()) {
    ++*n_find;
    const Um um = find(ix, kit.km);
    const int pos = kit.p;
    ++g_kind[um.isEmpty ? 1 : 0];
    if (!um.isEmpty) {
      if (partial && !and_partial(um)) { v.clear(); return; }
      v.push_back({um, kit.p});
      const Block& b = blk(ix, um);
      const size_t contig_start = b.lb;
      const size_t contig_length = b.ub - contig_start;
      const bool forward = um.strand;
      const int dist = forward ? (int)(contig_length - 1 - (um.dist - contig_start)) : (int)(um.dist - contig_start);
      if (dist >= 2) {
        int nextPos = pos + dist;
        if (pos + dist >= l - k) nextPos = l - k;
        KIt kit2 = kit;
        kit2.add(nextPos - pos);
        if (!kit2.invalid) {
          ++*n_find;
          const Um um2 = find(ix, kit2.km);
          {
            const long expect = forward ? (long)um.dist + (kit2.p - pos) : (long)um.dist - (kit2.p - pos);
            if (um2.isEmpty) ++g_kind[3];
            else if (um2.unitig == um.unitig && (long)um2.dist == expect && um2.strand == um.strand) ++g_kind[2];
            else ++g_kind[4];
          }
          bool found2 = false;
          int found2pos = pos + dist;
          if (um2.isEmpty) {
            found2 = true;
            found2pos = pos;
          } else if (same_unitig_ec(ix, um, um2)) {
            found2 = true;
            found2pos = pos + dist;
          }
          if (found2) {
            const bool is_in_dlist = um2.isEmpty && !ix.d_list.empty() && ix.d_list.count(rep(kit2.km));   // :1818
            if (found2pos >= l - k) {
              v.push_back({um, l - k});
              if (partial && is_in_dlist) { v.push_back({find(ix, ix.dummy_dfk), kit2.p}); return; }
              break;
            } else {
              v.push_back({um, found2pos});
              kit = kit2;
              if (partial && is_in_dlist) { v.push_back({find(ix, ix.dummy_dfk), kit2.p}); return; }
            }
          } else {
            bool foundMiddle = false;
            if (dist > 4) {
              const int middlePos = (pos + nextPos) / 2;
              int found3pos = pos + dist;
              KIt kit3 = kit;
              kit3.add(middlePos - pos);
              if (!kit3.invalid) {
                ++*n_find;
                ++g_kind[5];
                const Um um3 = find(ix, kit3.km);
                if (!um3.isEmpty) {
                  if (same_unitig_ec(ix, um, um3)) {
                    foundMiddle = true;
                    found3pos = middlePos;
                  } else if (same_unitig_ec(ix, um2, um3)) {
                    foundMiddle = true;
                    found3pos = pos + dist;
                  }
                }
                if (foundMiddle) {
                  if (partial && !and_partial(um3)) { v.clear(); return; }
                  v.push_back({um3, found3pos});
                  if (nextPos >= l - k) break;
                  else kit = kit2;
                }
              }
            }
            if (!foundMiddle) {
              kit.inc();
              backOff = true;
            }
          }
        } else {
          break;
        }
      }
    }
    if (backOff) {   // lines 1892-1925
      for (int j = 0; !kit.invalid; kit.inc(), ++j) {
        if (j == 0) {
          ++*n_find;
          ++g_kind[6];
          const Um um4 = find(ix, kit.km);
          if (!um4.isEmpty) {
            if (partial && !and_partial(um4)) { v.clear(); return; }
            v.push_back({um4, kit.p});
          }
        }
        if (kit.p >= nextPosOuter) {
          backOff = false;
          break;
        }
      }
    }
  }
  dlist_tail();
}

TidSet intersect(const TidSet& a, const TidSet& b) {
  TidSet t;
  std::set_intersection(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(t));
  return t;
}

// MinCollector::intersectECs, src/MinCollector.cpp:425-496 (min_range = 1; no shade, no D-list)
TidSet intersectECs(const OIndex& ix, HitVec& v, int k) {
  TidSet r;
  if (v.empty()) return r;
  std::stable_sort(v.begin(), v.end(), [&](const std::pair<Um, int>& a, const std::pair<Um, int>& b) {
    // the reference's comparator mixes two keys and is not a strict weak order; only the
    // resulting SET matters downstream, which no ordering can change.  Sorting by (unitig, pos)
    // keeps equal (unitig, EC) entries adjacent like the reference intends.
    if (a.first.unitig != b.first.unitig) return a.first.unitig < b.first.unitig;
    return a.second < b.second;
  });
  r = blk(ix, v[0].first).tids;
  bool found_nonempty = !r.empty();
  TidSet lastEC = r;
  TidSet ec;
  for (size_t i = 1; i < v.size(); ++i) {
    if (!found_nonempty) {
      r = blk(ix, v[i].first).tids;
      found_nonempty = !r.empty();
    }
    if (!same_unitig_ec(ix, v[i].first, v[i - 1].first)) {
      ec = blk(ix, v[i].first).tids;
      if (!(ec == lastEC) && !ec.empty()) {
        r = intersect(r, ec);
        if (r.empty()) return r;
        lastEC = ec;
      }
    }
  }
  int minpos = INT32_MAX, maxpos = 0;
  for (auto& x : v) {
    minpos = std::min(minpos, x.second);
    maxpos = std::max(maxpos, x.second);
  }
  if ((maxpos - minpos + k) < 1) return TidSet();
  return r;
}

// MinCollector::intersectKmers, src/MinCollector.cpp:160-218
int intersectKmers(const OIndex& ix, HitVec& v1, HitVec& v2, TidSet& r) {
  const TidSet u1 = intersectECs(ix, v1, ix.k), u2 = intersectECs(ix, v2, ix.k);
  if (u1.empty() && u2.empty()) return -1;
  if (u1.empty()) {
    if (v1.empty()) r = u2; else return -1;
  } else if (u2.empty()) {
    if (v2.empty()) r = u1; else return -1;
  } else {
    r = intersect(u1, u2);
  }
  if (r.empty()) return -1;
  return 1;
}

std::pair<Um, int> findFirstMappingKmer(const HitVec& v) {   // src/ProcessReads.cpp:45-59
  Um um;
  int p = -1;
  if (!v.empty()) {
    um = v[0].first;
    p = v[0].second;
    for (auto& x : v)
      if (x.second < p) { um = x.first; p = x.second; }
  }
  return {um, p};
}

// doStrandSpecificity (non-comprehensive), src/ProcessReads.cpp:61-124.  strand: 1 = FR, 2 = RF
void doStrandSpecificity(const OIndex& ix, TidSet& u, int strand, const HitVec& v, const HitVec& v2) {
  for (int mate = 0; mate < 2; ++mate) {
    const HitVec& vv = mate == 0 ? v : v2;
    if (vv.empty()) continue;
    const bool want = mate == 0 ? (strand == 1) : (strand == 2);
    const Um um = findFirstMappingKmer(vv).first;
    const Block& b = blk(ix, um);   // get_leading_vals(um.dist).back() == the block holding dist
    u = intersect(u, b.tids);
    TidSet vtmp;
    for (uint32_t tr : u) {
      const size_t rank = std::lower_bound(b.tids.begin(), b.tids.end(), tr) - b.tids.begin();
      const uint8_t sense = b.sense[rank];
      if ((um.strand == (bool)sense) == want || sense == 2) vtmp.push_back(tr);
    }
    if (vtmp.size() < u.size()) u = vtmp;
  }
}

// KmerIndex::findPosition, src/KmerIndex.cpp:2188-2292, restated literally on the flat block list
// (ecs = leading blocks, get_mc_contig = block bounds).  Returns {position, sense}.
std::pair<int, bool> findPosition(const OIndex& ix, uint32_t tr, const Um& um, int p) {
  const int k = ix.k;
  const bool csense = um.strand;
  const Unitig& un = ix.unitigs[um.unitig];
  const std::vector<Block>& B = un.blocks;
  const int64_t um_size = (int64_t)un.seq.size();
  int cur = um.block;                                   // ecs.size()-1
  const Block& v_ec = B[cur];
  const size_t rank = std::lower_bound(v_ec.tids.begin(), v_ec.tids.end(), tr) - v_ec.tids.begin();
  const uint32_t rawpos = v_ec.pos[rank].front();       // v_ec.get(tr, true).minimum()
  const int trpos = (int)(rawpos & 0x7FFFFFFF);
  const bool trsense = (rawpos == (uint32_t)trpos);
  auto contains = [&](int i) { return std::binary_search(B[i].tids.begin(), B[i].tids.end(), tr); };
  std::pair<int, bool> ret;
  if (trsense) {
    if (csense) {
      size_t padding = 0;
      if (trpos == 0) {
        int mc = cur;                                   // block whose bounds `mc` holds
        for (int i = cur - 1; i >= 0; i--) {
          if (!contains(i)) { padding = B[mc].lb; break; }
          mc = mc - 1;                                  // get_mc_contig(mc.first-1): the previous block
        }
      }
      ret = {(int)(static_cast<int64_t>(trpos) - p + static_cast<int64_t>(um.dist) + 1 - (int64_t)padding), csense};   // Case I
    } else {
      const int64_t initial = B[cur].ub;
      int right_one = 0, left_one = 0;
      int mc = cur;
      for (int i = cur; i >= 0; i--) {
        if (i == cur) right_one = B[mc].ub;
        if (!contains(i)) { left_one = B[mc].ub; break; }
        else if (i == 0) left_one = 0;
        mc = mc > 0 ? mc - 1 : (int)B.size() - 1;       // below 0: "returns the right-most block"
      }
      const int64_t padding = -(left_one + right_one - um_size + k - 1);
      ret = {(int)(trpos + p + k - (um_size - k - um.dist) + initial - 1 + padding), csense};   // Case III
    }
  } else {
    int left_one = 0, right_one = 0, unmapped_len = 0;
    bool found_first_mapped = false;
    for (size_t i = 0; i < B.size(); i++) {             // get_leading_vals(-1): every block
      if (!contains((int)i) && found_first_mapped) {
        if (unmapped_len == 0) left_one = B[i].lb;
        right_one = B[i].ub;
        unmapped_len += (B[i].ub - B[i].lb);
      }
      if (contains((int)i)) found_first_mapped = true;
    }
    if (csense) {
      int64_t start = 0;
      start -= right_one - left_one;
      start += um_size - k;
      ret = {(int)(trpos + (static_cast<int64_t>(-((int64_t)um.dist - start))) + k + p), !csense};   // Case IV
    } else {
      unmapped_len = right_one - left_one;
      const int64_t padding = um_size - um.dist - (unmapped_len) - k + 1;
      ret = {(int)(trpos + padding - p), !csense};      // case II
    }
  }
  return ret;
}

// KmerIndex::mapPair, src/KmerIndex.cpp:1622-1693
int mapPair(const OIndex& ix, const char* s1, const char* s2, uint64_t* n_find) {
  const int k = ix.k;
  int p1 = -1, p2 = -1;
  Um um1, um2;
  bool found1 = false, found2 = false;
  for (KIt kit = KIt::begin(s1, k); !kit.invalid; kit.inc()) {
    ++*n_find;
    um1 = find(ix, kit.km);
    if (!um1.isEmpty) {
      found1 = true;
      p1 = um1.strand ? (int)um1.dist - kit.p : (int)um1.dist + k + kit.p;
      break;
    }
  }
  if (!found1) return -1;
  for (KIt kit = KIt::begin(s2, k); !kit.invalid; kit.inc()) {
    ++*n_find;
    um2 = find(ix, kit.km);
    if (!um2.isEmpty) {
      found2 = true;
      p2 = um2.strand ? (int)um2.dist - kit.p : (int)um2.dist + k + kit.p;
      break;
    }
  }
  if (!found2) return -1;
  if (!same_unitig_ec(ix, um1, um2)) return -1;
  if (!(um1.strand ^ um2.strand)) return -1;
  if (blk(ix, um1).ub != blk(ix, um2).ub) return -1;
  return p1 > p2 ? p1 - p2 : p2 - p1;
}

// ---------------------------------------------------------------------------------------------
// A run: ReadProcessor::processBuffer + MasterProcessor::update with -t 1
// ---------------------------------------------------------------------------------------------
struct ORun {
  OIndex* ix;
  int paired, strand;   // strand: 0 none, 1 FR, 2 RF
  int fp_fl = -1;       // >= 0: !single_overhang && has_mean_fl, with (int) mean fragment length
  std::map<TidSet, int32_t> ecmapinv;       // EC set -> id (ids = insertion order, as with the ankerl map)
  std::vector<TidSet> ecs;
  std::vector<uint32_t> counts;
  std::vector<uint32_t> flens = std::vector<uint32_t>(1000, 0);
  int tlencount = 0;
  uint64_t numreads = 0, n_find = 0;
};

}  // namespace

extern "C" {

void* oracle_index_load(const char* path, char* err, int errlen) {
  try {
    return load_index(path);
  } catch (const std::exception& e) {
    if (err) snprintf(err, errlen, "%s", e.what());
    return nullptr;
  }
}
void oracle_index_free(void* ix) { delete (OIndex*)ix; }
int oracle_index_k(void* ix) { return ((OIndex*)ix)->k; }
uint32_t oracle_index_n_targets(void* ix) { return (uint32_t)((OIndex*)ix)->target_len.size(); }
uint64_t oracle_index_n_kmers(void* ix) { return ((OIndex*)ix)->kmers.size(); }
uint32_t oracle_index_n_unitigs(void* ix) { return (uint32_t)((OIndex*)ix)->unitigs.size(); }
uint64_t oracle_index_n_blocks(void* ix) {
  uint64_t n = 0;
  for (auto& u : ((OIndex*)ix)->unitigs) n += u.blocks.size();
  return n;
}
void oracle_index_target_lens(void* ix, uint32_t* out) {
  auto& v = ((OIndex*)ix)->target_len;
  memcpy(out, v.data(), v.size() * 4);
}
const char* oracle_index_target_name(void* ix, uint32_t i) { return ((OIndex*)ix)->target_name[i].c_str(); }

void* oracle_run_create(void* ix, int paired, int strand) {
  ORun* r = new ORun();
  r->ix = (OIndex*)ix;
  r->paired = paired;
  r->strand = strand;
  return r;
}
void oracle_run_free(void* r) { delete (ORun*)r; }
void oracle_run_set_fp(void* r, int fl) { ((ORun*)r)->fp_fl = fl; }

// One batch, reads given like kb_pseudoalign_batch.  ec_out: EC id per fragment (ids final: -t 1
// assigns them in first-occurrence order) or -1.  collect_fld mirrors opt.fld == 0.
void oracle_pseudoalign_batch(void* run, const char* bases, const uint32_t* off, uint32_t n_reads, uint32_t fixed_len,
                              int collect_fld, int32_t* ec_out) {
  ORun& R = *(ORun*)run;
  const OIndex& ix = *R.ix;
  const bool paired = R.paired != 0;
  HitVec v1, v2;
  std::string s1, s2;
  // ProcessReads.cpp:981-1017: the goal is fixed at batch start
  bool findFragmentLength = collect_fld && R.tlencount < 10000;
  int flengoal = findFragmentLength ? 10000 - R.tlencount : 0;
  int local_tlen = 0;
  uint32_t frag = 0;
  for (uint32_t i = 0; i < n_reads; ++i, ++frag) {
    auto fetch = [&](uint32_t idx, std::string& s) {
      const uint64_t b = off ? off[idx] : (uint64_t)idx * fixed_len;
      const uint64_t e = off ? off[idx + 1] : b + fixed_len;
      s.assign(bases + b, bases + e);
    };
    fetch(i, s1);
    if (paired) { ++i; fetch(i, s2); }
    ++R.numreads;
    v1.clear();
    v2.clear();
    TidSet u;
    match(ix, s1.c_str(), (int)s1.size(), v1, !paired, &R.n_find);
    if (paired) match(ix, s2.c_str(), (int)s2.size(), v2, !paired, &R.n_find);
    intersectKmers(ix, v1, v2, u);
    {   // diagnostics: distinct non-empty EC sets hit by the fragment (g_kind[7] = max over the run)
      std::vector<const TidSet*> seen;
      for (const HitVec* v : {&v1, &v2})
        for (const auto& h : *v) {
          const TidSet& t = blk(ix, h.first).tids;
          if (t.empty()) continue;
          bool dup = false;
          for (const TidSet* s : seen) dup = dup || (*s == t);
          if (!dup) seen.push_back(&t);
        }
      if (seen.size() > g_kind[7]) g_kind[7] = seen.size();
    }
    u = intersect(u, ix.onlist);                                        // :1072
    if (R.fp_fl >= 0 && !u.empty() && (!paired || v1.empty() || v2.empty())) {   // :1095-1136
      TidSet vtmp;
      const int fl = R.fp_fl;
      int p = -1;
      Um um;
      if (!v1.empty()) { auto res = findFirstMappingKmer(v1); um = res.first; p = res.second; }
      if (!v2.empty()) { auto res = findFirstMappingKmer(v2); um = res.first; p = res.second; }
      for (uint32_t tr : u) {
        auto x = findPosition(ix, tr, um, p);
        bool add = false;
        if (x.second && x.first + fl <= (int)ix.target_len[tr]) add = true;
        if (!x.second && x.first - fl >= 0) add = true;
        if (add) vtmp.push_back(tr);
      }
      if (vtmp.size() < u.size()) u = vtmp;
    }
    if (R.strand != 0 && !u.empty()) doStrandSpecificity(ix, u, R.strand, v1, v2);   // :1138-1145
    int32_t ec = -1;
    if (!u.empty()) {
      auto it = R.ecmapinv.find(u);
      if (it == R.ecmapinv.end()) {
        ec = (int32_t)R.ecs.size();
        R.ecmapinv[u] = ec;
        R.ecs.push_back(u);
        R.counts.push_back(1);
      } else {
        ec = it->second;
        ++R.counts[ec];
      }
      if (findFragmentLength && flengoal > 0 && paired && u.size() == 1 && !v1.empty() && !v2.empty()) {   // :1174-1181
        const int tl = mapPair(ix, s1.c_str(), s2.c_str(), &R.n_find);
        if (0 < tl && tl < 1000) {
          ++R.flens[tl];
          --flengoal;
          ++local_tlen;
        }
      }
    }
    if (ec_out) ec_out[frag] = ec;
  }
  R.tlencount += local_tlen;
}

uint32_t oracle_n_ecs(void* run) { return (uint32_t)((ORun*)run)->ecs.size(); }
uint64_t oracle_n_ec_entries(void* run) {
  uint64_t n = 0;
  for (auto& e : ((ORun*)run)->ecs) n += e.size();
  return n;
}
uint64_t oracle_n_find(void* run) { return ((ORun*)run)->n_find; }
void oracle_kind_counts(uint64_t* out) { for (int i = 0; i < 8; ++i) out[i] = g_kind[i]; }
void oracle_ec_table(void* run, uint64_t* off, uint32_t* tids, uint32_t* counts) {
  ORun& R = *(ORun*)run;
  uint64_t o = 0;
  for (size_t e = 0; e < R.ecs.size(); ++e) {
    off[e] = o;
    for (uint32_t t : R.ecs[e]) tids[o++] = t;
    counts[e] = R.counts[e];
  }
  off[R.ecs.size()] = o;
}
void oracle_get_flens(void* run, uint32_t* out) { memcpy(out, ((ORun*)run)->flens.data(), 4000); }

// compute_mean_frag_lens_trunc (MinCollector.cpp:629-651) or init_mean_fl_trunc/trunc_gaussian_fld
// (MinCollector.cpp:25-35, weights.cpp:248-271)
void oracle_mean_fl_trunc(const uint32_t* flens, double fld_mean, double fld_sd, double* out) {
  const int MAXF = 1000;
  for (int i = 0; i < MAXF; ++i) out[i] = 0.0;
  if (fld_mean == 0.0) {
    std::vector<int> counts(MAXF, 0);
    std::vector<double> mass(MAXF, 0.0);
    counts[0] = flens[0];
    for (size_t i = 1; i < (size_t)MAXF; ++i) {
      mass[i] = static_cast<double>(flens[i] * i) + mass[i - 1];
      counts[i] = flens[i] + counts[i - 1];
      if (counts[i] > 0) out[i] = mass[i] / static_cast<double>(counts[i]);
    }
  } else {
    double total_mass = 0.0, total_density = 0.0;
    for (size_t i = 0; i < (size_t)MAXF; ++i) {
      double x = static_cast<double>(i);
      x = (x - fld_mean) / fld_sd;
      const double cur_density = std::exp(-0.5 * x * x) / fld_sd;
      total_mass += cur_density * i;
      total_density += cur_density;
      if (total_mass > 0) out[i] = total_mass / total_density;
    }
  }
}

// get_frag_len_means + calc_eff_lens (weights.cpp:7-28, 58-79)
void oracle_eff_lens(const uint32_t* lens, uint32_t T, const double* fl_trunc, double* eff) {
  for (uint32_t t = 0; t < T; ++t) {
    const double mean = lens[t] >= 1000 ? fl_trunc[999] : fl_trunc[lens[t]];
    const double cur_len = static_cast<double>(lens[t]);
    double e = cur_len - mean + 1;
    if (e < 1.0) e = cur_len;
    eff[t] = e;
  }
}

// EMAlgorithm::run (EMAlgorithm.h:95-221) with calc_weights (weights.cpp:220-246).
// counts = the counts the EM fits (bootstrap sample or true counts); counts_w = tc.counts used for
// the weights.  Returns the number of rounds the reference prints.
int oracle_em(uint32_t n_ecs, const uint64_t* off, const uint32_t* tids, const uint32_t* counts,
              const uint32_t* counts_w, uint32_t T, const double* eff, int n_iter, int min_rounds, double* alpha_out) {
  std::vector<double> alpha(T, 1.0 / T), next_alpha(T, 0.0);
  std::vector<std::vector<double>> wmap(n_ecs);
  for (uint32_t e = 0; e < n_ecs; ++e)
    for (uint64_t j = off[e]; j < off[e + 1]; ++j) wmap[e].push_back(static_cast<double>(counts_w[e]) / eff[tids[j]]);
  const double alpha_limit = 1e-7, alpha_change_limit = 1e-2, alpha_change = 1e-2;
  const double TOLERANCE = std::numeric_limits<double>::denorm_min();
  bool finalRound = false;
  int i;
  for (i = 0; i < n_iter; ++i) {
    for (uint32_t e = 0; e < n_ecs; ++e)
      if (off[e + 1] - off[e] == 1) next_alpha[tids[off[e]]] = counts[e];
    for (uint32_t e = 0; e < n_ecs; ++e) {
      const uint64_t n = off[e + 1] - off[e];
      if (n == 1) continue;
      double denom = 0.0;
      if (counts[e] == 0) continue;
      const std::vector<double>& wv = wmap[e];
      const uint32_t* trs = tids + off[e];
      for (uint64_t t = 0; t < n; ++t) denom += alpha[trs[t]] * wv[t];
      if (denom < TOLERANCE) continue;
      const double countNorm = counts[e] / denom;
      for (uint64_t t = 0; t < n; ++t) next_alpha[trs[t]] += (wv[t] * alpha[trs[t]]) * countNorm;
    }
    bool stopEM = false;
    int chcount = 0;
    for (uint32_t ec = 0; ec < T; ++ec) {
      if (next_alpha[ec] > alpha_change_limit && (std::fabs(next_alpha[ec] - alpha[ec]) / next_alpha[ec]) > alpha_change) chcount++;
      alpha[ec] = next_alpha[ec];
      next_alpha[ec] = 0.0;
    }
    if (chcount == 0 && i > min_rounds) stopEM = true;
    if (finalRound) break;
    if (stopEM) {
      finalRound = true;
      for (uint32_t ec = 0; ec < T; ++ec)
        if (alpha[ec] < alpha_limit / 10.0) alpha[ec] = 0.0;
    }
  }
  memcpy(alpha_out, alpha.data(), T * sizeof(double));
  return i;
}

// counts_to_tpm, PlaintextWriter.cpp:5-27
void oracle_tpm(const double* est, const double* eff, uint32_t T, double* tpm) {
  double total_mass = 0.0;
  for (uint32_t i = 0; i < T; ++i) {
    tpm[i] = est[i] / eff[i];
    total_mass += tpm[i];
  }
  for (uint32_t i = 0; i < T; ++i) tpm[i] = (tpm[i] / total_mass) * 1e6;
}

// Bootstrap seeds (main.cpp:2746-2752) and Multinomial::sample (Multinomial.hpp:33-51) using the
// very same standard-library objects as the reference.
void oracle_bootstrap_sample(const uint32_t* counts, uint32_t n_ecs, uint64_t seed, int b, uint32_t* samp) {
  std::mt19937_64 rand;
  rand.seed(seed);
  size_t s = 0;
  for (int i = 0; i <= b; ++i) s = rand();
  std::vector<uint32_t> c(counts, counts + n_ecs);
  std::default_random_engine gen(s);
  std::discrete_distribution<int> dd(c.begin(), c.end());
  int n = 0;
  for (auto x : c) n += x;
  for (uint32_t e = 0; e < n_ecs; ++e) samp[e] = 0;
  for (int i = 0; i < n; ++i) ++samp[dd(gen)];
}

}  // extern "C"
