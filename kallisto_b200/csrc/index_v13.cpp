// Loader for the kallisto v13 index file -> kb::FlatIndex.  See index_v13.hpp.
#include "index_v13.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <unordered_map>

namespace kb {

namespace {

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  void need(size_t n) const {
    if ((size_t)(end - p) < n) throw std::runtime_error("kallisto index: truncated file");
  }
  template <class T> T get() {
    need(sizeof(T));
    T v;
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  const uint8_t* bytes(size_t n) {
    need(n);
    const uint8_t* r = p;
    p += n;
    return r;
  }
};

inline int base_code(char c) {  // Kmer::set_kmer, ext/bifrost/src/Kmer.cpp:92-107
  const unsigned x = (c & 4) >> 1;
  return x + ((x ^ (c & 2)) >> 1);
}

struct Mmap {
  const uint8_t* data = nullptr;
  size_t size = 0;
  int fd = -1;
  explicit Mmap(const std::string& path) {
    fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("Error: index input file could not be opened! (" + path + ")");
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); throw std::runtime_error("kallisto index: fstat failed"); }
    size = (size_t)st.st_size;
    if (size == 0) { close(fd); throw std::runtime_error("kallisto index: empty file"); }
    void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { close(fd); throw std::runtime_error("kallisto index: mmap failed"); }
    data = (const uint8_t*)m;
  }
  ~Mmap() {
    if (data) munmap((void*)data, size);
    if (fd >= 0) close(fd);
  }
};

// One parsed node: blocks with their member lists.  Parsed per thread, merged serially.
struct ParsedBlock {
  uint32_t lb, ub;
  uint32_t tid_begin, tid_n;   // into NodeChunk::tids / strands
};
struct NodeChunk {
  std::vector<uint32_t> unitig;        // per node: global unitig id
  std::vector<uint32_t> nblocks;       // per node
  std::vector<ParsedBlock> blocks;
  std::vector<uint32_t> tids;
  std::vector<uint8_t> strands;
  std::vector<uint32_t> pos_n;         // per member (when positions)
  std::vector<uint32_t> pos_val;
};

struct NodeRef {
  const uint8_t* head;   // k ASCII chars
  const uint8_t* blob;
  uint32_t size;
};

}  // namespace

void decode_roaring_portable(const uint8_t* p, size_t n, std::vector<uint32_t>& out) {
  // CRoaring portable format, ext/bifrost/src/roaring.h:5720-5726, roaring.c:10405-10450.
  Cursor c{p, p + n};
  const uint32_t cookie = c.get<uint32_t>();
  uint32_t ncont;
  bool hasrun = false;
  const uint8_t* runbits = nullptr;
  if ((cookie & 0xFFFF) == 12347) {
    hasrun = true;
    ncont = (cookie >> 16) + 1;
    runbits = c.bytes((ncont + 7) / 8);
  } else if (cookie == 12346) {
    ncont = c.get<uint32_t>();
  } else {
    throw std::runtime_error("kallisto index: bad Roaring cookie");
  }
  if (ncont > 65536) throw std::runtime_error("kallisto index: bad Roaring container count");
  const uint8_t* keycard = c.bytes((size_t)ncont * 4);
  if (!hasrun || ncont >= 4) c.bytes((size_t)ncont * 4);   // offset header
  for (uint32_t i = 0; i < ncont; ++i) {
    uint16_t key, cm1;
    memcpy(&key, keycard + 4 * i, 2);
    memcpy(&cm1, keycard + 4 * i + 2, 2);
    const uint32_t card = (uint32_t)cm1 + 1;
    const uint32_t hi = (uint32_t)key << 16;
    const bool isrun = hasrun && ((runbits[i >> 3] >> (i & 7)) & 1);
    if (isrun) {
      const uint16_t nruns = c.get<uint16_t>();
      for (uint16_t r = 0; r < nruns; ++r) {
        const uint16_t start = c.get<uint16_t>();
        const uint16_t lenm1 = c.get<uint16_t>();
        for (uint32_t v = start; v <= (uint32_t)start + lenm1; ++v) out.push_back(hi | v);
      }
    } else if (card > 4096) {
      const uint8_t* bits = c.bytes(8192);
      for (uint32_t w = 0; w < 1024; ++w) {
        uint64_t x;
        memcpy(&x, bits + 8 * w, 8);
        while (x) {
          const int b = __builtin_ctzll(x);
          out.push_back(hi | (w * 64 + b));
          x &= x - 1;
        }
      }
    } else {
      const uint8_t* arr = c.bytes((size_t)card * 2);
      for (uint32_t j = 0; j < card; ++j) {
        uint16_t v;
        memcpy(&v, arr + 2 * j, 2);
        out.push_back(hi | v);
      }
    }
  }
}

void decode_roaring_native(const uint8_t* p, size_t n, std::vector<uint32_t>& out) {
  // roaring_bitmap_deserialize, ext/bifrost/src/roaring.c:8554-8568
  if (n == 0) throw std::runtime_error("kallisto index: empty Roaring blob");
  if (p[0] == 1) {            // CROARING_SERIALIZATION_ARRAY_UINT32
    if (n < 5) throw std::runtime_error("kallisto index: truncated Roaring blob");
    uint32_t card;
    memcpy(&card, p + 1, 4);
    if (n < 5 + (size_t)card * 4) throw std::runtime_error("kallisto index: truncated Roaring blob");
    const size_t base = out.size();
    out.resize(base + card);
    memcpy(out.data() + base, p + 5, (size_t)card * 4);
  } else if (p[0] == 2) {     // CROARING_SERIALIZATION_CONTAINER
    decode_roaring_portable(p + 1, n - 1, out);
  } else {
    throw std::runtime_error("kallisto index: unknown Roaring serialization tag");
  }
}

static void parse_nodes(const std::vector<NodeRef>& nodes, size_t begin, size_t end, int k,
                        const std::unordered_map<uint64_t, uint32_t>& head2unitig, bool want_pos,
                        NodeChunk& out) {
  std::vector<uint32_t> tmp;
  for (size_t ni = begin; ni < end; ++ni) {
    const NodeRef& nr = nodes[ni];
    uint64_t km = 0;
    for (int i = 0; i < k; ++i) km = (km << 2) | (uint64_t)base_code((char)nr.head[i]);
    const uint64_t rc = kmer_revcomp(km, k);
    auto it = head2unitig.find(km < rc ? km : rc);
    if (it == head2unitig.end())
      throw std::runtime_error("Error: Corrupted index; unitig not found: " + std::string((const char*)nr.head, k));
    Cursor c{nr.blob, nr.blob + nr.size};
    c.get<uint32_t>();  // Node::id (only used as a sort key by the reference)
    const uint8_t flag = c.get<uint8_t>();
    uint64_t nb = 0;
    if (flag == 1) nb = 1;
    else if (flag >= 2) nb = c.get<uint64_t>();
    out.unitig.push_back(it->second);
    out.nblocks.push_back((uint32_t)nb);
    for (uint64_t b = 0; b < nb; ++b) {
      ParsedBlock pb;
      pb.lb = c.get<uint32_t>();
      pb.ub = c.get<uint32_t>();
      const uint64_t rbytes = c.get<uint64_t>();
      const uint8_t* rblob = c.bytes(rbytes);
      pb.tid_begin = (uint32_t)out.tids.size();
      decode_roaring_native(rblob, rbytes, out.tids);
      pb.tid_n = (uint32_t)out.tids.size() - pb.tid_begin;
      const uint64_t vsz = c.get<uint64_t>();
      if (vsz != pb.tid_n) throw std::runtime_error("kallisto index: SparseVector size mismatch");
      for (uint64_t j = 0; j < vsz; ++j) {
        const uint64_t pbytes = c.get<uint64_t>();
        const uint8_t* pblob = c.bytes(pbytes);
        tmp.clear();
        decode_roaring_native(pblob, pbytes, tmp);
        if (tmp.empty()) throw std::runtime_error("kallisto index: empty position set");
        // sorted ascending: values with bit 31 (antisense) sort last
        const bool min_sense = (tmp.front() & 0x80000000u) == 0;
        const bool max_sense = (tmp.back() & 0x80000000u) == 0;
        out.strands.push_back(min_sense != max_sense ? 2 : (min_sense ? 1 : 0));
        if (want_pos) {
          out.pos_n.push_back((uint32_t)tmp.size());
          out.pos_val.insert(out.pos_val.end(), tmp.begin(), tmp.end());
        }
      }
      out.blocks.push_back(pb);
    }
  }
}

// 4-7: target lengths, names and the on-list -- the part of the file `kallisto bus` copies into index.saved
// (KmerIndex::write(fn, false), src/KmerIndex.cpp:1296-1324)
static void parse_targets(Cursor& c, FlatIndex& fi) {
  // 4-6. targets (KmerIndex.cpp:1470-1519)
  int32_t num_trans = c.get<int32_t>();
  if ((int64_t)num_trans < (int64_t)fi.dlist_n) throw std::runtime_error("kallisto index: bad target count");
  num_trans -= (int32_t)fi.dlist_n;
  fi.target_len.resize(num_trans);
  for (int32_t i = 0; i < num_trans; ++i) fi.target_len[i] = (uint32_t)c.get<int32_t>();
  {
    // every transcript id of every equivalence class must name a target (an index with a D-list also uses id
    // num_trans: the off-list pseudo-target of the dummy k-mer's unitig), and the lists must be strictly ascending
    const uint32_t limit = (uint32_t)num_trans + (fi.dlist_n ? 1u : 0u);
    for (uint32_t e = 0; e < fi.n_ec(); ++e)
      for (uint64_t i = fi.ec_off[e]; i < fi.ec_off[e + 1]; ++i) {
        if (fi.ec_tid[i] >= limit) throw std::runtime_error("kallisto index: equivalence class with a transcript id out of range");
        if (i > fi.ec_off[e] && fi.ec_tid[i] <= fi.ec_tid[i - 1]) throw std::runtime_error("kallisto index: unsorted equivalence class");
      }
  }
  fi.target_name.resize(num_trans);
  for (int32_t i = 0; i < num_trans; ++i) {
    const uint64_t n = c.get<uint64_t>();
    const uint8_t* s = c.bytes(n);
    // the reference builds the name with std::string(buffer): stops at the first NUL
    fi.target_name[i] = std::string((const char*)s, strnlen((const char*)s, n));
  }
  // 7. on-list, Roaring portable (KmerIndex.cpp:1522-1526)
  {
    const uint64_t n = c.get<uint64_t>();
    const uint8_t* s = c.bytes(n);
    decode_roaring_portable(s, n, fi.onlist);
  }
}

void load_index_v13(const std::string& path, FlatIndex& fi, bool load_positions, int threads) {
  Mmap mm(path);
  Cursor c{mm.data, mm.data + mm.size};
  fi = FlatIndex();

  // 1. version (KmerIndex.cpp:1351-1360)
  const uint64_t version = c.get<uint64_t>();
  if (version != 13) {
    throw std::runtime_error("Error: incompatible indices. Found version " + std::to_string(version) +
                             ", expected version 13\nRerun with index to regenerate");
  }
  // 2. Bifrost blob (KmerIndex.cpp:1362-1382): GRAPH section parsed, INDEX section skipped
  uint64_t dbg_bytes = c.get<uint64_t>();
  dbg_bytes &= (~0ULL >> 1);
  if (dbg_bytes == 0) {
    // index.saved of `kallisto bus` (KmerIndex::write(fn, false), src/KmerIndex.cpp:1226-1327): no graph, no MPHF field,
    // an empty D-list, no nodes -- targets only.  `quant-tcc` runs on it (the reference's loader skips the graph and the
    // MPHF together when the size is 0, :1365-1383); k keeps the reference's default.
    fi.graphless = true;
    fi.k = 31;
    fi.dlist_n = c.get<uint64_t>();
    c.get<uint64_t>();      // overhang
    const uint64_t n_nodes = c.get<uint64_t>();
    if (fi.dlist_n != 0 || n_nodes != 0) throw std::runtime_error("kallisto index: empty de Bruijn graph");
    fi.blk_off.assign(1, 0);
    fi.blk_strand_off.assign(1, 0);
    fi.ec_off.assign(1, 0);
    fi.useq_byteoff.assign(1, 0);
    parse_targets(c, fi);
    return;
  }
  {
    const uint8_t* gb = c.bytes(dbg_bytes);       // validates the length before the sub-cursor is formed
    Cursor g{gb, gb + dbg_bytes};
    const uint64_t fmt = g.get<uint64_t>();
    if ((fmt >> 32) != 0x7e215f3fULL) throw std::runtime_error("kallisto index: bad Bifrost graph header");
    fi.k = g.get<int32_t>();
    fi.g = g.get<int32_t>();
    if (fi.k < 3 || fi.k > 31) throw std::runtime_error("kallisto index: unsupported k");
    const uint64_t n_long = g.get<uint64_t>();
    fi.n_long = (uint32_t)n_long;
    fi.useq_byteoff.reserve(n_long + 1);
    fi.useq_byteoff.push_back(0);
    // first pass for total size
    {
      Cursor h = g;
      uint64_t total = 0;
      for (uint64_t i = 0; i < n_long; ++i) {
        const uint64_t len = h.get<uint64_t>();
        const uint64_t nb = (len + 3) / 4;
        h.bytes(nb);
        total += nb;
      }
      fi.useq.resize(total + 16);  // padding so that device-side 8-byte reads never run off the end
    }
    uint64_t off = 0;
    for (uint64_t i = 0; i < n_long; ++i) {
      const uint64_t len = g.get<uint64_t>();
      if (len < (uint64_t)fi.k) throw std::runtime_error("kallisto index: unitig shorter than k");
      const uint64_t nb = (len + 3) / 4;
      memcpy(fi.useq.data() + off, g.bytes(nb), nb);
      off += nb;
      fi.useq_byteoff.push_back(off);
      fi.ulen.push_back((uint32_t)len);
      fi.n_kmers += len - fi.k + 1;
    }
    const uint64_t n_short = g.get<uint64_t>();
    fi.n_short = (uint32_t)n_short;
    for (uint64_t i = 0; i < n_short; ++i) {
      const uint64_t w = g.get<uint64_t>();               // left-aligned (Kmer.cpp:92-107)
      fi.skmer.push_back(w >> (64 - 2 * fi.k));
      fi.ulen.push_back((uint32_t)fi.k);
    }
    const uint64_t n_abund = g.get<uint64_t>();
    fi.n_abund = (uint32_t)n_abund;
    for (uint64_t i = 0; i < n_abund; ++i) {
      const uint64_t w = g.get<uint64_t>();
      fi.skmer.push_back(w >> (64 - 2 * fi.k));
      fi.ulen.push_back((uint32_t)fi.k);
    }
    fi.n_kmers += n_short + n_abund;
  }
  // MPHF blob: skipped
  {
    const uint64_t mphf_bytes = c.get<uint64_t>();
    c.bytes(mphf_bytes);
  }
  // 2.2 D-list (KmerIndex.cpp:1385-1403)
  fi.dlist_n = c.get<uint64_t>();
  c.get<uint64_t>();  // overhang
  {
    const uint8_t* dl = c.bytes(fi.dlist_n * 8);
    fi.dlist.resize(fi.dlist_n);
    for (uint64_t i = 0; i < fi.dlist_n; ++i) {
      uint64_t w;
      memcpy(&w, dl + i * 8, 8);
      fi.dlist[i] = w >> (64 - 2 * fi.k);     // Kmer: left-aligned 2-bit words, already canonical (rep(), KmerIndex.cpp:946)
    }
  }

  const int k = fi.k;
  const uint32_t nU = fi.n_unitigs();

  // head k-mer (canonical) -> unitig
  std::unordered_map<uint64_t, uint32_t> head2unitig;
  head2unitig.reserve((size_t)nU * 2);
  for (uint32_t u = 0; u < fi.n_long; ++u) {
    const uint8_t* s = fi.useq.data() + fi.useq_byteoff[u];
    uint64_t km = 0;
    for (int i = 0; i < k; ++i) km = (km << 2) | ((s[i >> 2] >> ((i & 3) * 2)) & 3);
    const uint64_t rc = kmer_revcomp(km, k);
    head2unitig[km < rc ? km : rc] = u;
  }
  for (uint32_t j = 0; j < fi.n_short + fi.n_abund; ++j) {
    const uint64_t km = fi.skmer[j];
    const uint64_t rc = kmer_revcomp(km, k);
    head2unitig[km < rc ? km : rc] = fi.n_long + j;
  }

  // 3. nodes (KmerIndex.cpp:1405-1468)
  const uint64_t n_nodes = c.get<uint64_t>();
  std::vector<NodeRef> nodes;
  nodes.reserve(n_nodes);
  for (uint64_t i = 0; i < n_nodes; ++i) {
    NodeRef nr;
    nr.head = c.bytes(k);
    nr.size = c.get<uint32_t>();
    nr.blob = c.bytes(nr.size);
    nodes.push_back(nr);
  }
  int nt = std::max(1, threads);
  if ((uint64_t)nt > n_nodes / 1024 + 1) nt = (int)(n_nodes / 1024 + 1);
  std::vector<NodeChunk> chunks(nt);
  {
    std::vector<std::thread> pool;
    std::vector<std::string> errs(nt);
    for (int t = 0; t < nt; ++t) {
      const size_t b = n_nodes * t / nt, e = n_nodes * (t + 1) / nt;
      pool.emplace_back([&, t, b, e] {
        try {
          parse_nodes(nodes, b, e, k, head2unitig, load_positions, chunks[t]);
        } catch (const std::exception& ex) {
          errs[t] = ex.what();
        }
      });
    }
    for (auto& th : pool) th.join();
    for (auto& e : errs)
      if (!e.empty()) throw std::runtime_error(e);
  }

  // Merge: group blocks by unitig (nodes may come in any order), de-duplicate EC sets by content.
  std::vector<uint32_t> ublocks(nU, 0);
  uint64_t total_blocks = 0, total_members = 0;
  for (auto& ch : chunks) {
    for (size_t i = 0; i < ch.unitig.size(); ++i) {
      if (ublocks[ch.unitig[i]] != 0) throw std::runtime_error("kallisto index: duplicate node for a unitig");
      ublocks[ch.unitig[i]] = ch.nblocks[i];
    }
    total_blocks += ch.blocks.size();
    total_members += ch.tids.size();
  }
  for (uint32_t u = 0; u < nU; ++u)
    if (ublocks[u] == 0) throw std::runtime_error("kallisto index: unitig without equivalence-class blocks");
  fi.blk_off.assign(nU + 1, 0);
  for (uint32_t u = 0; u < nU; ++u) fi.blk_off[u + 1] = fi.blk_off[u] + ublocks[u];
  fi.blk_lb.resize(total_blocks);
  fi.blk_ub.resize(total_blocks);
  fi.blk_ec.resize(total_blocks);
  fi.blk_strand_off.assign(total_blocks + 1, 0);

  // content-dedup table: open addressing over EC ids
  size_t cap = 64;
  while (cap < total_blocks * 2 + 16) cap <<= 1;
  std::vector<uint32_t> tab(cap, UINT32_MAX);
  fi.ec_off.push_back(0);
  auto hash_set = [](const uint32_t* v, uint32_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ULL ^ n;
    for (uint32_t i = 0; i < n; ++i) {
      h ^= v[i] + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
      h *= 0xFF51AFD7ED558CCDULL;
      h ^= h >> 32;
    }
    return h;
  };
  auto intern = [&](const uint32_t* v, uint32_t n) -> uint32_t {
    size_t h = hash_set(v, n) & (cap - 1);
    for (;;) {
      const uint32_t e = tab[h];
      if (e == UINT32_MAX) {
        const uint32_t id = (uint32_t)(fi.ec_off.size() - 1);
        fi.ec_tid.insert(fi.ec_tid.end(), v, v + n);
        fi.ec_off.push_back(fi.ec_tid.size());
        tab[h] = id;
        return id;
      }
      const uint64_t b = fi.ec_off[e];
      if (fi.ec_off[e + 1] - b == n && (n == 0 || memcmp(fi.ec_tid.data() + b, v, (size_t)n * 4) == 0)) return e;
      h = (h + 1) & (cap - 1);
    }
  };

  // First pass: per-block sizes at their final slots (so strand/pos CSR can be laid out in unitig order).
  {
    for (auto& ch : chunks) {
      size_t bi = 0;
      for (size_t i = 0; i < ch.unitig.size(); ++i) {
        const uint64_t base = fi.blk_off[ch.unitig[i]];
        for (uint32_t j = 0; j < ch.nblocks[i]; ++j, ++bi) fi.blk_strand_off[base + j + 1] = ch.blocks[bi].tid_n;
      }
    }
    for (uint64_t b = 0; b < total_blocks; ++b) fi.blk_strand_off[b + 1] += fi.blk_strand_off[b];
  }
  fi.strand.resize(total_members);
  fi.has_positions = load_positions;
  std::vector<uint32_t> pos_cnt;
  if (load_positions) pos_cnt.assign(total_members + 1, 0);
  for (auto& ch : chunks) {
    size_t bi = 0;
    for (size_t i = 0; i < ch.unitig.size(); ++i) {
      const uint32_t u = ch.unitig[i];
      const uint64_t base = fi.blk_off[u];
      const uint32_t nk = fi.ulen[u] - k + 1;
      uint32_t prev_ub = 0;
      for (uint32_t j = 0; j < ch.nblocks[i]; ++j, ++bi) {
        const ParsedBlock& pb = ch.blocks[bi];
        // The reference's BlockArray lookups (get_block_at / operator[], BlockArray.hpp:257-322)
        // are only well defined when the blocks tile [0, #kmers) -- which is what the index
        // builder writes.  Anything else is rejected loudly rather than guessed at.
        if (pb.lb >= pb.ub || pb.lb != prev_ub || (j + 1 == ch.nblocks[i] && pb.ub != nk))
          throw std::runtime_error("kallisto index: EC blocks do not tile the unitig (unsupported index)");
        prev_ub = pb.ub;
        fi.blk_lb[base + j] = pb.lb;
        fi.blk_ub[base + j] = pb.ub;
        fi.blk_ec[base + j] = intern(ch.tids.data() + pb.tid_begin, pb.tid_n);
        const uint64_t so = fi.blk_strand_off[base + j];
        memcpy(fi.strand.data() + so, ch.strands.data() + pb.tid_begin, pb.tid_n);
        if (load_positions)
          for (uint32_t m = 0; m < pb.tid_n; ++m) pos_cnt[so + m + 1] = ch.pos_n[pb.tid_begin + m];
      }
    }
  }
  if (load_positions) {
    fi.pos_off.assign(total_members + 1, 0);
    for (uint64_t i = 0; i < total_members; ++i) fi.pos_off[i + 1] = fi.pos_off[i] + pos_cnt[i + 1];
    fi.pos_val.resize(fi.pos_off[total_members]);
    for (auto& ch : chunks) {
      size_t bi = 0;
      size_t src = 0;  // running offset in ch.pos_val (members are stored in parse order)
      for (size_t i = 0; i < ch.unitig.size(); ++i) {
        const uint64_t base = fi.blk_off[ch.unitig[i]];
        for (uint32_t j = 0; j < ch.nblocks[i]; ++j, ++bi) {
          const ParsedBlock& pb = ch.blocks[bi];
          const uint64_t so = fi.blk_strand_off[base + j];
          for (uint32_t m = 0; m < pb.tid_n; ++m) {
            const uint32_t n = ch.pos_n[pb.tid_begin + m];
            memcpy(fi.pos_val.data() + fi.pos_off[so + m], ch.pos_val.data() + src, (size_t)n * 4);
            src += n;
          }
        }
      }
    }
  }
  chunks.clear();
  if (load_positions) {
    // findPosition constants.  `holds(b, tr)` = tr is a member of block b's EC.
    fi.fp_info.assign((size_t)total_members * 4, 0);
    for (uint32_t u = 0; u < nU; ++u) {
      const uint64_t b0 = fi.blk_off[u], b1 = fi.blk_off[u + 1];
      auto holds = [&](uint64_t b, uint32_t tr) {
        const uint64_t e = fi.blk_ec[b];
        const uint32_t* s = fi.ec_tid.data() + fi.ec_off[e];
        const uint32_t* t = s + (fi.ec_off[e + 1] - fi.ec_off[e]);
        return std::binary_search(s, t, tr);
      };
      for (uint64_t b = b0; b < b1; ++b) {
        const uint64_t e = fi.blk_ec[b];
        const uint64_t n = fi.ec_off[e + 1] - fi.ec_off[e];
        for (uint64_t m = 0; m < n; ++m) {
          const uint32_t tr = fi.ec_tid[fi.ec_off[e] + m];
          const uint64_t slot = fi.blk_strand_off[b] + m;
          uint32_t* out = fi.fp_info.data() + slot * 4;
          const uint32_t rawmin = fi.pos_val[fi.pos_off[slot]];   // sorted ascending: the minimum
          out[0] = rawmin;
          // case I: only when trpos == 0
          uint32_t pad = 0;
          if ((rawmin & 0x7FFFFFFFu) == 0) {
            uint64_t cur = b;
            for (uint64_t i = b; i-- > b0;) {
              if (!holds(i, tr)) { pad = fi.blk_lb[cur]; break; }
              cur = i;
            }
          }
          out[1] = pad;
          // case III
          uint32_t left3 = 0;
          for (uint64_t i = b; i-- > b0;) {
            if (!holds(i, tr)) { left3 = fi.blk_ub[i]; break; }
          }
          out[2] = left3;
          // cases II / IV: over all blocks of the unitig
          uint32_t left = 0, right = 0, unmapped = 0;
          bool found = false;
          for (uint64_t i = b0; i < b1; ++i) {
            const bool hs = holds(i, tr);
            if (!hs && found) {
              if (unmapped == 0) left = fi.blk_lb[i];
              right = fi.blk_ub[i];
              unmapped += fi.blk_ub[i] - fi.blk_lb[i];
            }
            if (hs) found = true;
          }
          out[3] = right - left;
        }
      }
    }
  }

  parse_targets(c, fi);
}

}  // namespace kb
