// abundance.h5 without libhdf5: a minimal emitter of the HDF5 file format (superblock version 0, "old style" groups,
// version-1 object headers) for exactly what H5Writer writes (src/H5Writer.cpp:4-71, src/h5utils.h:42-91):
//   /est_counts, /aux/{num_bootstrap, num_processed, fld, bias_observed, bias_normalized, kallisto_version,
//   index_version, call, start_time, ids, eff_lengths, lengths}, /bootstrap/bs<i>
// every one a 1-D dataset of native int (32-bit LE), IEEE double (LE) or fixed-length NUL-terminated strings
// (H5T_C_S1 of the longest string + 1), stored as ONE chunk with the deflate filter at level 6 (H5Pset_chunk(dims) +
// H5Pset_deflate, h5utils.h:54-60).  Host-only; zlib does the compression.
//
// File image (all addresses absolute, 8-byte sizes):
//   superblock v0 (96 B)  root symbol-table entry -> root object header
//   per group   object header {symbol table message} -> B-tree v1 node (type 0) -> symbol table nodes (SNOD) of <= 2*4
//               entries sorted by name; names in a local heap
//   per dataset object header {dataspace v1, datatype v1, fill value v2, filter pipeline v1 (deflate), layout v3 chunked}
//               -> B-tree v1 node (type 1, one key pair) -> the compressed chunk
#pragma once
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace kb {

class H5Writer {
 public:
  H5Writer() { groups_.push_back(Group{"/", {}, {}}); }

  // a group directly under the root; returns its handle (0 = the root itself)
  int group(const std::string& name) {
    groups_.push_back(Group{name, {}, {}});
    groups_[0].sub.push_back((int)groups_.size() - 1);
    return (int)groups_.size() - 1;
  }
  void add_i32(int g, const std::string& name, const int32_t* v, size_t n) { add_raw(g, name, I32, n, 4, v); }
  void add_f64(int g, const std::string& name, const double* v, size_t n) { add_raw(g, name, F64, n, 8, v); }
  void add_str(int g, const std::string& name, const std::vector<std::string>& v) {
    size_t w = 0;
    for (auto& s : v) w = std::max(w, s.size());
    w += 1;                                              // get_datatype_id: longest string + terminator (h5utils.cpp:34-52)
    std::vector<char> pool(w * v.size(), 0);
    for (size_t i = 0; i < v.size(); ++i) memcpy(pool.data() + i * w, v[i].data(), v[i].size());
    add_raw(g, name, STR, v.size(), (uint32_t)w, pool.data());
  }

  // Lays the file out and writes it.  false on an I/O or compression error.
  bool write(const std::string& path, int level = 6) {
    f_.clear();
    alloc(96);                                           // superblock, filled in last
    // every dataset is one independent zlib stream: compress them on all cores first (100 bootstrap vectors of a human
    // transcriptome are 13 s of deflate on one core)
    {
      std::vector<Dataset*> all;
      for (auto& g : groups_) for (auto& d : g.ds) all.push_back(&d);
      std::atomic<size_t> next{0};
      std::atomic<bool> ok{true};
      auto work = [&] {
        for (size_t i; (i = next.fetch_add(1)) < all.size();) {
          Dataset& d = *all[i];
          uLongf cl = compressBound((uLong)d.raw.size());
          d.z.resize(cl);
          if (d.raw.size() >= (1ull << 32) || compress2(d.z.data(), &cl, d.raw.data(), (uLong)d.raw.size(), level) != Z_OK) { ok = false; continue; }
          d.z.resize(cl);
        }
      };
      const unsigned nt = std::max(1u, std::min<unsigned>({std::thread::hardware_concurrency(), 32u, (unsigned)all.size()}));
      std::vector<std::thread> pool;
      for (unsigned t = 1; t < nt; ++t) pool.emplace_back(work);
      work();
      for (auto& th : pool) th.join();
      if (!ok) return false;
    }
    size_t max_entries = 0;
    for (auto& g : groups_) max_entries = std::max(max_entries, g.ds.size() + g.sub.size());
    const size_t max_nodes = (max_entries + 2 * kLeafK - 1) / (2 * kLeafK);
    ik_ = (uint16_t)std::max<size_t>(16, (max_nodes + 1) / 2);      // children per B-tree node <= 2 * ik
    std::vector<GroupAddr> ga(groups_.size());
    for (size_t g = groups_.size(); g-- > 0;) {          // children before the root
      std::vector<Entry> ent;
      for (auto& d : groups_[g].ds) {
        const uint64_t a = write_dataset(d, level);
        if (!a) return false;
        ent.push_back(Entry{d.name, a, false, 0, 0});
      }
      for (int s : groups_[g].sub) ent.push_back(Entry{groups_[s].name, ga[s].header, true, ga[s].btree, ga[s].heap});
      ga[g] = write_group(ent);
    }
    // superblock, version 0
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    memcpy(&f_[0], sig, 8);
    size_t o = 8;
    const uint8_t vers[8] = {0, 0, 0, 0, 0, 8, 8, 0};    // superblock, free space, root entry, -, shared header, offsets, lengths, -
    memcpy(&f_[o], vers, 8); o += 8;
    put16(o, kLeafK); put16(o + 2, ik_); put32(o + 4, 0); o += 8;   // group leaf K, group internal K, consistency flags
    put64(o, 0); put64(o + 8, kUndef); put64(o + 16, f_.size()); put64(o + 24, kUndef); o += 32;   // base, free space, EOF, driver
    put64(o, 0); put64(o + 8, ga[0].header); put32(o + 16, 1); put32(o + 20, 0);                  // root symbol table entry
    put64(o + 24, ga[0].btree); put64(o + 32, ga[0].heap);
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) return false;
    const bool ok = fwrite(f_.data(), 1, f_.size(), fp) == f_.size();
    return fclose(fp) == 0 && ok;
  }

 private:
  enum Kind { I32, F64, STR };
  struct Dataset { std::string name; Kind kind; uint64_t n; uint32_t elem; std::vector<uint8_t> raw, z; };
  struct Group { std::string name; std::vector<Dataset> ds; std::vector<int> sub; };
  struct Entry { std::string name; uint64_t header; bool is_group; uint64_t btree, heap; };
  struct GroupAddr { uint64_t header = 0, btree = 0, heap = 0; };
  static constexpr uint64_t kUndef = ~0ull;
  static constexpr uint16_t kLeafK = 4;                  // symbol table nodes hold <= 2 * 4 entries (the library's default)
  static constexpr int kChunkK = 32;                     // chunk B-tree nodes: the library's default for superblock v0

  void add_raw(int g, const std::string& name, Kind k, size_t n, uint32_t elem, const void* src) {
    Dataset d{name, k, n, elem, {}, {}};
    d.raw.resize(n * elem);
    if (n) memcpy(d.raw.data(), src, n * elem);
    groups_[g].ds.push_back(std::move(d));
  }
  size_t alloc(size_t n) {
    const size_t off = (f_.size() + 7) & ~(size_t)7;
    f_.resize(off + n, 0);
    return off;
  }
  void put16(size_t o, uint16_t v) { memcpy(&f_[o], &v, 2); }
  void put32(size_t o, uint32_t v) { memcpy(&f_[o], &v, 4); }
  void put64(size_t o, uint64_t v) { memcpy(&f_[o], &v, 8); }

  // object header message: type, size of the (padded) data, flags, 3 reserved bytes, data padded to 8 bytes
  static void message(std::vector<uint8_t>& m, uint16_t type, const std::vector<uint8_t>& data) {
    const uint16_t sz = (uint16_t)((data.size() + 7) & ~(size_t)7);
    const size_t o = m.size();
    m.resize(o + 8 + sz, 0);
    memcpy(&m[o], &type, 2);
    memcpy(&m[o + 2], &sz, 2);
    memcpy(&m[o + 8], data.data(), data.size());
  }
  template <class T> static void app(std::vector<uint8_t>& v, T x) {
    const size_t o = v.size();
    v.resize(o + sizeof(T));
    memcpy(&v[o], &x, sizeof(T));
  }
  uint64_t write_header(const std::vector<uint8_t>& msgs, uint16_t n_msgs) {
    const size_t h = alloc(16 + msgs.size());            // version-1 prefix: 12 bytes + 4 of alignment
    f_[h] = 1;
    put16(h + 2, n_msgs);
    put32(h + 4, 1);                                     // object reference count
    put32(h + 8, (uint32_t)msgs.size());                 // size of the message area
    memcpy(&f_[h + 16], msgs.data(), msgs.size());
    return h;
  }

  uint64_t write_dataset(const Dataset& d, int level) {
    if (d.n == 0 || d.n >= (1ull << 32) || d.raw.size() >= (1ull << 32) || d.z.empty()) return 0;    // one chunk: 32-bit chunk dimensions and size
    const size_t cl = d.z.size();                        // compressed by write() already
    const size_t chunk = alloc(cl);
    memcpy(&f_[chunk], d.z.data(), cl);
    // chunk index: B-tree v1, node type 1, one leaf with one chunk.  Key = {chunk bytes, filter mask, offsets[rank + 1]}
    const size_t key = 8 + 2 * 8;
    const size_t bt = alloc(24 + (2 * kChunkK + 1) * key + 2 * kChunkK * 8);
    memcpy(&f_[bt], "TREE", 4);
    f_[bt + 4] = 1;                                      // node type: raw data chunks
    f_[bt + 5] = 0;                                      // level
    put16(bt + 6, 1);                                    // entries used
    put64(bt + 8, kUndef); put64(bt + 16, kUndef);       // siblings
    size_t o = bt + 24;
    put32(o, (uint32_t)cl); put32(o + 4, 0); put64(o + 8, 0); put64(o + 16, 0); o += key;
    put64(o, chunk); o += 8;
    put32(o, 0); put32(o + 4, 0); put64(o + 8, d.n); put64(o + 16, 0);             // closing key: one chunk past the end
    std::vector<uint8_t> m, b;
    // dataspace, version 1: rank 1, maximum = current size (H5Screate_simple(1, dims, NULL))
    b = {1, 1, 1, 0, 0, 0, 0, 0};
    app<uint64_t>(b, d.n); app<uint64_t>(b, d.n);
    message(m, 0x0001, b);
    // datatype, version 1
    if (d.kind == I32) {          // class 0 fixed point: little-endian, signed; 4 bytes; bit offset 0, precision 32
      b = {0x10, 0x08, 0, 0, 4, 0, 0, 0, 0, 0, 32, 0};
    } else if (d.kind == F64) {   // class 1 floating point: little-endian, implied mantissa msb, sign bit 63; 8 bytes; bit offset 0,
                                  // precision 64, exponent at 52 (11 bits), mantissa at 0 (52 bits), bias 1023
      b = {0x11, 0x20, 0x3f, 0, 8, 0, 0, 0, 0, 0, 64, 0, 52, 11, 0, 52, 0xff, 0x03, 0, 0};
    } else {                      // class 3 string: NUL-terminated, ASCII; size = element width
      b = {0x13, 0, 0, 0};
      app<uint32_t>(b, d.elem);
    }
    message(m, 0x0003, b);
    // fill value, version 2: allocate incrementally (chunked), write the fill value if set, default value (size 0)
    b = {2, 3, 2, 1, 0, 0, 0, 0};
    message(m, 0x0005, b);
    // filter pipeline, version 1: one filter, deflate (id 1), optional, one client value = the level
    b = {1, 1, 0, 0, 0, 0, 0, 0};
    app<uint16_t>(b, 1); app<uint16_t>(b, 8); app<uint16_t>(b, 1); app<uint16_t>(b, 1);
    for (char c : std::string("deflate")) b.push_back((uint8_t)c);
    b.push_back(0);
    app<uint32_t>(b, (uint32_t)level); app<uint32_t>(b, 0);
    message(m, 0x000B, b);
    // data layout, version 3, chunked: rank + 1 dimensions (the last is the element size), address of the chunk B-tree
    b = {3, 2, 2};
    app<uint64_t>(b, bt); app<uint32_t>(b, (uint32_t)d.n); app<uint32_t>(b, d.elem);
    message(m, 0x0008, b);
    return write_header(m, 5);
  }

  GroupAddr write_group(std::vector<Entry>& ent) {
    std::sort(ent.begin(), ent.end(), [](const Entry& a, const Entry& b) { return a.name < b.name; });   // bytewise = strcmp
    std::vector<uint8_t> hd(8, 0);                       // local heap data: the empty name at offset 0
    std::vector<uint64_t> noff(ent.size());
    for (size_t i = 0; i < ent.size(); ++i) {
      noff[i] = hd.size();
      hd.insert(hd.end(), ent[i].name.begin(), ent[i].name.end());
      hd.push_back(0);
      while (hd.size() & 7) hd.push_back(0);
    }
    GroupAddr a;
    const size_t heap = alloc(32 + hd.size());
    memcpy(&f_[heap], "HEAP", 4);
    put64(heap + 8, hd.size());                          // data segment size
    put64(heap + 16, 1);                                 // head of the free list: 1 = none (H5HL_FREE_NULL)
    put64(heap + 24, heap + 32);                         // data segment address
    memcpy(&f_[heap + 32], hd.data(), hd.size());
    a.heap = heap;
    const size_t per = 2 * kLeafK;
    const size_t n_nodes = (ent.size() + per - 1) / per;
    std::vector<uint64_t> node(n_nodes);
    for (size_t s = 0; s < n_nodes; ++s) {
      const size_t sn = alloc(8 + per * 40);
      const size_t lo = s * per, hi = std::min(ent.size(), lo + per);
      memcpy(&f_[sn], "SNOD", 4);
      f_[sn + 4] = 1;
      put16(sn + 6, (uint16_t)(hi - lo));
      for (size_t i = lo; i < hi; ++i) {
        const size_t e = sn + 8 + (i - lo) * 40;
        put64(e, noff[i]);
        put64(e + 8, ent[i].header);
        put32(e + 16, ent[i].is_group ? 1u : 0u);        // cache type 1: the scratch pad holds the group's B-tree and heap
        if (ent[i].is_group) { put64(e + 24, ent[i].btree); put64(e + 32, ent[i].heap); }
      }
      node[s] = sn;
    }
    const size_t bt = alloc(24 + (2 * (size_t)ik_ + 1) * 8 + 2 * (size_t)ik_ * 8);
    memcpy(&f_[bt], "TREE", 4);
    f_[bt + 4] = 0;                                      // node type: group nodes
    f_[bt + 5] = 0;
    put16(bt + 6, (uint16_t)n_nodes);
    put64(bt + 8, kUndef); put64(bt + 16, kUndef);
    size_t o = bt + 24;
    put64(o, 0); o += 8;                                 // key 0: the empty name
    for (size_t s = 0; s < n_nodes; ++s) {
      put64(o, node[s]); o += 8;
      put64(o, noff[std::min(ent.size(), (s + 1) * per) - 1]); o += 8;     // key s + 1: the largest name in child s
    }
    a.btree = bt;
    std::vector<uint8_t> m, b;
    app<uint64_t>(b, bt); app<uint64_t>(b, heap);
    message(m, 0x0011, b);                               // symbol table message
    a.header = write_header(m, 1);
    return a;
  }

  std::vector<Group> groups_;
  std::vector<uint8_t> f_;
  uint16_t ik_ = 16;
};

}  // namespace kb
