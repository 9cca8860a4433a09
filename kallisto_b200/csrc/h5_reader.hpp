// Reading side of abundance.h5 for `kallisto_b200 h5dump` (H5Converter, src/H5Writer.cpp:75-200) without libhdf5: walks the
// HDF5 file format structures kallisto's files use -- superblock version 0/1, symbol-table groups (B-tree v1 of any depth,
// symbol table nodes, local heap), version-1 object headers with continuation blocks, 1-D datasets with compact,
// contiguous or chunked (B-tree v1 index, deflate filter) layout, fixed-point / IEEE float / fixed-length string types.
// That covers the files csrc/h5_writer.hpp writes and what libhdf5 writes for H5Writer's calls with default property
// lists; anything else (superblock 2/3, new-style groups, other filters) is reported as an error, never guessed.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace kb {

class H5Reader {
 public:
  explicit H5Reader(const std::string& path) {
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) throw std::runtime_error("Error: could not open " + path);
    fseek(fp, 0, SEEK_END);
    const long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    b_.resize(n > 0 ? (size_t)n : 0);
    const bool ok = b_.empty() || fread(b_.data(), 1, b_.size(), fp) == b_.size();
    fclose(fp);
    if (!ok) throw std::runtime_error("Error: could not read " + path);
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (b_.size() < 96 || memcmp(b_.data(), sig, 8) != 0) throw std::runtime_error("Error: " + path + " is not an HDF5 file");
    const int sbv = b_[8];
    if (sbv > 1) throw std::runtime_error("Error: HDF5 superblock version " + std::to_string(sbv) + " is not supported");
    if (b_[13] != 8 || b_[14] != 8) throw std::runtime_error("Error: HDF5 file with offsets / lengths that are not 8 bytes");
    size_t o = 24 + (sbv == 1 ? 4 : 0);       // version 1 adds the indexed-storage K and two reserved bytes
    base_ = u64(o);
    o += 32;                                   // base, free space, end of file, driver information
    root_header_ = u64(o + 8);                 // root symbol table entry: link name offset, object header address, ...
  }

  // path = "/name" or "/group/name"
  bool exists(const std::string& path) { return find(path) != 0; }
  std::vector<double> read_f64(const std::string& path) {
    Data d = read(path);
    std::vector<double> out(d.n);
    for (uint64_t i = 0; i < d.n; ++i) {
      const uint8_t* p = d.raw.data() + i * d.elem;
      if (d.cls == 1 && d.elem == 8) { double v; memcpy(&v, p, 8); out[i] = v; }
      else if (d.cls == 1 && d.elem == 4) { float v; memcpy(&v, p, 4); out[i] = v; }
      else if (d.cls == 0) out[i] = (double)as_int(p, d);
      else throw std::runtime_error("Error: dataset " + path + " is not numeric");
    }
    return out;
  }
  std::vector<int64_t> read_int(const std::string& path) {
    Data d = read(path);
    if (d.cls != 0) throw std::runtime_error("Error: dataset " + path + " is not an integer dataset");
    std::vector<int64_t> out(d.n);
    for (uint64_t i = 0; i < d.n; ++i) out[i] = as_int(d.raw.data() + i * d.elem, d);
    return out;
  }
  std::vector<std::string> read_str(const std::string& path) {
    Data d = read(path);
    if (d.cls != 3) throw std::runtime_error("Error: dataset " + path + " is not a string dataset");
    std::vector<std::string> out(d.n);
    for (uint64_t i = 0; i < d.n; ++i) {
      const char* p = (const char*)d.raw.data() + i * d.elem;
      out[i] = std::string(p, strnlen(p, d.elem));
    }
    return out;
  }

 private:
  struct Msg { uint16_t type; const uint8_t* p; uint16_t size; };
  struct Data { int cls = -1; bool is_signed = true; uint32_t elem = 0; uint64_t n = 0; std::vector<uint8_t> raw; };
  static constexpr uint64_t kUndef = ~0ull;

  const uint8_t* at(uint64_t addr, uint64_t n) const {
    addr += base_;
    if (addr == kUndef || addr > b_.size() || n > b_.size() - addr) throw std::runtime_error("Error: HDF5 file is truncated or damaged");
    return b_.data() + addr;
  }
  uint16_t u16(uint64_t o) const { uint16_t v; memcpy(&v, b_.data() + o, 2); return v; }
  uint64_t u64(uint64_t o) const { uint64_t v; memcpy(&v, b_.data() + o, 8); return v; }
  static uint16_t g16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
  static uint32_t g32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
  static uint64_t g64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
  static int64_t as_int(const uint8_t* p, const Data& d) {
    uint64_t v = 0;
    memcpy(&v, p, d.elem);
    if (d.is_signed && d.elem < 8 && (v >> (8 * d.elem - 1)) & 1) v |= ~0ull << (8 * d.elem);
    return (int64_t)v;
  }

  // all messages of a version-1 object header, continuation blocks included
  std::vector<Msg> messages(uint64_t hdr) const {
    const uint8_t* h = at(hdr, 16);
    if (h[0] != 1) throw std::runtime_error("Error: HDF5 object header version " + std::to_string(h[0]) + " is not supported");
    std::vector<Msg> out;
    std::vector<std::pair<uint64_t, uint64_t>> blocks{{hdr + 16, g32(h + 8)}};
    for (size_t bi = 0; bi < blocks.size() && bi < 64; ++bi) {
      const uint8_t* p = at(blocks[bi].first, blocks[bi].second);
      uint64_t o = 0;
      while (o + 8 <= blocks[bi].second) {
        const uint16_t type = g16(p + o), size = g16(p + o + 2);
        if (o + 8 + size > blocks[bi].second) break;
        if (type == 0x0010 && size >= 16) blocks.push_back({g64(p + o + 8), g64(p + o + 16)});     // object header continuation
        else if (type != 0) out.push_back(Msg{type, p + o + 8, size});
        o += 8 + (uint64_t)size;
      }
    }
    return out;
  }

  // name -> object header address of a group's members (symbol table: B-tree v1 of group nodes over symbol table nodes)
  std::map<std::string, uint64_t> members(uint64_t hdr) const {
    uint64_t bt = kUndef, heap = kUndef;
    for (auto& m : messages(hdr))
      if (m.type == 0x0011 && m.size >= 16) { bt = g64(m.p); heap = g64(m.p + 8); }
    if (bt == kUndef) throw std::runtime_error("Error: HDF5 group without a symbol table (new-style groups are not supported)");
    const uint8_t* hh = at(heap, 32);
    if (memcmp(hh, "HEAP", 4) != 0) throw std::runtime_error("Error: HDF5 local heap expected");
    const uint64_t hsize = g64(hh + 8), hdata = g64(hh + 24);
    const char* names = (const char*)at(hdata, hsize);
    std::map<std::string, uint64_t> out;
    walk_group_tree(bt, names, hsize, out, 0);
    return out;
  }
  void walk_group_tree(uint64_t addr, const char* names, uint64_t hsize, std::map<std::string, uint64_t>& out, int depth) const {
    if (depth > 16 || ++visits_ > 4000000) throw std::runtime_error("Error: HDF5 group B-tree too deep or cyclic");
    const uint8_t* n = at(addr, 24);
    if (memcmp(n, "TREE", 4) == 0) {
      if (n[4] != 0) throw std::runtime_error("Error: HDF5 group B-tree node expected");
      const unsigned used = g16(n + 6);
      const uint8_t* body = at(addr + 24, (uint64_t)used * 16 + 8);
      for (unsigned i = 0; i < used; ++i) walk_group_tree(g64(body + 8 + 16 * (uint64_t)i), names, hsize, out, depth + 1);
    } else if (memcmp(n, "SNOD", 4) == 0) {
      const unsigned used = g16(n + 6);
      const uint8_t* e = at(addr + 8, (uint64_t)used * 40);
      for (unsigned i = 0; i < used; ++i) {
        const uint64_t no = g64(e + 40 * (uint64_t)i);
        if (no >= hsize) throw std::runtime_error("Error: HDF5 link name outside its heap");
        out[std::string(names + no, strnlen(names + no, hsize - no))] = g64(e + 40 * (uint64_t)i + 8);
      }
    } else {
      throw std::runtime_error("Error: HDF5 group node expected");
    }
  }
  uint64_t find(const std::string& path) {
    uint64_t cur = root_header_;
    size_t o = 0;
    while (o < path.size()) {
      while (o < path.size() && path[o] == '/') ++o;
      size_t e = path.find('/', o);
      if (e == std::string::npos) e = path.size();
      if (e == o) break;
      auto it = cache_.find(cur);
      if (it == cache_.end()) it = cache_.emplace(cur, members(cur)).first;
      auto f = it->second.find(path.substr(o, e - o));
      if (f == it->second.end()) return 0;
      cur = f->second;
      o = e;
    }
    return cur;
  }

  Data read(const std::string& path) {
    const uint64_t hdr = find(path);
    if (!hdr) throw std::runtime_error("Error: no dataset " + path + " in the HDF5 file");
    Data d;
    const uint8_t* layout = nullptr;
    uint16_t layout_size = 0;
    bool deflate = false;
    for (auto& m : messages(hdr)) {
      if (m.type == 0x0001 && m.size >= 8) {              // dataspace, version 1 or 2
        const int ver = m.p[0], rank = m.p[1];
        const size_t dims = ver == 1 ? 8 : 4;
        if (rank == 0) d.n = 1;
        else if (rank == 1 && m.size >= dims + 8) d.n = g64(m.p + dims);
        else throw std::runtime_error("Error: dataset " + path + " is not one-dimensional");
      } else if (m.type == 0x0003 && m.size >= 8) {       // datatype
        d.cls = m.p[0] & 15;
        d.elem = g32(m.p + 4);
        if ((d.cls == 0 || d.cls == 1) && (m.p[1] & 1)) throw std::runtime_error("Error: big-endian dataset " + path);
        d.is_signed = d.cls == 0 && (m.p[1] & 8);
      } else if (m.type == 0x000B && m.size >= 16) {      // filter pipeline
        const int ver = m.p[0], nf = m.p[1];
        if (nf > 1) throw std::runtime_error("Error: dataset " + path + " uses several filters");
        if (nf == 1) {
          const uint8_t* f = m.p + (ver == 1 ? 8 : 2);
          if (g16(f) != 1) throw std::runtime_error("Error: dataset " + path + " uses a filter other than deflate");
          deflate = true;
        }
      } else if (m.type == 0x0008) {
        layout = m.p;
        layout_size = m.size;
      }
    }
    if (d.cls < 0 || !layout || d.elem == 0 || (d.elem > 8 && d.cls != 3)) throw std::runtime_error("Error: dataset " + path + " cannot be read");
    // deflate expands at most ~1032 times: a size beyond that cannot come out of this file (damaged header)
    if (d.n > (b_.size() * 1100ull + 4096) / d.elem) throw std::runtime_error("Error: dataset " + path + " is larger than the file can hold");
    const uint64_t bytes = d.n * d.elem;
    d.raw.assign(bytes, 0);
    if (layout[0] != 3 || layout_size < 3) throw std::runtime_error("Error: HDF5 data layout version " + std::to_string(layout[0]) + " is not supported");
    if (layout[1] == 0) {                                   // compact: the data is in the message
      const uint16_t sz = g16(layout + 2);
      if (sz < bytes || (uint64_t)4 + sz > layout_size) throw std::runtime_error("Error: dataset " + path + " is damaged");
      memcpy(d.raw.data(), layout + 4, bytes);
    } else if (layout[1] == 1) {                            // contiguous
      const uint64_t addr = g64(layout + 2);
      if (addr != kUndef && bytes) memcpy(d.raw.data(), at(addr, bytes), bytes);
    } else if (layout[1] == 2) {                            // chunked, rank 1 (+ the element size dimension)
      if (layout[2] != 2) throw std::runtime_error("Error: dataset " + path + " is not one-dimensional");
      const uint64_t bt = g64(layout + 3);
      const uint64_t cdim = g32(layout + 11);
      if (cdim == 0 || cdim > d.n + (1u << 20)) throw std::runtime_error("Error: dataset " + path + " has an implausible chunk size");
      if (bt != kUndef && bytes) read_chunks(bt, cdim, deflate, d, 0);
    } else {
      throw std::runtime_error("Error: unknown HDF5 data layout");
    }
    return d;
  }
  void read_chunks(uint64_t addr, uint64_t cdim, bool deflate, Data& d, int depth) const {
    if (depth > 16 || ++visits_ > 4000000) throw std::runtime_error("Error: HDF5 chunk B-tree too deep or cyclic");
    const uint8_t* n = at(addr, 24);
    if (memcmp(n, "TREE", 4) != 0 || n[4] != 1) throw std::runtime_error("Error: HDF5 chunk B-tree node expected");
    const unsigned level = n[5], used = g16(n + 6);
    const uint64_t key = 8 + 2 * 8;
    const uint8_t* body = at(addr + 24, (uint64_t)used * (key + 8) + key);
    for (unsigned i = 0; i < used; ++i) {
      const uint8_t* k = body + (uint64_t)i * (key + 8);
      const uint64_t child = g64(k + key);
      if (level > 0) { read_chunks(child, cdim, deflate, d, depth + 1); continue; }
      const uint32_t csize = g32(k), mask = g32(k + 4);
      const uint64_t off = g64(k + 8);                      // element offset of the chunk
      if (off >= d.n) continue;
      const uint64_t want = std::min<uint64_t>(cdim, d.n - off) * d.elem, full = cdim * d.elem;
      const uint8_t* src = at(child, csize);
      if (deflate && !(mask & 1)) {
        std::vector<uint8_t> tmp(full);
        uLongf got = (uLongf)full;
        if (uncompress(tmp.data(), &got, src, csize) != Z_OK || got < want) throw std::runtime_error("Error: damaged compressed chunk in the HDF5 file");
        memcpy(d.raw.data() + off * d.elem, tmp.data(), want);
      } else {
        if (csize < want) throw std::runtime_error("Error: damaged chunk in the HDF5 file");
        memcpy(d.raw.data() + off * d.elem, src, want);
      }
    }
  }

  std::vector<uint8_t> b_;
  uint64_t base_ = 0, root_header_ = 0;
  mutable uint64_t visits_ = 0;      // B-tree nodes visited: a damaged file may point a node at itself
  std::map<uint64_t, std::map<std::string, uint64_t>> cache_;
};

}  // namespace kb
