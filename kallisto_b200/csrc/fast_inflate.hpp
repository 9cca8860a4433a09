// Streaming gzip (RFC 1952 / DEFLATE RFC 1951) decoder for the command-line front end.
//
// The reference reads .gz input with zlib's gzread (src/kseq.h, src/common.h:216-225); on this path the
// inflate is the bottleneck once pseudoalignment runs on the device (zlib 1.3: ~270 MB/s of FASTQ text per
// stream).  This decoder produces the same bytes with the usual modern recipe -- 64-bit bit buffer refilled
// without branches, one table lookup per symbol with the extra-bit counts folded into the entry, word-wise
// match copies -- over the memory-mapped compressed file.  Multi-member files are concatenated like gzread
// does; CRC-32 and ISIZE of every member are verified (zlib's crc32()).  Anything malformed throws.
//
// Not a general-purpose library: input must be a regular file (it is mapped), output is pulled in chunks
// that stay valid until the next call.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace kb {

// CRC-32 (gzip polynomial, reflected) by carry-less multiplication folding (Gopal et al., "Fast CRC computation for
// generic polynomials using PCLMULQDQ", Intel 2009): ~2x zlib 1.3's braided tables; verified against crc32() in the
// tests.  Falls back to zlib where the CPU lacks PCLMULQDQ.
#if defined(__x86_64__)
__attribute__((target("pclmul,sse4.1"))) inline uint32_t crc32_clmul(uint32_t crc, const uint8_t* p, size_t n) {
  if (n < 64) return (uint32_t)crc32(crc, p, (uInt)n);
  const __m128i k1k2 = _mm_set_epi64x(0x00000001c6e41596, 0x0000000154442bd4);
  const __m128i k3k4 = _mm_set_epi64x(0x00000000ccaa009e, 0x00000001751997d0);
  const __m128i k5 = _mm_set_epi64x(0, 0x0000000163cd6124);
  const __m128i poly = _mm_set_epi64x(0x00000001f7011641, 0x00000001db710641);
  const __m128i mask32 = _mm_set_epi32(0, 0, 0, ~0);
  __m128i x1 = _mm_loadu_si128((const __m128i*)(p + 0)), x2 = _mm_loadu_si128((const __m128i*)(p + 16));
  __m128i x3 = _mm_loadu_si128((const __m128i*)(p + 32)), x4 = _mm_loadu_si128((const __m128i*)(p + 48));
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)~crc));
  p += 64;
  n -= 64;
  while (n >= 64) {
    const __m128i t1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), t2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
    const __m128i t3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), t4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
    x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11);
    x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
    x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11);
    x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, t1), _mm_loadu_si128((const __m128i*)(p + 0)));
    x2 = _mm_xor_si128(_mm_xor_si128(x2, t2), _mm_loadu_si128((const __m128i*)(p + 16)));
    x3 = _mm_xor_si128(_mm_xor_si128(x3, t3), _mm_loadu_si128((const __m128i*)(p + 32)));
    x4 = _mm_xor_si128(_mm_xor_si128(x4, t4), _mm_loadu_si128((const __m128i*)(p + 48)));
    p += 64;
    n -= 64;
  }
#define KB_CRC_FOLD(a, b)                                       \
  do {                                                          \
    const __m128i t_ = _mm_clmulepi64_si128(a, k3k4, 0x00);     \
    a = _mm_clmulepi64_si128(a, k3k4, 0x11);                    \
    a = _mm_xor_si128(_mm_xor_si128(a, t_), b);                 \
  } while (0)
  KB_CRC_FOLD(x1, x2);
  KB_CRC_FOLD(x1, x3);
  KB_CRC_FOLD(x1, x4);
  while (n >= 16) {
    const __m128i d = _mm_loadu_si128((const __m128i*)p);
    KB_CRC_FOLD(x1, d);
    p += 16;
    n -= 16;
  }
#undef KB_CRC_FOLD
  __m128i t = _mm_clmulepi64_si128(x1, k3k4, 0x10);              // 128 -> 64 bits
  x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), t);
  t = _mm_and_si128(x1, mask32);
  x1 = _mm_srli_si128(x1, 4);
  t = _mm_clmulepi64_si128(t, k5, 0x00);
  x1 = _mm_xor_si128(x1, t);
  t = _mm_and_si128(x1, mask32);                                 // Barrett reduction
  t = _mm_clmulepi64_si128(t, poly, 0x10);
  t = _mm_and_si128(t, mask32);
  t = _mm_clmulepi64_si128(t, poly, 0x00);
  x1 = _mm_xor_si128(x1, t);
  const uint32_t c = ~(uint32_t)_mm_extract_epi32(x1, 1);
  return n ? (uint32_t)crc32(c, p, (uInt)n) : c;
}
inline uint32_t fast_crc32(uint32_t crc, const uint8_t* p, size_t n) {
  static const bool have = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
  if (have) return crc32_clmul(crc, p, n);
  while (n > 0) {
    const size_t step = n < ((size_t)1 << 30) ? n : ((size_t)1 << 30);
    crc = (uint32_t)crc32(crc, p, (uInt)step);
    p += step;
    n -= step;
  }
  return crc;
}
#else
inline uint32_t fast_crc32(uint32_t crc, const uint8_t* p, size_t n) {
  while (n > 0) {
    const size_t step = n < ((size_t)1 << 30) ? n : ((size_t)1 << 30);
    crc = (uint32_t)crc32(crc, p, (uInt)step);
    p += step;
    n -= step;
  }
  return crc;
}
#endif

class FastGz {
 public:
  explicit FastGz(const std::string& path) : path_(path) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) throw std::runtime_error("Error: could not open file " + path);
    struct stat st;
    if (fstat(fd_, &st) != 0 || !S_ISREG(st.st_mode)) { ::close(fd_); throw std::runtime_error("Error: not a regular file " + path); }
    size_ = (size_t)st.st_size;
    if (size_ > 0) {
      void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
      if (m == MAP_FAILED) { ::close(fd_); throw std::runtime_error("Error: could not map file " + path); }
      in_ = (const uint8_t*)m;
      madvise((void*)in_, size_, MADV_SEQUENTIAL);
    }
    for (auto& b : bufs_) b.resize(kHist + kChunk + kSlack);
    out_ = bufs_[0].data();
    op_ = rp_ = crc_from_ = floor_ = kHist;
  }
  ~FastGz() {
    if (in_) munmap((void*)in_, size_);
    if (fd_ >= 0) ::close(fd_);
  }
  FastGz(const FastGz&) = delete;
  FastGz& operator=(const FastGz&) = delete;

  // gzip magic?
  static bool looks_gzip(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    unsigned char m[3] = {0, 0, 0};
    const size_t n = fread(m, 1, 3, f);
    fclose(f);
    return n == 3 && m[0] == 0x1f && m[1] == 0x8b && m[2] == 8;
  }

  // Next piece of decompressed data; false at the end of the file.  The decoder rotates through kBufs output
  // buffers, so a chunk stays valid until kBufs - 1 further calls have been made (GzPrefetch hands chunks to the
  // parser without copying them).
  static constexpr int kBufs = 4;
  bool next_chunk(const char*& p, size_t& n) {
    while (rp_ == op_) {
      if (state_ == kEnd) return false;
      // the last 32 KiB are the match history of what comes next: carry them to the front of the next buffer
      if (op_ > kHist) {
        flush_crc();
        const size_t delta = op_ - kHist;
        cur_buf_ = (cur_buf_ + 1) % kBufs;
        uint8_t* next = bufs_[cur_buf_].data();
        memcpy(next, out_ + delta, kHist);
        out_ = next;
        op_ = rp_ = crc_from_ = kHist;
        floor_ = floor_ > delta ? floor_ - delta : 0;
      }
      decode(kHist + kChunk);
    }
    p = (const char*)out_ + rp_;
    n = op_ - rp_;
    rp_ = op_;
    return true;
  }

 private:
  static constexpr size_t kHist = 32768, kChunk = (size_t)4 << 20, kSlack = 1024;
  enum State { kMemberHeader, kBlockHeader, kStored, kHuffman, kTrailer, kEnd };

  [[noreturn]] void bad(const char* what) const { throw std::runtime_error("Error: corrupt gzip data in " + path_ + " (" + what + ")"); }

  // ---- bit input (LSB first) ----
  inline void refill() {
    if (ip_ + 8 <= size_) {
      uint64_t w;
      memcpy(&w, in_ + ip_, 8);
      bitbuf_ |= w << bitcnt_;
      ip_ += (size_t)((63 - bitcnt_) >> 3);
      bitcnt_ |= 56;
    } else {
      while (bitcnt_ <= 56 && ip_ < size_) {
        bitbuf_ |= (uint64_t)in_[ip_++] << bitcnt_;
        bitcnt_ += 8;
      }
    }
  }
  inline uint32_t peek(int n) const { return (uint32_t)(bitbuf_ & ((1ull << n) - 1)); }
  inline void drop(int n) { bitbuf_ >>= n; bitcnt_ -= n; }
  inline uint32_t take(int n) {
    if (bitcnt_ < n) { refill(); if (bitcnt_ < n) bad("unexpected end of data"); }
    const uint32_t v = peek(n);
    drop(n);
    return v;
  }
  // bytes that have been loaded into the bit buffer but not consumed are given back
  void align_to_byte() {
    drop(bitcnt_ & 7);
    ip_ -= (size_t)(bitcnt_ >> 3);
    bitbuf_ = 0;
    bitcnt_ = 0;
  }

  // ---- Huffman tables ----
  // entry: bits 0-7 code length consumed at this level; bit 8: literal; bit 9: sub-table link; bit 10: end of block;
  //        bits 11-15: number of extra bits; bits 16-31: literal value / base length / base distance / sub-table start
  static constexpr int kLitBits = 11, kDistBits = 8;
  static constexpr uint32_t kLit = 1u << 8, kSub = 1u << 9, kEob = 1u << 10;
  static constexpr uint32_t kLit2 = 1u << 11;    // literal entries only: a second literal in bits 24-31

  void build(const uint8_t* lens, int n, int table_bits, bool is_litlen, std::vector<uint32_t>& tab) {
    static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    int count[16] = {0};
    for (int i = 0; i < n; ++i) ++count[lens[i]];
    count[0] = 0;
    // over-subscribed or incomplete sets (a single code of length 1 is allowed by zlib for distances)
    int left = 1, n_codes = 0;
    for (int l = 1; l <= 15; ++l) {
      left <<= 1;
      left -= count[l];
      if (left < 0) bad("over-subscribed Huffman code");
      n_codes += count[l];
    }
    // zlib (inftrees.c): an incomplete set is only accepted when it is empty or a single 1-bit code
    if (left > 0 && !(n_codes == 0 || (n_codes == 1 && count[1] == 1))) bad("incomplete Huffman code");
    (void)is_litlen;
    uint32_t next_code[16];
    {
      uint32_t code = 0;
      for (int l = 1; l <= 15; ++l) {
        code = (code + (uint32_t)count[l - 1]) << 1;
        next_code[l] = code;
      }
    }
    auto entry_for = [&](int sym, int consumed) -> uint32_t {
      if (is_litlen) {
        if (sym < 256) return (uint32_t)consumed | kLit | ((uint32_t)sym << 16);      // bits 24-31 stay 0
        if (sym == 256) return (uint32_t)consumed | kEob;
        if (sym > 285) return 0;                                             // invalid symbol: flagged at use (length 0)
        return (uint32_t)consumed | ((uint32_t)len_extra[sym - 257] << 11) | ((uint32_t)len_base[sym - 257] << 16);
      }
      if (sym > 29) return 0;
      return (uint32_t)consumed | ((uint32_t)dist_extra[sym] << 11) | ((uint32_t)dist_base[sym] << 16);
    };
    auto rev = [](uint32_t c, int l) { uint32_t r = 0; for (int i = 0; i < l; ++i) { r = (r << 1) | (c & 1); c >>= 1; } return r; };
    const uint32_t tsize = 1u << table_bits;
    tab.assign(tsize, 0);                                                     // 0 = no code: flagged at use
    // sub-tables: for every primary prefix, the longest code sharing it
    std::vector<uint8_t> sub_bits(tsize, 0);
    std::vector<uint32_t> codes((size_t)n);
    for (int i = 0; i < n; ++i) {
      const int l = lens[i];
      if (!l) continue;
      codes[i] = rev(next_code[l]++, l);
      if (l > table_bits) {
        const uint32_t pre = codes[i] & (tsize - 1);
        if (l - table_bits > sub_bits[pre]) sub_bits[pre] = (uint8_t)(l - table_bits);
      }
    }
    for (uint32_t pre = 0; pre < tsize; ++pre)
      if (sub_bits[pre]) {
        const uint32_t start = (uint32_t)tab.size();
        if (start + (1u << sub_bits[pre]) > 0xFFFF) bad("Huffman table too large");
        tab[pre] = (uint32_t)table_bits | kSub | ((uint32_t)sub_bits[pre] << 11) | (start << 16);
        tab.resize(start + (1u << sub_bits[pre]), 0);
      }
    for (int i = 0; i < n; ++i) {
      const int l = lens[i];
      if (!l) continue;
      if (l <= table_bits) {
        const uint32_t e = entry_for(i, l);
        for (uint32_t j = codes[i]; j < tsize; j += 1u << l) tab[j] = e;
      } else {
        const uint32_t pre = codes[i] & (tsize - 1);
        const uint32_t start = tab[pre] >> 16, sb = (tab[pre] >> 11) & 31;
        const uint32_t e = entry_for(i, l - table_bits);
        for (uint32_t j = codes[i] >> table_bits; j < (1u << sb); j += 1u << (l - table_bits)) tab[start + j] = e;
      }
    }
    // Two literals per lookup where both codes fit in the primary index (skewed alphabets such as FASTQ text
    // have 2-4 bit codes for their frequent bytes): index = code(a) | code(b) << len(a).
    if (is_litlen) {
      int shorts[256], n_short = 0;
      for (int a = 0; a < 256 && a < n; ++a)
        if (lens[a] && lens[a] < table_bits) shorts[n_short++] = a;
      for (int ia = 0; ia < n_short; ++ia) {
        const int a = shorts[ia], la = lens[a];
        for (int ib = 0; ib < n_short; ++ib) {
          const int b = shorts[ib], lb = lens[b];
          if (la + lb > table_bits) continue;
          const uint32_t e = (uint32_t)(la + lb) | kLit | kLit2 | ((uint32_t)a << 16) | ((uint32_t)b << 24);
          for (uint32_t j = codes[a] | (codes[b] << la); j < tsize; j += 1u << (la + lb)) tab[j] = e;
        }
      }
    }
  }

  void fixed_tables() {
    uint8_t l[288];
    for (int i = 0; i < 144; ++i) l[i] = 8;
    for (int i = 144; i < 256; ++i) l[i] = 9;
    for (int i = 256; i < 280; ++i) l[i] = 7;
    for (int i = 280; i < 288; ++i) l[i] = 8;
    build(l, 288, kLitBits, true, lit_);
    uint8_t d[32];
    for (int i = 0; i < 32; ++i) d[i] = 5;
    build(d, 32, kDistBits, false, dist_);
  }

  void dynamic_tables() {
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    const int hlit = (int)take(5) + 257, hdist = (int)take(5) + 1, hclen = (int)take(4) + 4;
    if (hlit > 286 || hdist > 30) bad("too many length or distance symbols");
    uint8_t cl[19] = {0};
    for (int i = 0; i < hclen; ++i) cl[order[i]] = (uint8_t)take(3);
    std::vector<uint32_t> ct;
    build_codelen(cl, ct);
    uint8_t lens[320];
    int i = 0;
    while (i < hlit + hdist) {
      if (bitcnt_ < 15 + 7) refill();
      const uint32_t e = ct[peek(7)];
      if ((e & 0xFF) == 0) bad("invalid code length code");
      if ((int)(e & 0xFF) > bitcnt_) bad("unexpected end of data");
      drop((int)(e & 0xFF));
      const int sym = (int)(e >> 16);
      if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
      int rep, val = 0;
      if (sym == 16) {
        if (i == 0) bad("repeat without a previous length");
        val = lens[i - 1];
        rep = 3 + (int)take(2);
      } else if (sym == 17) {
        rep = 3 + (int)take(3);
      } else {
        rep = 11 + (int)take(7);
      }
      if (i + rep > hlit + hdist) bad("code length repeat runs past the end");
      while (rep--) lens[i++] = (uint8_t)val;
    }
    if (lens[256] == 0) bad("no end-of-block code");
    build(lens, hlit, kLitBits, true, lit_);
    build(lens + hlit, hdist, kDistBits, false, dist_);
  }
  // the 19-symbol code that describes the code lengths: direct 7-bit table, entry = len | sym << 16
  void build_codelen(const uint8_t* cl, std::vector<uint32_t>& tab) {
    int count[8] = {0};
    for (int i = 0; i < 19; ++i) ++count[cl[i]];
    count[0] = 0;
    int left = 1, n_codes = 0;
    for (int l = 1; l <= 7; ++l) { left <<= 1; left -= count[l]; if (left < 0) bad("over-subscribed code length code"); n_codes += count[l]; }
    if (left > 0) bad("incomplete code length code");
    (void)n_codes;
    uint32_t next_code[8], code = 0;
    for (int l = 1; l <= 7; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next_code[l] = code; }
    tab.assign(128, 0);
    for (int i = 0; i < 19; ++i) {
      const int l = cl[i];
      if (!l) continue;
      uint32_t c = next_code[l]++, r = 0;
      for (int b = 0; b < l; ++b) { r = (r << 1) | (c & 1); c >>= 1; }
      for (uint32_t j = r; j < 128; j += 1u << l) tab[j] = (uint32_t)l | ((uint32_t)i << 16);
    }
  }

  // ---- gzip member framing ----
  void member_header() {
    if (ip_ >= size_) { state_ = kEnd; return; }
    // gzread treats trailing garbage after a complete member as the end of the data
    if (size_ - ip_ < 18 || in_[ip_] != 0x1f || in_[ip_ + 1] != 0x8b) {
      if (n_members_ == 0) bad("not in gzip format");
      state_ = kEnd;
      return;
    }
    if (in_[ip_ + 2] != 8) bad("unknown compression method");
    const uint8_t flg = in_[ip_ + 3];
    size_t p = ip_ + 10;
    auto need = [&](size_t k) { if (p + k > size_) bad("truncated header"); };
    if (flg & 4) { need(2); const size_t xl = in_[p] | ((size_t)in_[p + 1] << 8); p += 2; need(xl); p += xl; }
    if (flg & 8) { while (true) { need(1); if (in_[p++] == 0) break; } }
    if (flg & 16) { while (true) { need(1); if (in_[p++] == 0) break; } }
    if (flg & 2) { need(2); p += 2; }
    ip_ = p;
    bitbuf_ = 0;
    bitcnt_ = 0;
    crc_ = 0;
    isize_ = 0;
    ++n_members_;
    floor_ = op_;                 // members are independent streams: no match may reach into the previous one
    state_ = kBlockHeader;
  }
  void member_trailer() {
    align_to_byte();
    if (ip_ + 8 > size_) bad("truncated trailer");
    flush_crc();
    uint32_t c, n;
    memcpy(&c, in_ + ip_, 4);
    memcpy(&n, in_ + ip_ + 4, 4);
    ip_ += 8;
    if (c != (uint32_t)crc_) bad("crc mismatch");
    if (n != (uint32_t)isize_) bad("length mismatch");
    state_ = kMemberHeader;
  }
  // brings crc_ / isize_ up to date with everything decoded so far
  void flush_crc() {
    if (op_ > crc_from_) {
      crc_ = fast_crc32((uint32_t)crc_, out_ + crc_from_, op_ - crc_from_);
      isize_ += op_ - crc_from_;
      crc_from_ = op_;
    }
  }

  // ---- decoding: fills out_ up to `limit` (or the end of the data) ----
  void decode(size_t limit) {
    for (;;) {
      switch (state_) {
        case kEnd:
          return;
        case kMemberHeader:
          member_header();
          break;
        case kTrailer:
          member_trailer();
          break;
        case kBlockHeader: {
          final_ = take(1) != 0;
          const uint32_t type = take(2);
          if (type == 0) {
            align_to_byte();
            if (ip_ + 4 > size_) bad("truncated stored block");
            const uint32_t len = in_[ip_] | ((uint32_t)in_[ip_ + 1] << 8), nlen = in_[ip_ + 2] | ((uint32_t)in_[ip_ + 3] << 8);
            if ((len ^ 0xFFFFu) != nlen) bad("stored block length check");
            ip_ += 4;
            stored_left_ = len;
            state_ = kStored;
          } else if (type == 1) {
            fixed_tables();
            state_ = kHuffman;
          } else if (type == 2) {
            dynamic_tables();
            state_ = kHuffman;
          } else {
            bad("invalid block type");
          }
          break;
        }
        case kStored: {
          const size_t room = limit > op_ ? limit - op_ : 0;
          if (room == 0) return;
          const size_t n = stored_left_ < room ? stored_left_ : room;
          if (ip_ + n > size_) bad("truncated stored block");
          memcpy(out_ + op_, in_ + ip_, n);
          ip_ += n;
          op_ += n;
          stored_left_ -= n;
          if (stored_left_ == 0) state_ = final_ ? kTrailer : kBlockHeader;
          if (op_ >= limit) return;
          break;
        }
        case kHuffman:
          if (huffman(limit)) state_ = final_ ? kTrailer : kBlockHeader;
          if (op_ >= limit) return;
          break;
      }
    }
  }

  // true when the end-of-block symbol was consumed; false when the output limit was reached first.
  // The bit reader lives in locals here (the members would be reloaded around every store through `out`).
  bool huffman(size_t limit) {
    uint8_t* const out = out_;
    size_t op = op_;
    const uint32_t* const lit = lit_.data();
    const uint32_t* const dst = dist_.data();
    const uint32_t lmask = (1u << kLitBits) - 1, dmask = (1u << kDistBits) - 1;
    const uint8_t* const in = in_;
    const size_t in_fast_end = size_ >= 8 ? size_ - 8 : 0;    // refilling with one 8-byte load is allowed up to here
    size_t ip = ip_;
    uint64_t bb = bitbuf_;
    int bc = bitcnt_;
    const size_t floor = floor_;
    bool eob = false;
    const char* err = nullptr;
#define KB_REFILL()                                                   \
  do {                                                                \
    if (ip <= in_fast_end) {                                          \
      uint64_t w_;                                                    \
      memcpy(&w_, in + ip, 8);                                        \
      bb |= w_ << bc;                                                 \
      ip += (size_t)((63 - bc) >> 3);                                 \
      bc |= 56;                                                       \
    } else {                                                          \
      while (bc <= 56 && ip < size_) { bb |= (uint64_t)in[ip++] << bc; bc += 8; } \
    }                                                                 \
  } while (0)
    // ---- fast loop: at least 8 input bytes ahead, so every refill leaves >= 56 valid bits and no availability
    //      checks are needed (longest chain: 15 + 5 + 15 + 13 = 48 bits for a match, 3 x 11 for the literal run)
    // The table entry of the NEXT symbol is looked up as soon as its index bits are known (a refill only adds
    // bits above the ones already counted), so its L1 latency overlaps the stores of the current symbol.
#define KB_FAST_REFILL()                       \
  do {                                         \
    uint64_t w_;                               \
    memcpy(&w_, in + ip, 8);                   \
    bb |= w_ << bc;                            \
    ip += (size_t)((63 - bc) >> 3);            \
    bc |= 56;                                  \
  } while (0)
    if (op < limit && ip <= in_fast_end) {
      KB_FAST_REFILL();
      uint32_t e = lit[bb & lmask];
      for (;;) {
        // invariant: e = lit[bb & lmask] for the current bit position, bc >= 48 valid bits
        if (e & kSub) {
          bb >>= (e & 0xFF);
          bc -= (int)(e & 0xFF);
          e = lit[(e >> 16) + (uint32_t)(bb & ((1u << ((e >> 11) & 31)) - 1))];
        }
        int cl = (int)(e & 0xFF);
        if (cl == 0) { err = "invalid literal/length code"; break; }
        bb >>= cl;
        bc -= cl;
        if (e & kLit) {
          uint32_t e2 = lit[bb & lmask];
          out[op] = (uint8_t)(e >> 16);
          out[op + 1] = (uint8_t)(e >> 24);
          op += 1 + ((e >> 11) & 1);
          if ((e2 & (kLit | kSub)) == kLit) {           // primary literal entries have 1 <= cl <= 11
            cl = (int)(e2 & 0xFF);
            bb >>= cl;
            bc -= cl;
            e = lit[bb & lmask];
            out[op] = (uint8_t)(e2 >> 16);
            out[op + 1] = (uint8_t)(e2 >> 24);
            op += 1 + ((e2 >> 11) & 1);
            if ((e & (kLit | kSub)) == kLit) {
              cl = (int)(e & 0xFF);
              bb >>= cl;
              bc -= cl;
              out[op] = (uint8_t)(e >> 16);
              out[op + 1] = (uint8_t)(e >> 24);
              op += 1 + ((e >> 11) & 1);
              if (!(op < limit && ip <= in_fast_end)) break;
              KB_FAST_REFILL();
              e = lit[bb & lmask];
              continue;
            }
          } else {
            e = e2;
          }
          // e is the entry at the current position (bc >= 56 - 15 - 11 - 11 = 19 >= 11 index bits): top up and go on
          if (!(op < limit && ip <= in_fast_end)) break;
          KB_FAST_REFILL();
          continue;
        }
        if (e & kEob) { eob = true; break; }
        const int leb = (int)((e >> 11) & 31);
        uint32_t len = (e >> 16) + (uint32_t)(bb & ((1u << leb) - 1));
        bb >>= leb;
        bc -= leb;
        uint32_t d = dst[bb & dmask];
        if (d & kSub) {
          bb >>= (d & 0xFF);
          bc -= (int)(d & 0xFF);
          d = dst[(d >> 16) + (uint32_t)(bb & ((1u << ((d >> 11) & 31)) - 1))];
        }
        const int dl = (int)(d & 0xFF);
        if (dl == 0) { err = "invalid distance code"; break; }
        const int deb = (int)((d >> 11) & 31);
        bb >>= dl;
        const uint32_t distance = (d >> 16) + (uint32_t)(bb & ((1u << deb) - 1));
        bb >>= deb;
        bc -= dl + deb;
        if (distance > op - floor) { err = "distance too far back"; break; }
        const uint8_t* src = out + op - distance;
        uint8_t* dp = out + op;
        uint8_t* const end = dp + len;
        op += len;
        // next symbol: refill and look up before the copy
        const bool more = op < limit && ip <= in_fast_end;
        if (more) {
          KB_FAST_REFILL();
          e = lit[bb & lmask];
        }
        if (distance >= 16) {
          do { memcpy(dp, src, 16); dp += 16; src += 16; } while (dp < end);     // may overshoot into the slack
        } else if (distance >= 8) {
          do { memcpy(dp, src, 8); dp += 8; src += 8; } while (dp < end);
        } else if (distance == 1) {
          const uint64_t v = 0x0101010101010101ull * *src;
          do { memcpy(dp, &v, 8); dp += 8; } while (dp < end);
        } else {
          while (dp < end) *dp++ = *src++;
        }
        if (!more) break;
      }
    }
#undef KB_FAST_REFILL
    // ---- careful loop: the last bytes of the input (and the remainder after an error-free fast loop)
    // every iteration may write one match of up to 258 bytes (+ word-copy overshoot inside kSlack)
    while (!err && !eob && op < limit) {
      KB_REFILL();                                // >= 56 bits unless the input is exhausted
      uint32_t e = lit[bb & lmask];
      if (e & kSub) {
        bb >>= (e & 0xFF);
        bc -= (int)(e & 0xFF);
        e = lit[(e >> 16) + (uint32_t)(bb & ((1u << ((e >> 11) & 31)) - 1))];
      }
      int cl = (int)(e & 0xFF);
      if (cl == 0) { err = "invalid literal/length code"; break; }
      if (cl > bc) { err = "unexpected end of data"; break; }
      bb >>= cl;
      bc -= cl;
      if (e & kLit) {
        // one or two literals per entry: both bytes are stored, the position advances by 1 or 2
        out[op] = (uint8_t)(e >> 16);
        out[op + 1] = (uint8_t)(e >> 24);
        op += 1 + ((e >> 11) & 1);
        // up to two more lookups without refilling (3 x 15 bits <= 56)
        e = lit[bb & lmask];
        cl = (int)(e & 0xFF);
        if ((e & (kLit | kSub)) == kLit && cl <= bc && cl) {
          bb >>= cl;
          bc -= cl;
          out[op] = (uint8_t)(e >> 16);
          out[op + 1] = (uint8_t)(e >> 24);
          op += 1 + ((e >> 11) & 1);
          e = lit[bb & lmask];
          cl = (int)(e & 0xFF);
          if ((e & (kLit | kSub)) == kLit && cl <= bc && cl) {
            bb >>= cl;
            bc -= cl;
            out[op] = (uint8_t)(e >> 16);
            out[op + 1] = (uint8_t)(e >> 24);
            op += 1 + ((e >> 11) & 1);
          }
        }
        continue;
      }
      if (e & kEob) { eob = true; break; }
      // length: base + extra bits (<= 5)
      const int leb = (int)((e >> 11) & 31);
      uint32_t len = (e >> 16) + (uint32_t)(bb & ((1u << leb) - 1));
      bb >>= leb;
      bc -= leb;
      if (bc < 15 + 13) KB_REFILL();
      uint32_t d = dst[bb & dmask];
      if (d & kSub) {
        bb >>= (d & 0xFF);
        bc -= (int)(d & 0xFF);
        d = dst[(d >> 16) + (uint32_t)(bb & ((1u << ((d >> 11) & 31)) - 1))];
      }
      const int dl = (int)(d & 0xFF);
      if (dl == 0) { err = "invalid distance code"; break; }
      const int deb = (int)((d >> 11) & 31);
      if (dl + deb > bc) { err = "unexpected end of data"; break; }
      bb >>= dl;
      const uint32_t distance = (d >> 16) + (uint32_t)(bb & ((1u << deb) - 1));
      bb >>= deb;
      bc -= dl + deb;
      if (distance > op - floor) { err = "distance too far back"; break; }
      const uint8_t* src = out + op - distance;
      uint8_t* dp = out + op;
      op += len;
      if (distance >= 8) {
        // no overlap at word granularity: 8 bytes at a time (may overshoot into the slack behind the limit)
        const uint8_t* const end = dp + len;
        do { memcpy(dp, src, 8); dp += 8; src += 8; } while (dp < end);
      } else if (distance == 1) {
        memset(dp, *src, len);
      } else {
        while (len--) *dp++ = *src++;
      }
    }
#undef KB_REFILL
    op_ = op;
    ip_ = ip;
    bitbuf_ = bb;
    bitcnt_ = bc;
    if (err) bad(err);
    return eob;
  }

 private:
  std::string path_;
  int fd_ = -1;
  const uint8_t* in_ = nullptr;
  size_t size_ = 0, ip_ = 0;
  uint64_t bitbuf_ = 0;
  int bitcnt_ = 0;
  std::vector<uint8_t> bufs_[kBufs];
  int cur_buf_ = 0;
  uint8_t* out_ = nullptr;
  size_t op_ = 0, rp_ = 0, crc_from_ = 0;
  size_t floor_ = 0;     // oldest position in out_ a match may reference
  State state_ = kMemberHeader;
  bool final_ = false;
  size_t stored_left_ = 0;
  std::vector<uint32_t> lit_, dist_;
  uLong crc_ = 0;
  uint64_t isize_ = 0;
  size_t n_members_ = 0;
};

}  // namespace kb
