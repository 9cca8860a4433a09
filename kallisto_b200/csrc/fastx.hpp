// FASTA/FASTQ reader for the command-line front end (plain or gzip, through zlib like the
// reference's kseq + gzread, src/kseq.h, src/common.h:216-225).  Record grammar follows kseq_read:
// a header line starting with '>' or '@'; sequence lines up to the next line starting with '>', '@'
// or '+'; for '+', a quality string at least as long as the sequence (possibly over several lines).
// Sequences are appended to a caller-owned batch (concatenated bases + offsets), the layout
// kb_pseudoalign_batch* takes; names and qualities are skipped (quant/bus never use them).
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace kb {

struct ReadBatch {
  char* bases = nullptr;      // capacity cap_bases (+ slack)
  uint32_t* off = nullptr;    // capacity cap_reads + 1
  size_t cap_bases = 0, cap_reads = 0;
  size_t n = 0;               // reads in the batch
  uint32_t max_len = 0;
  void clear() { n = 0; max_len = 0; if (off) off[0] = 0; }
  size_t n_bases() const { return off ? off[n] : 0; }
};

class FastxFile {
 public:
  explicit FastxFile(const std::string& path) : path_(path), buf_(1 << 22) {
    f_ = gzopen(path.c_str(), "rb");
    if (!f_) throw std::runtime_error("Error: could not open file " + path);
    gzbuffer(f_, 1 << 20);
  }
  ~FastxFile() { if (f_) gzclose(f_); }
  FastxFile(const FastxFile&) = delete;
  FastxFile& operator=(const FastxFile&) = delete;

  // Append records to `b` until it holds max_reads reads, the next read might not fit, or the file
  // ends.  Returns false once the file is exhausted and nothing was appended.
  bool fill(ReadBatch& b, size_t max_reads) {
    size_t added = 0;
    while (b.n < max_reads && b.n < b.cap_reads) {
      // a record of unknown length is coming: keep a generous margin in the base buffer
      if (b.n_bases() + kMaxRead > b.cap_bases) break;
      if (!next(b)) break;
      ++added;
    }
    return added > 0;
  }
  bool eof() const { return eof_ && pos_ >= end_ && last_ == 0; }
  const std::string& path() const { return path_; }

  static constexpr size_t kMaxRead = 1 << 20;   // longest single sequence accepted (bases)

 private:
  int getc() {
    if (pos_ >= end_) {
      if (eof_) return -1;
      const int n = gzread(f_, buf_.data(), (unsigned)buf_.size());
      if (n < 0) throw std::runtime_error("Error: failed reading " + path_);
      pos_ = 0;
      end_ = (size_t)n;
      if (n < (int)buf_.size()) eof_ = true;
      if (n == 0) return -1;
    }
    return (unsigned char)buf_[pos_++];
  }
  // copy the rest of the current line to dst (may be null = discard); returns chars copied,
  // without the line terminator (a trailing '\r' is dropped); *got_nl tells whether '\n' was seen
  size_t rest_of_line(char* dst, size_t room, bool* got_nl) {
    size_t copied = 0;
    *got_nl = false;
    for (;;) {
      if (pos_ >= end_) {
        const int c = getc();
        if (c < 0) break;
        --pos_;
      }
      const char* s = buf_.data() + pos_;
      const size_t avail = end_ - pos_;
      const char* nl = (const char*)memchr(s, '\n', avail);
      const size_t take = nl ? (size_t)(nl - s) : avail;
      if (dst) {
        if (copied + take > room) throw std::runtime_error("Error: sequence too long in " + path_);
        memcpy(dst + copied, s, take);
      }
      copied += take;
      pos_ += take;
      if (nl) {
        ++pos_;
        *got_nl = true;
        break;
      }
    }
    if (copied > 0 && dst && dst[copied - 1] == '\r') --copied;
    return copied;
  }

  bool next(ReadBatch& b) {
    int c;
    if (last_ == 0) {
      while ((c = getc()) != -1 && c != '>' && c != '@') {}
      if (c == -1) return false;
    }
    last_ = 0;
    bool nl;
    rest_of_line(nullptr, 0, &nl);                      // name / comment
    char* dst = b.bases + b.off[b.n];
    const size_t room = b.cap_bases - b.off[b.n];
    size_t len = 0;
    while ((c = getc()) != -1 && c != '>' && c != '+' && c != '@') {
      if (c == '\n') continue;
      if (len + 1 > room) throw std::runtime_error("Error: sequence too long in " + path_);
      dst[len++] = (char)c;
      len += rest_of_line(dst + len, room - len, &nl);
    }
    if (c == '>' || c == '@') last_ = c;                // header of the next record already consumed
    if (c == '+') {
      rest_of_line(nullptr, 0, &nl);                    // rest of the '+' line
      size_t q = 0;
      for (;;) {                                        // at least one line, like kseq_read
        bool got;
        q += rest_of_line(nullptr, 0, &got);
        if (!got || q >= len) break;                    // EOF inside the quality string, or complete
      }
    }
    b.off[b.n + 1] = b.off[b.n] + (uint32_t)len;
    if (len > b.max_len) b.max_len = (uint32_t)len;
    ++b.n;
    return true;
  }

  std::string path_;
  gzFile f_ = nullptr;
  std::vector<char> buf_;
  size_t pos_ = 0, end_ = 0;
  bool eof_ = false;
  int last_ = 0;
};

}  // namespace kb
