// FASTA/FASTQ reader for the command-line front end (plain or gzip, through zlib like the
// reference's kseq + gzread, src/kseq.h, src/common.h:216-225).  Record grammar follows kseq_read:
// a header line starting with '>' or '@'; sequence lines up to the next line starting with '>', '@'
// or '+'; for '+', a quality string at least as long as the sequence (possibly over several lines).
// Sequences are appended to a caller-owned batch (concatenated bases + offsets), the layout
// kb_pseudoalign_batch* takes; names and qualities are skipped (quant/bus never use them).
//
// ParallelFastx is the ingest path for plain (uncompressed) files: the file is mapped, cut into byte
// segments, and the segments are parsed concurrently with the same grammar.  A segment may only start
// where the sequential parser would start a record; that is not decidable locally (a quality line may
// begin with '@'), so every start is a guess that is then PROVEN: segment i, which starts at a proven
// boundary, must stop exactly on the guessed start of segment i+1.  If it does not, the remaining
// segments of the window are re-parsed sequentially from the proven position -- the result is always
// the sequential parse.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "fast_inflate.hpp"

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <future>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace kb {

struct ReadBatch {
  char* bases = nullptr;      // capacity cap_bases (+ slack)
  uint32_t* off = nullptr;    // capacity cap_reads + 1
  size_t cap_bases = 0, cap_reads = 0;
  size_t n = 0;               // reads in the batch
  uint32_t max_len = 0;
  bool eof = false;           // marker batch: the stream's current input file ended here (csrc/cli_main.cpp LockStep)
  void clear() { n = 0; max_len = 0; eof = false; if (off) off[0] = 0; }
  size_t n_bases() const { return off ? off[n] : 0; }
};

// Decompression of a regular .gz file on its own thread (csrc/fast_inflate.hpp), up to two chunks ahead of the
// parser.  Chunks are handed over by pointer: the decoder rotates through FastGz::kBufs buffers and a slot of the
// ring is reserved BEFORE a chunk is decoded, so a buffer is only reused after the parser has released it.
class GzPrefetch {
 public:
  explicit GzPrefetch(const std::string& path) : gz_(path) {
    th_ = std::thread([this] { run(); });
  }
  ~GzPrefetch() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    if (th_.joinable()) th_.join();
  }
  GzPrefetch(const GzPrefetch&) = delete;
  GzPrefetch& operator=(const GzPrefetch&) = delete;

  // next piece of decompressed data, valid until the next call; false at the end; rethrows decoder errors
  bool next_chunk(const char*& p, size_t& n) {
    std::unique_lock<std::mutex> lk(m_);
    if (held_) {                       // give the previous chunk back
      --outstanding_;
      tail_ = (tail_ + 1) % kRing;
      held_ = false;
      cv_.notify_all();
    }
    cv_.wait(lk, [&] { return ready_ > 0 || done_; });
    if (ready_ == 0) {
      if (err_) std::rethrow_exception(err_);
      return false;
    }
    p = ring_[tail_].p;
    n = ring_[tail_].n;
    --ready_;
    held_ = true;
    return true;
  }

 private:
  static constexpr size_t kRing = FastGz::kBufs - 1;     // chunks decoded and not yet released
  struct Slot { const char* p; size_t n; };
  void run() {
    try {
      for (;;) {
        {
          std::unique_lock<std::mutex> lk(m_);
          cv_.wait(lk, [&] { return outstanding_ < kRing || stop_; });
          if (stop_) return;
          ++outstanding_;              // reserve before decoding: the buffer about to be written is free
        }
        const char* p = nullptr;
        size_t n = 0;
        if (!gz_.next_chunk(p, n)) break;
        std::lock_guard<std::mutex> lk(m_);
        ring_[head_] = Slot{p, n};
        head_ = (head_ + 1) % kRing;
        ++ready_;
        cv_.notify_all();
      }
    } catch (...) {
      std::lock_guard<std::mutex> lk(m_);
      err_ = std::current_exception();
    }
    std::lock_guard<std::mutex> lk(m_);
    done_ = true;
    cv_.notify_all();
  }
  FastGz gz_;
  Slot ring_[kRing];
  size_t head_ = 0, tail_ = 0, ready_ = 0, outstanding_ = 0;
  bool held_ = false, done_ = false, stop_ = false;
  std::exception_ptr err_;
  std::mutex m_;
  std::condition_variable cv_;
  std::thread th_;
};

class FastxFile {
 public:
  // Regular gzip files are decoded by FastGz on a helper thread (KB_FASTGZ=0: zlib's gzread like the reference);
  // everything else (plain files, pipes, stdin-like paths) goes through gzread, which passes plain data through.
  explicit FastxFile(const std::string& path) : path_(path) {
    const char* knob = getenv("KB_FASTGZ");
    struct stat st;
    if (!(knob && knob[0] == '0') && stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode) && FastGz::looks_gzip(path)) {
      fast_.reset(new GzPrefetch(path));
      return;
    }
    buf_.resize(1 << 22);
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
      if (fast_) {
        const char* p = nullptr;
        size_t n = 0;
        if (!fast_->next_chunk(p, n)) { eof_ = true; pos_ = end_ = 0; return -1; }
        cur_ = p;
        pos_ = 0;
        end_ = n;
      } else {
        const int n = gzread(f_, buf_.data(), (unsigned)buf_.size());
        if (n < 0) throw std::runtime_error("Error: failed reading " + path_);
        cur_ = buf_.data();
        pos_ = 0;
        end_ = (size_t)n;
        if (n < (int)buf_.size()) eof_ = true;
        if (n == 0) return -1;
      }
    }
    return (unsigned char)cur_[pos_++];
  }
  // copy the rest of the current line to dst (may be null = discard); returns chars copied,
  // without the line terminator (a trailing '\r' is dropped); *got_nl tells whether '\n' was seen
  size_t rest_of_line(char* dst, size_t room, bool* got_nl) {
    size_t copied = 0;
    char last_c = 0;
    *got_nl = false;
    for (;;) {
      if (pos_ >= end_) {
        const int c = getc();
        if (c < 0) break;
        --pos_;
      }
      const char* s = cur_ + pos_;
      const size_t avail = end_ - pos_;
      const char* nl = (const char*)memchr(s, '\n', avail);
      const size_t take = nl ? (size_t)(nl - s) : avail;
      if (take) last_c = s[take - 1];
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
    // ks_getuntil2 (src/kseq.h:137) drops a trailing '\r' from copied and from discarded lines alike
    if (copied > 0 && last_c == '\r') --copied;
    return copied;
  }

  bool next(ReadBatch& b) {
    int c;
    if (stopped_) return false;
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
      if (q != len) {
        // kseq_read returns -2 (quality string of a different length); FastqSequenceReader::fetchSequences treats any
        // negative length as the end of this file (src/ProcessReads.cpp:3178-3182): the record is dropped, the file ends
        stopped_ = true;
        return false;
      }
    }
    b.off[b.n + 1] = b.off[b.n] + (uint32_t)len;
    if (len > b.max_len) b.max_len = (uint32_t)len;
    ++b.n;
    return true;
  }

  std::string path_;
  gzFile f_ = nullptr;
  std::unique_ptr<GzPrefetch> fast_;
  std::vector<char> buf_;
  const char* cur_ = nullptr;
  size_t pos_ = 0, end_ = 0;
  bool eof_ = false;
  bool stopped_ = false;      // a record with a quality string of the wrong length ended the file (kseq's -2)
  int last_ = 0;
};

// ---------------------------------------------------------------------------------------------
// Memory-range parser with FastxFile::next's grammar.  parse_range() appends the records whose header
// character lies in [pos, stop) and returns the position of the first header at or after `stop` (or
// `size` at the end of the data): the place where the sequential parser would begin its next record.
struct ParsedSegment {
  std::vector<char> bases;
  std::vector<uint32_t> cum{0};   // cum[i] = bases before read i of the segment; cum.size() = reads + 1
  uint32_t max_len = 0;
  size_t end_pos = 0;
  size_t n_reads() const { return cum.size() - 1; }
  void reset() { bases.clear(); cum.assign(1, 0); max_len = 0; end_pos = 0; }
};

inline size_t parse_range(const char* d, size_t size, size_t pos, size_t stop, ParsedSegment& out) {
  auto skip_line = [&](size_t p) -> size_t {   // position after the next '\n' (or size)
    const char* nl = (const char*)memchr(d + p, '\n', size - p);
    return nl ? (size_t)(nl - d) + 1 : size;
  };
  for (;;) {
    // header search: anything up to the next '>' or '@' is skipped, like kseq_read
    while (pos < size && d[pos] != '>' && d[pos] != '@') ++pos;
    if (pos >= size) return size;
    if (pos >= stop) return pos;
    pos = skip_line(pos + 1);                                   // name / comment
    const size_t first = out.bases.size();
    while (pos < size) {
      const char c = d[pos];
      if (c == '>' || c == '+' || c == '@') break;
      if (c == '\n') { ++pos; continue; }
      const char* nl = (const char*)memchr(d + pos, '\n', size - pos);
      size_t e = nl ? (size_t)(nl - d) : size;
      const size_t next = nl ? e + 1 : size;
      if (e > pos + 1 && d[e - 1] == '\r') --e;   // FastxFile keeps the first character of a line whatever it is
      out.bases.insert(out.bases.end(), d + pos, d + e);
      pos = next;
    }
    const size_t len = out.bases.size() - first;
    if (len > FastxFile::kMaxRead) throw std::runtime_error("Error: sequence too long");
    if (pos < size && d[pos] == '+') {
      pos = skip_line(pos + 1);                                 // rest of the '+' line
      size_t q = 0;
      for (;;) {                                                // at least one line, like kseq_read
        if (pos >= size) break;
        const char* nl = (const char*)memchr(d + pos, '\n', size - pos);
        size_t e = nl ? (size_t)(nl - d) : size;
        const bool got = nl != nullptr;
        const size_t next = nl ? e + 1 : size;
        // a trailing '\r' does not count (ks_getuntil2, src/kseq.h:137)
        q += (e > pos && d[e - 1] == '\r') ? e - pos - 1 : e - pos;
        pos = next;
        if (!got || q >= len) break;
      }
      if (q != len) {
        // kseq's -2: the record is dropped and the file ends here (see FastxFile::next); reporting the end of the data as
        // this segment's end makes the caller discard every later segment of the window and stop
        out.bases.resize(first);
        return size;
      }
    }
    if (out.bases.size() > 0xFFFFFFFFull) throw std::runtime_error("Error: parse window too large");
    out.cum.push_back((uint32_t)out.bases.size());
    if (len > out.max_len) out.max_len = (uint32_t)len;
  }
}

class ParallelFastx {
 public:
  // threads >= 2.  Throws if the file cannot be mapped; is_plain_regular() tells whether to try.
  ParallelFastx(const std::string& path, int threads) : path_(path), threads_(threads < 1 ? 1 : threads) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) throw std::runtime_error("Error: could not open file " + path);
    struct stat st;
    if (fstat(fd_, &st) != 0) { ::close(fd_); throw std::runtime_error("Error: could not stat file " + path); }
    size_ = (size_t)st.st_size;
    if (size_ > 0) {
      void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
      if (m == MAP_FAILED) { ::close(fd_); throw std::runtime_error("Error: could not map file " + path); }
      data_ = (const char*)m;
      madvise((void*)data_, size_, MADV_SEQUENTIAL);
    }
    int cap = 16, copy_cap = 4;               // defaults from the round-1 measurements; KB_FASTX_CAP / KB_FASTX_COPY: experiments
    if (const char* s = getenv("KB_FASTX_CAP")) { const int v = atoi(s); if (v > 0) cap = v; }
    if (const char* s = getenv("KB_FASTX_COPY")) { const int v = atoi(s); if (v > 0) copy_cap = v; }
    if (threads_ > cap) threads_ = cap;
    copy_threads_ = std::max(1, std::min(copy_cap, threads_ / 2));
    window_ = (size_t)4 << 20;                // bytes per parser thread and window
    if (const char* s = getenv("KB_FASTX_WINDOW")) { const long long v = atoll(s); if (v > 0) window_ = (size_t)v; }   // tests
    window_ *= (size_t)threads_;
    launch_next();
  }
  ~ParallelFastx() {
    if (pending_.valid()) pending_.wait();
    if (data_ && unmapped_ < size_) munmap((void*)(data_ + unmapped_), size_ - unmapped_);
    data_ = nullptr;
    if (getenv("KB_FASTX_DEBUG"))
      fprintf(stderr, "[fastx] %s: %zu windows, %zu segments, %zu rejected segment starts, parse %.3f s, wait %.3f s\n", path_.c_str(),
              n_windows_, n_segments_, n_fallbacks_, t_parse_, t_wait_);
    if (fd_ >= 0) ::close(fd_);
  }
  ParallelFastx(const ParallelFastx&) = delete;
  ParallelFastx& operator=(const ParallelFastx&) = delete;

  // plain (not gzip) regular file?
  static bool is_plain_regular(const std::string& path) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) return false;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    unsigned char m[2] = {0, 0};
    const size_t n = fread(m, 1, 2, f);
    fclose(f);
    return !(n == 2 && m[0] == 0x1f && m[1] == 0x8b);
  }

  // same contract as FastxFile::fill
  bool fill(ReadBatch& b, size_t max_reads) {
    const size_t n_before = b.n;
    for (;;) {
      const int r = fill_from_window(b, max_reads);       // copies what the current window offers
      if (r != kNeedWindow) break;
      if (!next_window()) break;
    }
    return b.n > n_before;
  }
  const std::string& path() const { return path_; }

 private:
  typedef std::vector<std::unique_ptr<ParsedSegment>> Window;

  enum { kBatchFull = 0, kNeedWindow = 1 };
  int fill_from_window(ReadBatch& b, size_t max_reads) {
    struct Job { const ParsedSegment* s; size_t rd, cnt; uint32_t dst_off; size_t dst_rd; };
    std::vector<Job> jobs;
    const size_t limit = std::min(max_reads, b.cap_reads);
    size_t n = b.n;
    uint64_t o = b.off[b.n];
    size_t total_bytes = 0;
    int status = kBatchFull;
    while (n < limit) {
      if (cur_.empty() || seg_ == cur_.size()) { status = kNeedWindow; break; }
      const ParsedSegment& s = *cur_[seg_];
      if (rd_ == s.n_reads()) { ++seg_; rd_ = 0; continue; }
      // as many whole reads of this segment as fit (FastxFile keeps kMaxRead bytes of head room)
      size_t take = std::min(s.n_reads() - rd_, limit - n);
      if (o + FastxFile::kMaxRead > b.cap_bases) break;
      const uint64_t room = b.cap_bases - FastxFile::kMaxRead - o;
      const uint32_t c0 = s.cum[rd_];
      bool full = false;
      if ((uint64_t)(s.cum[rd_ + take] - c0) > room) {
        // largest i with cum[rd_ + i] - c0 <= room
        const uint32_t* lo = s.cum.data() + rd_;
        const uint32_t* it = std::upper_bound(lo, lo + take + 1, (uint32_t)std::min<uint64_t>(room + c0, 0xFFFFFFFFull));
        take = (size_t)(it - lo) - 1;
        full = true;
        if (take == 0) break;
      }
      jobs.push_back({&s, rd_, take, (uint32_t)o, n});
      total_bytes += s.cum[rd_ + take] - c0;
      o += s.cum[rd_ + take] - c0;
      n += take;
      rd_ += take;
      if (full) break;                                             // base buffer full
    }
    auto run_job = [&b](const Job& j) {
      const uint32_t c0 = j.s->cum[j.rd];
      memcpy(b.bases + j.dst_off, j.s->bases.data() + c0, j.s->cum[j.rd + j.cnt] - c0);
      const uint32_t* c = j.s->cum.data() + j.rd;
      uint32_t* d = b.off + j.dst_rd;
      const uint32_t shift = j.dst_off - c0;     // modular arithmetic: dst = cum - c0 + dst_off
      for (size_t i = 1; i <= j.cnt; ++i) d[i] = c[i] + shift;
    };
    const size_t n_workers = total_bytes < ((size_t)4 << 20) ? 1 : std::min<size_t>((size_t)copy_threads_, jobs.size());
    if (n_workers <= 1) {
      for (const Job& j : jobs) run_job(j);
    } else {
      std::vector<std::future<void>> fu;
      for (size_t w = 0; w < n_workers; ++w)
        fu.push_back(std::async(std::launch::async, [&, w] { for (size_t i = w; i < jobs.size(); i += n_workers) run_job(jobs[i]); }));
      for (auto& f : fu) f.get();
    }
    for (const Job& j : jobs) if (j.s->max_len > b.max_len) b.max_len = j.s->max_len;   // upper bound
    b.n = n;
    return status;
  }

  // first position >= s that starts a line "@..." whose next-but-one line starts with '+', or npos
  size_t guess_start(size_t s, size_t limit) const {
    const char* nl = (const char*)memchr(data_ + s, '\n', size_ - s);
    size_t p = nl ? (size_t)(nl - data_) + 1 : size_;
    for (int tries = 0; p < limit && tries < 64; ++tries) {
      const char* n1 = (const char*)memchr(data_ + p, '\n', size_ - p);
      if (!n1) return (size_t)-1;
      const size_t l1 = (size_t)(n1 - data_) + 1;
      if (data_[p] == '@' && l1 < size_) {
        const char* n2 = (const char*)memchr(data_ + l1, '\n', size_ - l1);
        if (n2) {
          const size_t l2 = (size_t)(n2 - data_) + 1;
          if (l2 < size_ && data_[l2] == '+') return p;
        }
      }
      p = l1;
    }
    return (size_t)-1;
  }

  static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }

  Window parse_window(size_t begin, size_t end) {
    const double t0 = now_s();
    // begin is a proven record boundary; records whose header lies in [begin, end) belong to this window
    std::vector<size_t> starts(1, begin);
    const size_t step = std::max<size_t>((end - begin) / (size_t)threads_, 1);
    for (int t = 1; t < threads_; ++t) {
      const size_t want = begin + step * (size_t)t;
      if (want <= starts.back() || want >= end) continue;
      const size_t g = guess_start(want, end);
      if (g != (size_t)-1 && g > starts.back() && g < end) starts.push_back(g);
    }
    const size_t n = starts.size();
    Window w(n);
    std::vector<std::future<void>> fu;
    for (size_t i = 0; i < n; ++i) {
      w[i] = take_segment();
      const size_t a = starts[i], z = i + 1 < n ? starts[i + 1] : end;
      ParsedSegment* seg = w[i].get();
      fu.push_back(std::async(std::launch::async, [this, a, z, seg] {
        seg->bases.reserve((z - a) / 2 + 64);
        seg->cum.reserve((z - a) / 64 + 16);
        seg->end_pos = parse_range(data_, size_, a, z, *seg);
      }));
    }
    for (auto& f : fu) f.get();
    // proof: every segment must stop exactly where the next one started
    for (size_t i = 0; i + 1 < n; ++i) {
      if (w[i]->end_pos != starts[i + 1]) {
        // the guess was not a record boundary: everything after segment i is re-parsed from the proven position
        for (size_t j = i + 1; j < w.size(); ++j) give_segment(std::move(w[j]));
        w.resize(i + 1);
        w.emplace_back(take_segment());
        w.back()->end_pos = parse_range(data_, size_, w[i]->end_pos, end, *w.back());
        ++n_fallbacks_;
        break;
      }
    }
    ++n_windows_;
    n_segments_ += w.size();
    t_parse_ += now_s() - t0;
    return w;
  }

  // parsed-segment buffers are recycled: fresh 100 MB vectors would page-fault under the process-wide mmap lock
  std::unique_ptr<ParsedSegment> take_segment() {
    std::lock_guard<std::mutex> lk(pool_m_);
    if (pool_.empty()) return std::unique_ptr<ParsedSegment>(new ParsedSegment());
    std::unique_ptr<ParsedSegment> s = std::move(pool_.back());
    pool_.pop_back();
    s->reset();
    return s;
  }
  void give_segment(std::unique_ptr<ParsedSegment> s) {
    std::lock_guard<std::mutex> lk(pool_m_);
    pool_.push_back(std::move(s));
  }

  void launch_next() {
    if (next_begin_ >= size_) return;
    const size_t begin = next_begin_;
    const size_t end = std::min(size_, begin + window_);
    pending_ = std::async(std::launch::async, [this, begin, end] { return parse_window(begin, end); });
  }

  bool next_window() {
    for (auto& s : cur_) give_segment(std::move(s));
    cur_.clear();
    seg_ = rd_ = 0;
    if (!pending_.valid()) return false;
    const double t0 = now_s();
    cur_ = pending_.get();
    t_wait_ += now_s() - t0;
    next_begin_ = cur_.empty() ? size_ : cur_.back()->end_pos;   // proven: the sequential parser continues here
    // the file bytes before next_begin_ are parsed: give the pages back now, not in one long munmap at the end
    const size_t page = 4096, upto = std::min(next_begin_, size_) / page * page;
    if (upto > unmapped_) {
      munmap((void*)(data_ + unmapped_), upto - unmapped_);
      unmapped_ = upto;
    }
    launch_next();
    return true;
  }

  std::string path_;
  int threads_;
  int fd_ = -1;
  const char* data_ = nullptr;
  size_t size_ = 0, window_ = 0, next_begin_ = 0, unmapped_ = 0;
  int copy_threads_ = 1;
  std::future<Window> pending_;
  Window cur_;
  size_t seg_ = 0, rd_ = 0;
  size_t n_fallbacks_ = 0, n_windows_ = 0, n_segments_ = 0;
  double t_parse_ = 0, t_wait_ = 0;
  std::mutex pool_m_;
  std::vector<std::unique_ptr<ParsedSegment>> pool_;
};

// Either reader behind one interface: ParallelFastx for plain regular files when more than one parser
// thread is available, FastxFile (zlib) otherwise.
class FastxReader {
 public:
  FastxReader(const std::string& path, int threads) {
    if (threads > 1 && ParallelFastx::is_plain_regular(path)) par_.reset(new ParallelFastx(path, threads));
    else ser_.reset(new FastxFile(path));
  }
  bool fill(ReadBatch& b, size_t max_reads) { return par_ ? par_->fill(b, max_reads) : ser_->fill(b, max_reads); }
  bool parallel() const { return (bool)par_; }

 private:
  std::unique_ptr<ParallelFastx> par_;
  std::unique_ptr<FastxFile> ser_;
};

}  // namespace kb
