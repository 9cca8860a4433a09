// Host-side engine: owns the device tables of one index and the device state of one
// quantification run.  This is the C++ layer right under the C ABI (include/kallisto_b200.h);
// it mirrors the reference objects that sit on the hot path:
//
//   kb::Index  <->  KmerIndex after KmerIndex::load            (src/KmerIndex.cpp:1330-1559)
//   kb::Quant  <->  MinCollector + MasterProcessor/ReadProcessor (src/MinCollector.h:17-119,
//                   src/ProcessReads.cpp:307-483, 934-1237) followed by EMAlgorithm / Bootstrap
//
// There is no CPU implementation of any of the per-read or per-iteration work in here:
// without a CUDA device every entry point throws.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "index_v13.hpp"
#include "kernels.hpp"

namespace kb {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { release(); }
  void alloc(size_t count);
  void release();
  void upload(const T* src, size_t count, cudaStream_t st = 0);
  void download(T* dst, size_t count, size_t offset = 0, cudaStream_t st = 0) const;
  void zero(cudaStream_t st = 0);
};

// Grow-only device work buffers.  They belong to the Index and are lent to one run at a time, so that
// consecutive runs on the same index (the normal case) do not pay cudaMalloc again; a second run
// created while the first is still alive gets private ones.
struct BatchWs {
  DBuf<uint8_t> stage_b[2][2];      // double-buffered input staging: H2D of batch i+1 overlaps the kernels of batch i
  DBuf<uint32_t> stage_o[2][2];
  DBuf<uint32_t> d_qcount, d_qentries, d_scratch, d_packed;
  DBuf<uint32_t> d_spill, d_qbig_count, d_qbig;   // fragments with more than KB_MAX_E distinct EC sets
  DBuf<int32_t> d_handles;
  DBuf<uint16_t> d_tl;
  DBuf<uint8_t> d_skip;             // per fragment: holds a D-list k-mer (only when the index has a D-list)
};

struct EmWs {   // grow-only device workspace of run_em_device
  DBuf<uint32_t> used, scal, idx_in, order, handle, count, len, multi_len, is_multi, ec_off, m_off, multi_index;
  DBuf<unsigned long long> key_in, key_out;
  DBuf<uint8_t> tmp;
  DBuf<uint32_t> ec_tid, multi_ec, m_rowoff, m_tid, m_row, m_iota, sortv, t_deg, t_off, t_midx;
  DBuf<uint32_t> minkey, ckey, cval, ckey_out, rlen;     // row order of the EM matrices (emprep_rows)
  DBuf<unsigned long long> k64_in, k64_out;              // CSC sort keys
  DBuf<unsigned> bar;                                    // grid-barrier counter of em_kernel
  DBuf<uint32_t> cnt_row;                                // row-ordered counts (launch_em fills them)
  DBuf<double> single_cnt;
  // bootstrap over the same matrices (run_bootstrap_device)
  DBuf<uint32_t> bs_counts, bs_x0;
  DBuf<double> bs_alpha, bs_norm, bs_cp;
  DBuf<int> bs_emi;
  DBuf<unsigned int> bs_ch;
  DBuf<double> m_w, t_w, eff, alpha, norm;
  DBuf<int32_t> t_single;
  DBuf<int> emi;
  DBuf<unsigned int> chcount;
};

class Index {
 public:
  static std::unique_ptr<Index> load(const std::string& path, int device, bool load_positions, int threads);
  ~Index();

  FlatIndex flat;
  int device = 0;
  DevIndex dev{};
  uint64_t table_cap = 0;
  uint64_t dict_cap = 0;          // set-dictionary capacity used by every run on this index
  uint32_t empty_ec = 0xFFFFFFFFu;
  uint32_t max_set_len = 0;
  uint32_t n_index_tids = 0;      // pool entries occupied by the index's own EC sets
  double load_seconds = 0, build_seconds = 0;

  DBuf<KmerSlot> slots;
  DBuf<unsigned long long> dfk;   // D-list k-mer set
  DBuf<uint32_t> filter;          // presence filter of the k-mer table (L2-resident, see DevIndex)
  size_t l2_persist_bytes = 0;    // persisting-L2 carve-out set aside for it
  DBuf<uint32_t> ec_off;
  DBuf<uint32_t> index_pool;      // the index's EC sets (copied to the front of every run's pool)
  DBuf<unsigned long long> dslots_init;
  DBuf<int32_t> ec_handle;
  std::vector<int32_t> h_ec_handle;   // host copy: index EC-set id -> dictionary handle
  DBuf<uint32_t> blk_ec;
  DBuf<uint64_t> blk_strand_off;
  DBuf<uint8_t> strand;
  BatchWs shared_bws;
  EmWs* shared_emws = nullptr;
  bool ws_in_use = false;
  DBuf<uint4> fp_info;            // only when loaded with positions
  DBuf<uint32_t> blk_usize, target_len;
};

// One NCCL communicator per process/GPU plus the receive area of the merge (csrc/comm.cu).
struct CommImpl;
class Comm {
 public:
  static void unique_id(void* out128);                                  // ncclGetUniqueId (rank 0, then broadcast by the caller)
  Comm(int n_ranks, int rank, const void* id128, int device);            // ncclCommInitRank
  Comm(void* nccl_comm, int n_ranks, int rank, int device, bool take_ownership);   // an existing ncclComm_t
  static std::vector<Comm*> init_all(const std::vector<int>& devices);   // one process, one communicator per device
  ~Comm();
  Comm(const Comm&) = delete;
  Comm& operator=(const Comm&) = delete;
  void reserve(size_t n_sets_per_rank, size_t n_entries_per_rank);       // root: size the receive area ahead of time
  int n_ranks = 1, rank = 0, device = 0;
  CommImpl* impl_ = nullptr;
};

struct QuantOptions {
  int paired = 1;          // !opt.single_end
  int strand_mode = 0;     // 0 unstranded, 1 --fr-stranded, 2 --rf-stranded
  int collect_fld = 1;     // opt.fld == 0: estimate the fragment-length distribution from the data
  uint32_t max_batch_reads = 1u << 22;     // staging capacity (reads per batch)
  uint64_t max_batch_bases = 1ull << 29;   // staging capacity (bases per batch)
  int threads_per_block = 256;
  int fp_fl = -1;          // >= 0: fragment-position filter with this mean fragment length (!single_overhang && -l given)
  bool bus = false;        // `kallisto bus` run: records instead of (only) counts
  BusSpec bus_spec{};
  int refill_min = 16;     // match_kernel: finished lanes per warp that trigger a finalise + refill round
};

// Equivalence classes of a finished run, ids in order of first occurrence (== reference -t 1).
struct EcTable {
  std::vector<uint64_t> off;      // n_ec + 1
  std::vector<uint32_t> tid;
  std::vector<uint32_t> count;
  std::vector<int32_t> handle;    // device handle of each EC (to translate per-fragment results)
  uint32_t n() const { return (uint32_t)count.size(); }
};

struct EmResult {
  std::vector<double> alpha;      // est_counts
  std::vector<double> eff_lens;
  int rounds = 0;
  double seconds = 0;
};

struct Stats {
  uint64_t n_processed = 0, n_pseudoaligned = 0, n_unique = 0;
  uint64_t n_probes = 0, n_slot_visits = 0, n_resolved = 0, n_memo_hits = 0;
};

class Quant {
 public:
  Quant(Index& ix, const QuantOptions& opt);
  ~Quant();

  // One batch of reads (mates interleaved when paired).  `off` has n_reads+1 entries or is null
  // when every read has `fixed_len` bases.  Pointers are HOST memory; the copy to the device, the
  // kernels and (if handles_out != null) the copy back of one handle per fragment happen inside.
  void pseudoalign_host(const char* bases, const uint32_t* off, uint32_t n_reads, uint32_t fixed_len,
                        int32_t* handles_out);
  // Paired batch with one buffer per mate (what a FASTQ reader produces: R1 and R2 parsed
  // separately), n_pairs fragments; off1/off2 have n_pairs + 1 entries or are null with fixed_len.
  void pseudoalign_host_pe(const char* bases1, const uint32_t* off1, const char* bases2, const uint32_t* off2,
                           uint32_t n_pairs, uint32_t fixed_len, int32_t* handles_out);
  // `kallisto bus`: one batch of read sets (bases[k]/offs[k] = file k of the technology, n_sets + 1
  // offsets each).  Writes the BUS records of the pseudoaligned sets, in read order, EC ids final.
  void bus_batch_host(const char* const* bases, const uint32_t* const* offs, uint32_t n_sets, BusRecord* records_out,
                      uint32_t* n_records_out);
  uint32_t bus_batch_device(const uint8_t* const* d_bases, const uint32_t* const* d_offs, uint32_t n_sets, uint32_t max_seq_len);
  const BusRecord* bus_records_device() const { return bus_rec_.p; }
  void bus_lengths(uint32_t* bc_hist, uint32_t* umi_hist);
  // Batch mode (`bus -x BULK`, src/ProcessReads.cpp:371-404,1603-1607): the read sets that follow belong to sample
  // `barcode` (the fake barcode of their records); its fragment-length sampling starts from an empty histogram.
  void bus_begin_sample(uint64_t barcode);
  // Same, inputs already resident in device memory; handles stay on the device
  // (device_handles(), valid until the next batch).
  void pseudoalign_device(const uint8_t* d_bases, const uint32_t* d_off, uint32_t n_reads, uint32_t fixed_len,
                          uint32_t max_read_len);
  const int32_t* device_handles() const { return bws_->d_handles.p; }
  void sync();

  // MasterProcessor tail flush + EC id assignment.
  const EcTable& finalize_ecs();
  const std::vector<uint32_t>& flens() const { return flens_; }
  void set_flens(const uint32_t* f);     // e.g. after an all-reduce across ranks
  Stats stats();

  // Effective lengths from the fragment-length distribution (or a given mean/sd), then the EM.
  std::vector<double> mean_fl_trunc(double fld_mean, double fld_sd) const;
  EmResult run_em(const EcTable& ecs, const std::vector<double>& fl_trunc, int max_iter = 10000, int min_rounds = 50);
  // Same result, EC table built and kept on the device (the `quant` fast path).
  EmResult run_em_device(const std::vector<double>& fl_trunc, int max_iter = 10000, int min_rounds = 50);
  // B bootstrap EMs (Bootstrap::run_em): alpha_out is B x n_targets.  Returns rounds per bootstrap.
  std::vector<int> run_bootstrap(const EcTable& ecs, const std::vector<double>& fl_trunc, uint64_t seed, int B,
                                 std::vector<double>& alpha_out, std::vector<uint32_t>* samples_out = nullptr);

  // ---- multi-GPU: ship this rank's equivalence classes to another rank / fold another rank's in ----
  // export_prepare numbers the ECs (first occurrence) and lays the table out on the device; returns
  // {n_sets, n_entries}.  export_copy then fills caller-provided DEVICE buffers (e.g. torch tensors
  // about to go through NCCL): off[n_sets+1], tids[n_entries], counts[n_sets], first[n_sets].
  void export_prepare(uint32_t* n_sets, uint32_t* n_entries);
  void export_copy(uint32_t* d_off, uint32_t* d_tids, uint32_t* d_counts, unsigned long long* d_first);
  void import_sets_device(uint32_t n_sets, const uint32_t* d_off, const uint32_t* d_tids, const uint32_t* d_counts,
                          const unsigned long long* d_first, unsigned long long first_offset);
  void add_processed(uint64_t n) { n_frag_total_ += n; }
  // The whole exchange in one collective call (csrc/comm.cu): tables gathered to rank 0 with NCCL send/recv and
  // folded in by content with one kernel launch, fragment-length samples completed in rank order.  Returns the
  // number of fragments processed by all ranks.
  uint64_t merge_to_root(Comm& comm, uint64_t first_stride);
  // Same merge when all runs live in THIS process (one host thread drives several GPUs, `kallisto_b200 quant --devices`):
  // the other runs' tables are copied with cudaMemcpyPeerAsync (NVLink) -- no communicator to set up.  Called on the root.
  uint64_t merge_local(const std::vector<Quant*>& others, uint64_t first_stride);
  // Global index of the first fragment of the NEXT batch (multi-GPU drivers that deal batches of one read
  // stream to several runs: first-occurrence order then is the order of the stream).  Default: running count.
  void set_frag_base(uint64_t base) { frag_base_ = base; have_frag_base_ = true; }
  // Size the EM / EC-numbering workspace ahead of time (no cudaMalloc on the first kb_em_run).
  void reserve_em(size_t n_ecs, size_t n_entries);

  // Same result on the EM matrices run_em_device left on the device (no EC table on the host, no second set-up);
  // the B problems are solved `chunk` at a time so that their alpha / norm vectors stay in L2.  ms_out (optional):
  // {resample ms, EM ms} measured with CUDA events.
  std::vector<int> run_bootstrap_device(const std::vector<double>& fl_trunc, uint64_t seed, int B, std::vector<double>& alpha_out,
                                        std::vector<uint32_t>* samples_out = nullptr, double* ms_out = nullptr);
  bool dev_problem_valid() const { return dev_problem_valid_; }
  // Run on a caller-provided stream (e.g. the framework's current stream) instead of the run's own.
  void set_stream(cudaStream_t st);
  // Per-kernel device time, measured with CUDA events on the launching stream.
  struct Timings { double match_ms = 0, resolve_ms = 0, pack_ms = 0; uint64_t match_launches = 0, resolve_launches = 0; };
  void enable_timing(bool on) { timing_ = on; }
  Timings timings();

  Index& index() { return ix_; }
  const QuantOptions& options() const { return opt_; }
  cudaStream_t stream() const { return stream_; }
  double last_em_seconds = 0, last_prep_seconds = 0, last_bs_resample_ms = 0, last_bs_em_ms = 0;
  uint64_t n_kernel_launches = 0;   // launches of this library's own kernels by this run (CUB's are not counted)
  // filled by run_em_device
  bool dev_stats_valid_ = false, dev_problem_valid_ = false;
  uint32_t dev_n_multi_ = 0;
  uint64_t dev_n_ecs_ = 0, dev_nnz_ = 0, dev_pseudoaligned_ = 0, dev_unique_ = 0;

 private:
  void run_batch(const uint8_t* d_bases, const uint32_t* d_off, uint32_t n_reads, uint32_t fixed_len,
                 uint32_t max_read_len, const uint8_t* d_bases2 = nullptr, const uint32_t* d_off2 = nullptr);
  void check_device_errors();
  void apply_l2_window();
  uint32_t bus_core(const uint8_t* const* db, const uint32_t* const* dofs, uint32_t n_sets, uint32_t maxlen);

  Index& ix_;
  QuantOptions opt_;
  cudaStream_t stream_ = nullptr;
  bool own_stream_ = true;
  bool timing_ = false;
  std::vector<cudaEvent_t> events_;   // triples
  Timings tacc_;
  DevDict dd_{};
  // run state on the device
  DBuf<uint32_t> pool_;
  DBuf<Memo2Entry> m2_;
  DBuf<unsigned long long> dslots_, first_, mn_key_, counters_;   // counters_: pool_top, tpool_top, stats[4], one 128-byte line each
  DBuf<uint32_t> count_, tpool_;
  DBuf<int32_t> mn_val_;
  DBuf<int> error_;
  // batch staging
  BatchWs* bws_ = nullptr;
  bool own_ws_ = false;
  cudaStream_t copy_stream_ = nullptr;
  cudaEvent_t ev_copied_[2] = {nullptr, nullptr}, ev_done_[2] = {nullptr, nullptr};
  int stage_idx_ = 0;
  uint32_t n_resolve_warps_ = 0, scratch_stride_ = 0, resolve_group_ = 32;
  uint32_t* h_off_pinned_ = nullptr;
  // host-side run state
  uint64_t n_frag_total_ = 0;
  std::vector<uint32_t> flens_;
  uint32_t tlencount_ = 0;
  std::vector<uint16_t> tl_list_;        // the samples behind flens_, in read order (shipped to rank 0 by merge_to_root)
  uint64_t frag_base_ = 0;
  bool have_frag_base_ = false;
  std::vector<uint16_t> h_tl_;
  EcTable ecs_;
  bool ecs_valid_ = false;
  struct EmWs* emws_ = nullptr;
  // bus mode
  DBuf<uint8_t> bus_b_[4], bus_skip_, bus_notag_;
  DBuf<uint32_t> bus_o_[4], bus_flags_, bus_hist_, bus_isnew_, bus_newrank_, bus_ismapped_, bus_rank_;
  DBuf<unsigned long long> bus_bc_, bus_umi_, bus_nvalid_;
  DBuf<int32_t> bus_idof_;
  DBuf<BusRecord> bus_rec_;
  DBuf<uint8_t> bus_tmp_;
  uint32_t exp_n_ = 0, exp_nnz_ = 0;
  DBuf<uint32_t> lm_off_, lm_tids_, lm_counts_;       // merge_local: receive area on the root
  DBuf<unsigned long long> lm_first_;
  uint32_t bus_next_id_ = 0;
  uint64_t bus_valid_total_ = 0, bus_sample_base_ = 0;
  const uint8_t* cur_skip_ = nullptr;
  uint32_t cur_start_ = 0, cur_start2_ = 0, cur_alt_start_ = 0, cur_alt_start2_ = 0;
  const uint8_t* cur_notag_ = nullptr;
};

std::vector<double> mean_fl_trunc_of(const uint32_t* flens /* 1000 */, double fld_mean, double fld_sd);

// `kallisto quant-tcc` (src/main.cpp:2802-3220): one EM per sample (row of a transcript-compatibility-count matrix) over
// one shared equivalence-class table (the lines of matrix.ec), on the device in chunks of samples.
struct TccInput {
  uint32_t n_ecs = 0;
  const uint64_t* ec_off = nullptr;     // n_ecs + 1
  const uint32_t* tids = nullptr;       // sorted transcript ids of every EC
  uint32_t n_samples = 0;
  const uint64_t* row_off = nullptr;    // n_samples + 1 offsets into ec_ids / counts
  const uint32_t* ec_ids = nullptr;
  const uint32_t* counts = nullptr;
  const double* eff_lens = nullptr;     // n_targets, or n_samples x n_targets when per_sample_eff
  bool per_sample_eff = false;
};
std::vector<int> tcc_run(Index& ix, const TccInput& in, std::vector<double>& alpha_out /* n_samples x n_targets */);

}  // namespace kb
