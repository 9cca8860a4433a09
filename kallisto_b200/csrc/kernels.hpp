// Host-visible kernel argument blocks and launch wrappers (plain structs, no torch types).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "kb_device.cuh"

namespace kb {

struct TableBuildArgs {
  const uint8_t* useq;
  const uint64_t* useq_byteoff;
  const uint64_t* skmer;
  const uint64_t* kstart;     // n_unitigs + 1: number of k-mers before each unitig
  const uint64_t* blk_off;
  const uint32_t* blk_lb;
  const uint32_t* blk_ub;
  const uint32_t* blk_ec;     // per block: set handle
  uint32_t n_long, n_unitigs;
  int k;
  uint64_t n_kmers;
  KmerSlot* slots;
  uint64_t mask;
  uint32_t* filter;         // optional presence filter (zeroed by the caller)
  uint32_t filter_mask;
  int* error;
};

struct DictInitArgs {
  const uint32_t* ec_off;
  const uint32_t* pool;
  uint32_t n_ec;
  unsigned long long* dslots;
  uint64_t dmask;
  int32_t* ec_handle;
};

// One batch of reads for the pseudoalignment kernels (ReadProcessor::processBuffer's `seqs`,
// src/ProcessReads.cpp:968-1046): concatenated ASCII bases, mates interleaved when paired.
struct BatchArgs {
  const uint8_t* bases;
  const uint32_t* off;      // n_reads + 1 offsets into bases, or nullptr when every read has fixed_len bases
  const uint8_t* bases2;    // optional: second-mate buffer (then `bases`/`off` hold the first mates only and
  const uint32_t* off2;     //           read 2f+m is entry f of buffer m); nullptr = mates interleaved in `bases`
  uint32_t fixed_len;
  uint32_t n_frag;          // pairs (paired) or reads (single)
  int paired;
  int strand_mode;          // 0 unstranded, 1 FR (--fr-stranded), 2 RF (--rf-stranded)
  uint64_t frag_base;       // global index of fragment 0
  int32_t* handle_out;      // per fragment: set handle, or KB_H_UNMAPPED
  uint16_t* tl_out;         // per fragment fragment-length candidate (mapPair), or nullptr
  uint32_t* q_count;        // resolve queue
  uint32_t* q_entries;      // stride KB_Q_STRIDE
  uint32_t* spill;          // KB_SPILL words per resident lane of match_kernel: set handles beyond KB_MAX_E
  uint32_t* qbig_count;     // wide resolve queue (fragments that hit more than KB_MAX_E distinct EC sets)
  uint32_t* qbig_entries;   // stride KB_QBIG_STRIDE
  uint32_t qbig_cap;        // entries
  const uint32_t* packed;   // pack_kernel output: per read, nb 64-bit base words then nb 32-bit invalid masks
  uint32_t nb;              // 32-base words per packed read = ceil(max_read_len / 32)
  uint32_t pstride;         // 32-bit words per packed read (multiple of 8 = 32 bytes)
  uint32_t empty_ec;        // handle of the empty index EC set, or 0xFFFFFFFF
  int refill_min;           // finished lanes of a warp that trigger a finalise + refill round
  const uint8_t* skip;      // optional per fragment: 1 = treat as having no sequence (bus: bad barcode/UMI; D-list hit)
  uint8_t* skip_w;          // the same array, writable: set when the index has a D-list (dlist_scan_kernel marks fragments)
  int fp_fl;                // >= 0: apply the fragment-position filter of ProcessReads.cpp:1095-1136 with this mean fragment length
  uint32_t start;           // first base of every read that is matched (bus: BUSOptionSubstr.start of the sequence)
  uint32_t start2;          // the same for the second mate of a pair (paired bus technologies, e.g. STORM-seq: 14)
  // UMI tag sequences (bus --tag / SMARTSEQ3, src/ProcessReads.cpp:1512-1530): notag[f] = 1 marks a fragment without the
  // tag ("ignore_umi"): its reads start at alt_start / alt_start2, the strand filter is off for it, and ONLY such
  // fragments sample fragment lengths.  nullptr = no tag sequence in play.
  const uint8_t* notag;
  uint32_t alt_start, alt_start2;
};
static constexpr int KB_Q_STRIDE = 2 + KB_MAX_E + 6;   // frag, n|flags, handles, 2 strand words, 4 position-filter words
static constexpr int KB_SPILL = 112;                   // a fragment may hit KB_MAX_E + KB_SPILL = 128 distinct EC sets
static constexpr int KB_QBIG_STRIDE = 2 + KB_MAX_E + KB_SPILL + 6;
static constexpr uint32_t KB_QBIG_CAP = 1u << 16;      // wide-queue entries per batch

struct ResolveArgs {
  uint32_t* scratch;        // per lane group: scratch_stride entries
  uint32_t scratch_stride;
  uint32_t n_warps;         // number of lane groups (one fragment each at a time)
  uint32_t group;           // lanes per group: 32, 16, 8 or 4
};

void launch_fill_u64(unsigned long long* p, uint64_t n, unsigned long long v, cudaStream_t st);
void launch_fill_i32(int32_t* p, uint64_t n, int32_t v, cudaStream_t st);
void launch_fill_f64(double* p, uint64_t n, double v, cudaStream_t st);
void launch_fill_memo2(Memo2Entry* p, uint64_t n, cudaStream_t st);
void launch_build_table(const TableBuildArgs& a, cudaStream_t st);
void launch_dict_init(const DictInitArgs& a, cudaStream_t st);

// Pseudoalignment of one batch: match kernel (thread per fragment) + resolve kernel (warp per
// queued fragment) [+ fragment-length finalisation].
// ev (optional): three events recorded before match_kernel, between the kernels, after resolve_kernel.
void launch_pseudoalign(const DevIndex& ix, const DevDict& dd, const BatchArgs& ba, const ResolveArgs& ra,
                        int threads_per_block, cudaStream_t st, cudaEvent_t* ev = nullptr);
void launch_fld_finalize(const DevDict& dd, const BatchArgs& ba, cudaStream_t st);
void launch_import_sets(const DevDict& dd, uint32_t n_sets, const uint32_t* off, const uint32_t* tids, const uint32_t* counts,
                        const unsigned long long* first, unsigned long long first_offset, cudaStream_t st);
// Same for the tables of several ranks at once (Quant::merge_to_root): one launch, warp per incoming set.
struct ImportSeg {
  uint32_t n_sets;
  const uint32_t* off;
  const uint32_t* tids;
  const uint32_t* counts;
  const unsigned long long* first;
};
static constexpr int KB_IMPORT_SEGS = 16;
void launch_import_segments(const DevDict& dd, const ImportSeg* segs, int n_segs, cudaStream_t st);
int device_sm_count();
// Compact the handles with count > 0: used[0..*n_used)
void launch_collect_used(const DevDict& dd, uint32_t* used, uint32_t* n_used, cudaStream_t st);

// ---- EM / bootstrap (kernels_em.cu) ----
struct EmProblem {
  // structure shared by all problems of a batch
  uint32_t n_ec, n_targets;
  // ECs with >= 2 members and their members, CSR in EC-id order (denominator pass)
  uint32_t n_multi;
  const uint32_t* multi_ec;     // n_multi: EC id
  const uint32_t* m_off;        // n_multi + 1
  const uint32_t* m_tid;        // nnz
  const double* m_w;            // nnz: counts_orig[ec] / eff_len[tid]  (calc_weights, src/weights.cpp:220-246)
  // CSC by transcript over the same nnz, entries in increasing EC id (numerator pass)
  const uint32_t* t_off;        // n_targets + 1
  const uint32_t* t_midx;       // nnz: index into the multi arrays (row of the EC)
  const double* t_w;            // nnz
  const int32_t* t_single;      // n_targets: EC id of the singleton EC {t}, or -1
  // per problem (nb of them)
  int nb;
  const uint32_t* counts;       // nb x n_ec
  double* alpha;                // nb x n_targets (in/out)
  double* norm;                 // nb x n_multi scratch: counts/denom or 0
  int* rounds;                  // nb: iterations run (the reference's "ran for i rounds")
  unsigned* bar;                // arrival counter of the kernel's grid barrier (zeroed by launch_em)
  // filled by launch_em before the EM kernel starts (one dependent load less per row and per round):
  uint32_t* cnt_row;            // nb x n_multi: counts of the multi-transcript ECs in row order
  double* single_cnt;           // nb x n_targets: count of the singleton EC {t}, or 0
  int* fstate;                  // nb: final state (2 finished, 3 finished + host must zero small alphas)
  unsigned int* chcount;        // nb x 2 (double-buffered) change counters
  int max_iter, min_rounds;
  // 0: the weights m_w / t_w are shared by all problems (quant, bootstrap); nnz: problem b has its own at [b * w_stride]
  // (quant-tcc: the weights are the SAMPLE's counts / eff_len, src/weights.cpp:220-246)
  uint64_t w_stride;
};
int em_max_blocks(int threads_per_block);
// quant-tcc helpers: dense per-sample count vectors from the sparse TCC rows, and per-sample weights in CSR and CSC order
struct TccFill {
  uint32_t n_ec, n_targets, nb;        // samples in this chunk
  const unsigned long long* row_off;    // chunk's rows: nb + 1 offsets into ec_ids / vals
  const uint32_t* ec_ids;
  const uint32_t* vals;
  uint32_t* counts;                     // nb x n_ec (zeroed by the launcher)
  // entry -> (EC id, transcript) of the CSR (m_*) and CSC (t_*) layouts
  uint64_t nnz;
  const uint32_t* m_ec; const uint32_t* m_tid; const uint32_t* t_ec; const uint32_t* t_tid;
  const double* eff; uint64_t eff_stride;   // nb x n_targets when eff_stride == n_targets, shared when 0
  double* m_w; double* t_w;             // nb x nnz
};
void launch_tcc_fill(const TccFill& a, cudaStream_t st);
void launch_em(const EmProblem& p, int threads_per_block, cudaStream_t st);

// Device-side EM problem construction (kernels_emprep.cu)
struct EmPrep {
  uint32_t n_ec, n_multi, n_targets;
  // per EC (id = order of first occurrence); arrays of n_ec + 1 where a scan total is stored
  uint32_t* handle;
  uint32_t* count;
  uint32_t* len;           // n_ec + 1
  uint32_t* ec_off;        // n_ec + 1: offsets of the EC table
  uint32_t* m_off;         // n_ec + 1: offsets into the multi-EC entry arrays (0-length for singletons)
  uint32_t* multi_index;   // n_ec + 1: after the scan the rank among the multi-transcript ECs, after emprep_rows the ROW of the EC
  uint32_t* minkey;        // n_ec: smallest transcript id of the EC
  uint32_t* ec_tid;        // EC table entries
  // multi-transcript ECs, CSR
  uint32_t* multi_ec;      // n_multi
  uint32_t* m_rowoff;      // n_multi + 1
  uint32_t* m_tid;
  double* m_w;
  uint32_t* m_row;         // entry -> row
  uint32_t* m_iota;        // entry -> entry (values of the CSC sort)
  unsigned long long* k64_in;   // entry -> tid << 32 | EC id (keys of the CSC sort)
  // CSC
  uint32_t* t_deg;         // n_targets + 1 (zeroed by the caller)
  uint32_t* t_off;         // n_targets + 1
  uint32_t* t_midx;
  double* t_w;
  int32_t* t_single;       // n_targets (filled with -1 by the caller)
  const double* eff;       // n_targets
};
size_t emprep_sort_bytes(uint32_t n_used, uint32_t nnz_max);
void emprep_sort_by_first(const DevDict& dd, const uint32_t* used, uint32_t n_used, unsigned long long* key_in,
                          unsigned long long* key_out, uint32_t* idx_in, uint32_t* order_out, void* tmp, size_t tmp_bytes,
                          cudaStream_t st);
void emprep_meta(const DevDict& dd, const uint32_t* used, const uint32_t* order, uint32_t n, const EmPrep& p,
                 uint32_t* multi_len, uint32_t* is_multi, void* tmp, size_t tmp_bytes, cudaStream_t st);
void emprep_rows(const EmPrep& p, const uint32_t* is_multi, uint32_t* ckey, uint32_t* cval, uint32_t* ckey_out, uint32_t* rlen,
                 void* tmp, size_t tmp_bytes, cudaStream_t st);
void emprep_fill_table(const DevDict& dd, const EmPrep& p, cudaStream_t st);
void emprep_fill(const DevDict& dd, const EmPrep& p, uint32_t nnz, unsigned long long* sort_keys_out, uint32_t* sort_vals_out,
                 void* tmp, size_t tmp_bytes, unsigned long long* stats2, cudaStream_t st);

// ---- BUS (kernels_bus.cu) ----
struct BusRecord {      // BUSData, src/BUSData.h:30-38
  uint64_t barcode, umi;
  int32_t ec;
  uint32_t count, flags, pad;
};
struct BusSpec {        // BUSOptions (src/common.h:38-91): where barcode / UMI / sequence sit in the files of a read set
  int nfiles;
  int n_bc, n_umi;
  int bc_f[4], bc_a[4], bc_b[4];
  int umi_f[4], umi_a[4], umi_b[4];
  int seq_file, seq_start;
  int num_flag;         // --num: flags = read number
  int paired;           // busopt.paired: two sequence reads, pseudoaligned as a pair (src/ProcessReads.cpp:1550-1567)
  int seq2_file, seq2_start;
  int no_umi;           // umi[0].fileno == -1 ("bulk_like", :1393): UMI field = ~0, one count in umi_len[1]
  unsigned long long fake_bc;   // n_bc == 0: the barcode every record gets (0 = 16 x 'A'; batch mode: the sample's id, :1603-1607)
  int tag_len;                  // --tag: length of the tag sequence that precedes the UMI (0 = none); umi_a[0] is already advanced by it
  unsigned long long tag_bin;   // stringToBinary(tag)
};
struct BusArgs {
  const uint8_t* bases[4];
  const uint32_t* off[4];
  uint32_t n_sets;
  uint64_t set_base;
  BusSpec spec;
  uint64_t* barcode;
  uint64_t* umi;
  uint32_t* flags;
  uint8_t* skip;
  uint8_t* notag;       // tag runs: 1 = the read set does not carry the tag
  uint32_t* bc_hist;    // 33 bins
  uint32_t* umi_hist;   // 33 bins
  unsigned long long* n_valid;
};
size_t bus_scan_bytes(uint32_t n);
void launch_bus_fields(const BusArgs& a, cudaStream_t st);
void launch_bus_records(const DevDict& dd, const int32_t* handle, uint32_t n, uint64_t base, uint32_t next_id,
                        int32_t* id_of, uint32_t* is_new, uint32_t* new_rank, uint32_t* is_mapped, uint32_t* rank,
                        const uint64_t* barcode, const uint64_t* umi, const uint32_t* flags, BusRecord* out, void* tmp,
                        size_t tmp_bytes, cudaStream_t st);

struct ResampleArgs {
  const double* cp;          // n_ec cumulative probabilities (discrete_distribution::_M_cp)
  uint32_t n_ec;
  uint64_t n_draws;          // N = sum(counts)
  int nb;
  const uint32_t* x0;        // nb initial minstd_rand0 states
  uint32_t* samp;            // nb x n_ec output counts (zeroed by the launcher)
};
void launch_resample(const ResampleArgs& a, cudaStream_t st);

}  // namespace kb
