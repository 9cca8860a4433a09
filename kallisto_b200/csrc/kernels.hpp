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
  uint32_t fixed_len;
  uint32_t n_frag;          // pairs (paired) or reads (single)
  int paired;
  int strand_mode;          // 0 unstranded, 1 FR (--fr-stranded), 2 RF (--rf-stranded)
  uint64_t frag_base;       // global index of fragment 0
  int32_t* handle_out;      // per fragment: set handle, or KB_H_UNMAPPED
  uint16_t* tl_out;         // per fragment fragment-length candidate (mapPair), or nullptr
  uint32_t* q_count;        // resolve queue
  uint32_t* q_entries;      // stride KB_Q_STRIDE
  uint32_t bwords, iwords;  // shared-memory words per read (2-bit bases / invalid mask)
  const uint32_t* packed;   // pack_kernel output: per read, nb 64-bit base words then nb 32-bit invalid masks
  uint32_t nb;              // 32-base words per packed read = ceil(max_read_len / 32)
  uint32_t pstride;         // 32-bit words per packed read (multiple of 8 = 32 bytes)
  uint32_t empty_ec;        // handle of the empty index EC set, or 0xFFFFFFFF
  int refill_min;           // finished lanes of a warp that trigger a finalise + refill round
};
static constexpr int KB_Q_STRIDE = 2 + KB_MAX_E + 2;

struct ResolveArgs {
  uint32_t* scratch;        // per warp: scratch_stride entries
  uint32_t scratch_stride;
  uint32_t n_warps;
};

void launch_fill_u64(unsigned long long* p, uint64_t n, unsigned long long v, cudaStream_t st);
void launch_fill_i32(int32_t* p, uint64_t n, int32_t v, cudaStream_t st);
void launch_fill_memo2(Memo2Entry* p, uint64_t n, cudaStream_t st);
void launch_build_table(const TableBuildArgs& a, cudaStream_t st);
void launch_dict_init(const DictInitArgs& a, cudaStream_t st);

// Pseudoalignment of one batch: match kernel (thread per fragment) + resolve kernel (warp per
// queued fragment) [+ fragment-length finalisation].
// ev (optional): three events recorded before match_kernel, between the kernels, after resolve_kernel.
void launch_pseudoalign(const DevIndex& ix, const DevDict& dd, const BatchArgs& ba, const ResolveArgs& ra,
                        int threads_per_block, cudaStream_t st, cudaEvent_t* ev = nullptr);
void launch_fld_finalize(const DevDict& dd, const BatchArgs& ba, cudaStream_t st);
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
  int* state;                   // nb: 0 running, 1 final round pending, 2 done
  unsigned int* chcount;        // nb x 2 (double-buffered) change counters
  unsigned int* barrier;        // grid barrier words
  int max_iter, min_rounds;
};
int em_max_blocks(int threads_per_block);
void launch_em(const EmProblem& p, int threads_per_block, cudaStream_t st);

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
