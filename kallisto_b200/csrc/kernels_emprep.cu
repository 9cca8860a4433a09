// Device-side construction of the EM problem straight from the run's set dictionary: no EC table
// ever travels to the host on the quant path.
//
// Replaces MasterProcessor::update's id assignment (EC ids = order of first occurrence, what the
// reference produces with -t 1; src/ProcessReads.cpp:323-334,424-483) and calc_weights
// (src/weights.cpp:220-246), and lays the equivalence classes out twice: CSR by EC for the
// denominator pass, CSC by transcript (entries in increasing EC id) for the numerator pass of
// em_kernel.  Sorting / scanning uses CUB device primitives (library plumbing, not a hot path:
// ~1e6 keys once per run); the gather / weight / transpose kernels are ours.
#include <cub/cub.cuh>

#include "kb_device.cuh"
#include "kernels.hpp"

namespace kb {

namespace {

__global__ void gather_used_kernel(DevDict dd, const uint32_t* used, uint32_t n, unsigned long long* first, uint32_t* idx) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  first[i] = dd.first[used[i]];
  idx[i] = i;
}

// After the sort: EC id e <- used[order[e]]
__global__ void ec_meta_kernel(DevDict dd, const uint32_t* used, const uint32_t* order, uint32_t n, uint32_t* handle,
                               uint32_t* count, uint32_t* len, uint32_t* multi_len, uint32_t* is_multi, uint32_t* minkey) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint32_t h = used[order[e]];
  const unsigned long long word = dd.dslots[h];
  const uint32_t l = (uint32_t)((word >> 32) & 0xFFFFFFu);
  handle[e] = h;
  count[e] = dd.count[h];
  len[e] = l;
  multi_len[e] = l > 1 ? l : 0;
  is_multi[e] = l > 1 ? 1u : 0u;
  minkey[e] = l > 0 ? dd.pool[(uint32_t)word] : 0u;     // smallest transcript id of the set (lists are sorted)
}

// Rows of the EM matrices = the multi-transcript ECs, laid out in order of their SMALLEST TRANSCRIPT ID instead of EC
// id: ECs of one gene (adjacent transcript ids) become adjacent rows, so the alpha gathers of neighbouring rows and
// the norm gathers of neighbouring transcripts fall into the same 32-byte sectors.  The arithmetic does not change:
// a row is still accumulated in its own order and a transcript's entries stay in increasing EC id.
__global__ void multi_compact_kernel(const uint32_t* is_multi, const uint32_t* multi_index, const uint32_t* minkey, uint32_t n,
                                     uint32_t* ckey, uint32_t* cval) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || !is_multi[e]) return;
  const uint32_t r0 = multi_index[e];
  ckey[r0] = minkey[e];
  cval[r0] = e;
}
__global__ void row_len_kernel(const uint32_t* multi_ec, const uint32_t* len, uint32_t n_multi, uint32_t* rlen, uint32_t* rowpos) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_multi) return;
  if (r == n_multi) { rlen[r] = 0; return; }     // the scan runs over n_multi + 1 items
  const uint32_t e = multi_ec[r];
  rlen[r] = len[e];
  rowpos[e] = r;
}

// One warp per EC: copy its transcript ids, compute the weights, count transcript degrees.
__global__ void ec_fill_kernel(DevDict dd, EmPrep p) {
  const uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (e >= p.n_ec) return;
  const uint32_t h = p.handle[e];
  const uint32_t off = (uint32_t)dd.dslots[h];
  const uint32_t l = p.len[e];
  const uint32_t* src = dd.pool + off;
  // the EC table itself (ids in order of first occurrence)
  const uint32_t eo = p.ec_off[e];
  for (uint32_t j = lane; j < l; j += 32) p.ec_tid[eo + j] = src[j];
  if (l == 1) {
    if (lane == 0) p.t_single[src[0]] = (int32_t)e;
    return;
  }
  const uint32_t r = p.multi_index[e];     // row of this EC (rows are ordered by smallest transcript id)
  const uint32_t mo = p.m_rowoff[r];
  const double c = (double)p.count[e];
  for (uint32_t j = lane; j < l; j += 32) {
    const uint32_t t = src[j];
    p.m_tid[mo + j] = t;
    p.m_w[mo + j] = __ddiv_rn(c, p.eff[t]);        // calc_weights: counts[ec] / eff_lens[tr]
    p.m_row[mo + j] = r;
    p.m_iota[mo + j] = mo + j;
    p.k64_in[mo + j] = ((unsigned long long)t << 32) | e;    // CSC order: transcript, then EC id
    atomicAdd(&p.t_deg[t], 1u);
  }
}

// EC table only (export to other ranks): one warp per EC copies its transcript ids
__global__ void ec_table_kernel(DevDict dd, EmPrep p) {
  const uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (e >= p.n_ec) return;
  const uint32_t* src = dd.pool + (uint32_t)dd.dslots[p.handle[e]];
  const uint32_t l = p.len[e], eo = p.ec_off[e];
  for (uint32_t j = lane; j < l; j += 32) p.ec_tid[eo + j] = src[j];
}

// CSC entries in (transcript, EC id) order from the stable sort of (tid, entry index)
__global__ void csc_fill_kernel(EmPrep p, const uint32_t* sorted_entry, uint32_t nnz) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  const uint32_t j = sorted_entry[i];
  p.t_midx[i] = p.m_row[j];
  p.t_w[i] = p.m_w[j];
}

__global__ void stats_kernel(const uint32_t* count, const uint32_t* len, uint32_t n, unsigned long long* out) {
  unsigned long long a = 0, u = 0;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    a += count[e];
    if (len[e] == 1) u += count[e];
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
    u += __shfl_xor_sync(0xFFFFFFFFu, u, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out[0], a);
    atomicAdd(&out[1], u);
  }
}

}  // namespace

size_t emprep_sort_bytes(uint32_t n_used, uint32_t nnz_max) {
  size_t a = 0, b = 0, c = 0, d = 0;
  const int big = (int)std::max(n_used, nnz_max);
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, big);
  cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, big);
  cub::DeviceScan::ExclusiveSum(nullptr, c, (const uint32_t*)nullptr, (uint32_t*)nullptr, big + 1);
  (void)d;
  return std::max(a, std::max(b, c)) + 256;
}

void emprep_sort_by_first(const DevDict& dd, const uint32_t* used, uint32_t n_used, unsigned long long* key_in,
                          unsigned long long* key_out, uint32_t* idx_in, uint32_t* order_out, void* tmp, size_t tmp_bytes,
                          cudaStream_t st) {
  if (n_used == 0) return;
  gather_used_kernel<<<(n_used + 255) / 256, 256, 0, st>>>(dd, used, n_used, key_in, idx_in);
  cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_in, key_out, idx_in, order_out, (int)n_used, 0, 64, st);
}

void emprep_meta(const DevDict& dd, const uint32_t* used, const uint32_t* order, uint32_t n, const EmPrep& p,
                 uint32_t* multi_len, uint32_t* is_multi, void* tmp, size_t tmp_bytes, cudaStream_t st) {
  if (n == 0) return;
  ec_meta_kernel<<<(n + 255) / 256, 256, 0, st>>>(dd, used, order, n, p.handle, p.count, p.len, multi_len, is_multi, p.minkey);
  // n + 1 items so that the totals land in [n]
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, p.len, p.ec_off, (int)n + 1, st);
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, multi_len, p.m_off, (int)n + 1, st);
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, is_multi, p.multi_index, (int)n + 1, st);
}

// Row order of the EM matrices: multi-transcript ECs sorted by their smallest transcript id (stable: ties in EC id
// order).  Fills multi_ec (row -> EC id), m_rowoff (n_multi + 1) and overwrites multi_index (EC id -> row).
void emprep_rows(const EmPrep& p, const uint32_t* is_multi, uint32_t* ckey, uint32_t* cval, uint32_t* ckey_out, uint32_t* rlen,
                 void* tmp, size_t tmp_bytes, cudaStream_t st) {
  if (p.n_multi == 0) return;
  multi_compact_kernel<<<(p.n_ec + 255) / 256, 256, 0, st>>>(is_multi, p.multi_index, p.minkey, p.n_ec, ckey, cval);
  int bits = 1;
  while ((1u << bits) < p.n_targets && bits < 32) ++bits;
  cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, ckey, ckey_out, cval, p.multi_ec, (int)p.n_multi, 0, bits, st);
  row_len_kernel<<<(p.n_multi + 1 + 255) / 256, 256, 0, st>>>(p.multi_ec, p.len, p.n_multi, rlen, p.multi_index);
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, rlen, p.m_rowoff, (int)p.n_multi + 1, st);
}

void emprep_fill_table(const DevDict& dd, const EmPrep& p, cudaStream_t st) {
  if (p.n_ec == 0) return;
  const uint64_t threads = (uint64_t)p.n_ec * 32;
  ec_table_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(dd, p);
}

void emprep_fill(const DevDict& dd, const EmPrep& p, uint32_t nnz, unsigned long long* sort_keys_out, uint32_t* sort_vals_out,
                 void* tmp, size_t tmp_bytes, unsigned long long* stats2, cudaStream_t st) {
  if (p.n_ec == 0) return;
  const uint64_t threads = (uint64_t)p.n_ec * 32;
  ec_fill_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(dd, p);
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, p.t_deg, p.t_off, (int)p.n_targets + 1, st);
  if (nnz) {
    int bits = 1;
    while ((1u << bits) < p.n_targets && bits < 32) ++bits;
    // (transcript, EC id) order: a transcript's entries are accumulated in increasing EC id (EMAlgorithm.h:125-169
    // walks the ECs in id order), whatever the row order of the matrices
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, p.k64_in, sort_keys_out, p.m_iota, sort_vals_out, (int)nnz, 0, 32 + bits, st);
    csc_fill_kernel<<<(nnz + 255) / 256, 256, 0, st>>>(p, sort_vals_out, nnz);
  }
  cudaMemsetAsync(stats2, 0, 16, st);
  stats_kernel<<<device_sm_count(), 256, 0, st>>>(p.count, p.len, p.n_ec, stats2);
}

}  // namespace kb
