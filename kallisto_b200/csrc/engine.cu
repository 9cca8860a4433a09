// Host-side engine: see engine.hpp.
#include "engine.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <numeric>
#include <random>
#include <thread>

namespace kb {

namespace {

inline void ck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw Error(std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}
#define KB_CK(x) ck((x), #x)

inline uint64_t pow2_ge(uint64_t v) {
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

template <class T> void DBuf<T>::alloc(size_t count) {
  release();
  n = count;
  if (count) KB_CK(cudaMalloc((void**)&p, count * sizeof(T)));
}
template <class T> void DBuf<T>::release() {
  if (p) cudaFree(p);
  p = nullptr;
  n = 0;
}
template <class T> void DBuf<T>::upload(const T* src, size_t count, cudaStream_t st) {
  if (count > n) alloc(count);
  if (count) KB_CK(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
}
template <class T> void DBuf<T>::download(T* dst, size_t count, size_t offset, cudaStream_t st) const {
  if (count) KB_CK(cudaMemcpyAsync(dst, p + offset, count * sizeof(T), cudaMemcpyDeviceToHost, st));
}
template <class T> void DBuf<T>::zero(cudaStream_t st) {
  if (n) KB_CK(cudaMemsetAsync(p, 0, n * sizeof(T), st));
}


// ------------------------------------------------------------------------------------------
// Index
// ------------------------------------------------------------------------------------------
Index::~Index() { delete shared_emws; }

std::unique_ptr<Index> Index::load(const std::string& path, int device, bool load_positions, int threads) {
  std::unique_ptr<Index> ix(new Index());
  ix->device = device;
  const double t0 = now_s();
  // the file is parsed on its own thread while this one initialises the driver and brings the CUDA context up
  // (0.3 s + 0.3-0.5 s in a fresh process)
  std::exception_ptr parse_err;
  std::thread parser([&] {
    try {
      load_index_v13(path, ix->flat, load_positions, threads);
    } catch (...) {
      parse_err = std::current_exception();
    }
  });
  int ndev = 0;
  const bool have_dev = cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0;
  if (!have_dev || device < 0 || device >= ndev) {
    parser.join();
    if (!have_dev) throw Error("kallisto_b200: no CUDA device available (this build has no CPU path)");
    throw Error("kallisto_b200: invalid CUDA device ordinal");
  }
  cudaError_t init_err = cudaSetDevice(device);
  if (init_err == cudaSuccess) init_err = cudaFree(0);
  if (init_err == cudaSuccess) {
    // The k-mer table probes touch one random 32-byte sector each.  With the default L2 fetch granularity
    // every miss pulls 64-128 bytes from HBM (ncu: 5.2 GB per 2 M pairs against 1.5 GB algorithmic); with
    // 32 bytes the DRAM traffic equals the algorithmic bytes (1.85 GB) at the same probe rate -- the rate is
    // bounded by the random-sector throughput of the L2-miss path (tools/randbench), not by bytes.
    size_t gran = 32;
    if (const char* s = getenv("KB_L2_FETCH")) gran = (size_t)atoi(s);    // 0: leave the device default
    if (gran > 0 && cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran) != cudaSuccess) cudaGetLastError();
  }
  parser.join();
  if (parse_err) std::rethrow_exception(parse_err);
  KB_CK(init_err);
  const double t1 = now_s();
  ix->load_seconds = t1 - t0;
  FlatIndex& f = ix->flat;
  if (f.graphless) {      // index.saved: targets only; nothing to build on the device, only quant-tcc can use it
    ix->build_seconds = 0;
    return ix;
  }
  if (f.onlist.size() != f.target_len.size())
    throw Error("kallisto_b200: indices whose on-list does not cover every target are not supported yet");
  if (f.ec_tid.size() >= 0xFFFFFFFFull) throw Error("kallisto_b200: index EC sets exceed 2^32 entries");
  if (f.num_targets() >= (1u << 24)) throw Error("kallisto_b200: more than 2^24 targets are not supported");

  const uint32_t nU = f.n_unitigs();
  std::vector<uint64_t> kstart(nU + 1, 0);
  for (uint32_t u = 0; u < nU; ++u) kstart[u + 1] = kstart[u] + (f.ulen[u] - f.k + 1);
  if (kstart[nU] != f.n_kmers) throw Error("kallisto_b200: k-mer count mismatch");

  cudaStream_t st = 0;
  // permanent arrays
  std::vector<uint32_t> ec_off32(f.ec_off.size());
  for (size_t i = 0; i < f.ec_off.size(); ++i) ec_off32[i] = (uint32_t)f.ec_off[i];
  ix->ec_off.upload(ec_off32.data(), ec_off32.size(), st);
  ix->n_index_tids = (uint32_t)f.ec_tid.size();
  ix->index_pool.alloc(std::max<size_t>(1, f.ec_tid.size()));
  ix->index_pool.upload(f.ec_tid.data(), f.ec_tid.size(), st);
  ix->blk_strand_off.upload(f.blk_strand_off.data(), f.blk_strand_off.size(), st);
  ix->strand.alloc(std::max<size_t>(1, f.strand.size()));
  ix->strand.upload(f.strand.data(), f.strand.size(), st);
  if (f.has_positions) {
    ix->fp_info.alloc(std::max<size_t>(1, f.fp_info.size() / 4));
    if (!f.fp_info.empty())
      KB_CK(cudaMemcpyAsync(ix->fp_info.p, f.fp_info.data(), f.fp_info.size() * 4, cudaMemcpyHostToDevice, st));
    std::vector<uint32_t> busize(f.blk_lb.size());
    for (uint32_t u = 0; u < f.n_unitigs(); ++u)
      for (uint64_t b = f.blk_off[u]; b < f.blk_off[u + 1]; ++b) busize[b] = f.ulen[u];
    ix->blk_usize.upload(busize.data(), busize.size(), st);
    ix->target_len.upload(f.target_len.data(), f.target_len.size(), st);
    KB_CK(cudaStreamSynchronize(st));
  }
  for (uint32_t e = 0; e < f.n_ec(); ++e) {
    const uint32_t len = (uint32_t)(f.ec_off[e + 1] - f.ec_off[e]);
    if (len == 0) ix->empty_ec = e;
    ix->max_set_len = std::max(ix->max_set_len, len);
  }

  cudaStream_t st0 = st;
  // set dictionary, initial state: the index's own EC sets.  Done first: the k-mer slots store the
  // dictionary handle of their block's set, not its index-local id.
  ix->dict_cap = pow2_ge((uint64_t)f.n_ec() * 4 + (1u << 20));
  ix->dslots_init.alloc(ix->dict_cap);
  launch_fill_u64(ix->dslots_init.p, ix->dict_cap, ~0ULL, st0);
  ix->ec_handle.alloc(std::max<uint32_t>(1, f.n_ec()));
  {
    DictInitArgs a{};
    a.ec_off = ix->ec_off.p; a.pool = ix->index_pool.p; a.n_ec = f.n_ec();
    a.dslots = ix->dslots_init.p; a.dmask = ix->dict_cap - 1; a.ec_handle = ix->ec_handle.p;
    launch_dict_init(a, st0);
    KB_CK(cudaGetLastError());
  }
  ix->h_ec_handle.resize(f.n_ec());
  ix->ec_handle.download(ix->h_ec_handle.data(), f.n_ec(), 0, st0);
  KB_CK(cudaStreamSynchronize(st0));
  {
    std::vector<uint32_t> blk_handle(f.blk_ec.size());
    for (size_t i = 0; i < blk_handle.size(); ++i) blk_handle[i] = (uint32_t)ix->h_ec_handle[f.blk_ec[i]];
    ix->blk_ec.upload(blk_handle.data(), blk_handle.size(), st0);
    KB_CK(cudaStreamSynchronize(st0));
  }
  if (ix->empty_ec != 0xFFFFFFFFu) ix->empty_ec = (uint32_t)ix->h_ec_handle[ix->empty_ec];

  // k-mer table.  Every probe costs one random 32-byte sector whatever the table size, and the sector rate of the
  // L2-miss path is the bound of match_kernel (profiles/README.md), so HBM capacity is traded for shorter probe
  // sequences: slots >= 4 x k-mers (load 0.14-0.27: 1.14 visits per lookup; human: 34 GB of the 180 GB) when that
  // leaves three quarters of the free device memory to the run, else 2 x (load <= 0.5; 1.37 visits measured at
  // 0.27).  KB_TABLE_FACTOR overrides.
  double factor = 4.0;
  {
    size_t free_b = 0, total_b = 0;
    KB_CK(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t cap4 = pow2_ge(std::max<uint64_t>(1024, f.n_kmers * 4));
    if (cap4 * sizeof(KmerSlot) > free_b / 4) factor = 2.0;
  }
  if (const char* s = getenv("KB_TABLE_FACTOR")) { const double v = atof(s); if (v >= 1.25 && v <= 64.0) factor = v; }
  ix->table_cap = pow2_ge(std::max<uint64_t>(1024, (uint64_t)((double)f.n_kmers * factor)));
  ix->slots.alloc(ix->table_cap);
  // presence filter: 2^KB_FILTER_LOG2 bits (default: about 3.5 bits per k-mer, at most 2^29 bits = 64 MB so that it
  // fits the persisting part of the 126 MB L2; 0 = off).  Single hash: a miss passes it with probability
  // 1 - exp(-n / bits).
  uint32_t filter_bits = 0;
  {
    int lg = 0;
    while ((1ull << lg) < f.n_kmers * 3 && lg < 29) ++lg;
    if (lg < 16) lg = 16;
    if (const char* s = getenv("KB_FILTER_LOG2")) lg = atoi(s);
    if (lg >= 10 && lg <= 32) filter_bits = lg == 32 ? 0 : (1u << lg);
    if (lg == 32) filter_bits = 0;
  }
  if (filter_bits) {
    ix->filter.alloc(filter_bits / 32);
    ix->filter.zero(st);
  }
  DBuf<int> err;
  err.alloc(1);
  err.zero(st);
  {
    DBuf<uint8_t> d_useq;
    DBuf<uint64_t> d_byteoff, d_skmer, d_kstart, d_blkoff;
    DBuf<uint32_t> d_lb, d_ub;
    d_useq.upload(f.useq.data(), f.useq.size(), st);
    d_byteoff.upload(f.useq_byteoff.data(), f.useq_byteoff.size(), st);
    d_skmer.alloc(std::max<size_t>(1, f.skmer.size()));
    d_skmer.upload(f.skmer.data(), f.skmer.size(), st);
    d_kstart.upload(kstart.data(), kstart.size(), st);
    d_blkoff.upload(f.blk_off.data(), f.blk_off.size(), st);
    d_lb.upload(f.blk_lb.data(), f.blk_lb.size(), st);
    d_ub.upload(f.blk_ub.data(), f.blk_ub.size(), st);
    TableBuildArgs a{};
    a.useq = d_useq.p; a.useq_byteoff = d_byteoff.p; a.skmer = d_skmer.p; a.kstart = d_kstart.p;
    a.blk_off = d_blkoff.p; a.blk_lb = d_lb.p; a.blk_ub = d_ub.p; a.blk_ec = ix->blk_ec.p;
    a.n_long = f.n_long; a.n_unitigs = nU; a.k = f.k; a.n_kmers = f.n_kmers;
    a.slots = ix->slots.p; a.mask = ix->table_cap - 1; a.error = err.p;
    a.filter = filter_bits ? ix->filter.p : nullptr; a.filter_mask = filter_bits ? filter_bits - 1 : 0;
    launch_build_table(a, st);
    KB_CK(cudaGetLastError());
    KB_CK(cudaStreamSynchronize(st));
  }
  int herr = 0;
  err.download(&herr, 1, 0, st);
  KB_CK(cudaStreamSynchronize(st));
  if (herr & KB_DEVERR_TABLE_DUP) throw Error("kallisto_b200: corrupt index (a k-mer occurs in two unitigs)");

  DevIndex& d = ix->dev;
  d.slots = ix->slots.p;
  d.mask = ix->table_cap - 1;
  d.filter = filter_bits ? ix->filter.p : nullptr;
  d.filter_mask = filter_bits ? filter_bits - 1 : 0;
  if (filter_bits) {
    // keep the filter in L2: persisting carve-out as large as the device allows (the access window itself is set on the
    // stream of every run, Quant::apply_l2_window)
    int max_persist = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
    size_t want = std::min<size_t>((size_t)filter_bits / 8, (size_t)std::max(0, max_persist));
    if (const char* s = getenv("KB_L2_PERSIST_MB")) want = std::min<size_t>((size_t)atoll(s) << 20, (size_t)std::max(0, max_persist));
    if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) ix->l2_persist_bytes = want;
    else cudaGetLastError();
  }
  d.k = f.k;
  d.n_ec = f.n_ec();
  d.n_targets = f.num_targets();
  d.ec_off = ix->ec_off.p;
  d.ec_handle = ix->ec_handle.p;
  d.blk_ec = ix->blk_ec.p;
  d.blk_strand_off = ix->blk_strand_off.p;
  d.strand = ix->strand.p;
  d.fp_info = f.has_positions ? ix->fp_info.p : nullptr;
  d.blk_usize = f.has_positions ? ix->blk_usize.p : nullptr;
  d.target_len = f.has_positions ? ix->target_len.p : nullptr;
  if (f.dlist_n) {
    // D-list k-mers: open-addressing set, built on the host (the list is a small fraction of the k-mer table)
    const uint64_t cap = pow2_ge(2 * f.dlist_n + 16);
    std::vector<unsigned long long> tab(cap, ~0ULL);
    for (uint64_t km : f.dlist) {
      uint64_t h = kb_mix64(km) & (cap - 1);
      while (tab[h] != ~0ULL && tab[h] != km) h = (h + 1) & (cap - 1);
      tab[h] = km;
    }
    ix->dfk.upload(tab.data(), cap, st);
    KB_CK(cudaStreamSynchronize(st));
    d.dfk = ix->dfk.p;
    d.dfk_mask = cap - 1;
  }
  ix->build_seconds = now_s() - t1;
  return ix;
}

// ------------------------------------------------------------------------------------------
// Quant
// ------------------------------------------------------------------------------------------
Quant::Quant(Index& ix, const QuantOptions& opt) : ix_(ix), opt_(opt), flens_(1000, 0) {
  if (ix.flat.graphless) throw Error("kallisto_b200: this index has no k-mers (an index.saved written by `bus`): only quant-tcc can use it");
  if (!ix_.ws_in_use) {   // borrow the index's work buffers
    ix_.ws_in_use = true;
    if (!ix_.shared_emws) ix_.shared_emws = new EmWs();
    bws_ = &ix_.shared_bws;
    emws_ = ix_.shared_emws;
  } else {
    own_ws_ = true;
    bws_ = new BatchWs();
    emws_ = new EmWs();
  }
  if (opt_.fp_fl >= 0 && !ix_.flat.has_positions)
    throw Error("kallisto_b200: the fragment-position filter needs an index loaded with positions (load_positions = 1)");
  if (const char* s = getenv("KB_REFILL_MIN")) opt_.refill_min = std::max(1, std::min(32, atoi(s)));   // tuning knob
  KB_CK(cudaSetDevice(ix_.device));
  KB_CK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  KB_CK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    KB_CK(cudaEventCreateWithFlags(&ev_copied_[i], cudaEventDisableTiming));
    KB_CK(cudaEventCreateWithFlags(&ev_done_[i], cudaEventDisableTiming));
  }
  cudaStream_t st = stream_;
  apply_l2_window();
  const uint64_t nE = ix_.flat.n_ec();
  // pools and tables of this run
  const uint64_t pool_cap =
      std::min<uint64_t>(0xFFFFFFF0ull, (uint64_t)ix_.n_index_tids + std::max<uint64_t>(1u << 24, 4ull * ix_.n_index_tids));
  pool_.alloc(pool_cap);
  KB_CK(cudaMemcpyAsync(pool_.p, ix_.index_pool.p, (size_t)ix_.n_index_tids * 4, cudaMemcpyDeviceToDevice, st));
  dslots_.alloc(ix_.dict_cap);
  KB_CK(cudaMemcpyAsync(dslots_.p, ix_.dslots_init.p, ix_.dict_cap * 8, cudaMemcpyDeviceToDevice, st));
  count_.alloc(ix_.dict_cap);
  count_.zero(st);
  first_.alloc(ix_.dict_cap);
  launch_fill_u64(first_.p, ix_.dict_cap, ~0ULL, st);
  const uint64_t m2_cap = pow2_ge(nE * 8 + (1u << 20));
  const uint64_t mn_cap = pow2_ge(nE * 4 + (1u << 20));
  m2_.alloc(m2_cap);
  launch_fill_memo2(m2_.p, m2_cap, st);
  mn_key_.alloc(mn_cap);
  launch_fill_u64(mn_key_.p, mn_cap, ~0ULL, st);
  mn_val_.alloc(mn_cap);
  launch_fill_i32(mn_val_.p, mn_cap, KB_H_NOTREADY, st);
  const uint64_t tpool_cap = mn_cap * 8;
  tpool_.alloc(tpool_cap);
  // pool_top, tpool_top and the statistics live in separate 128-byte lines: atomics on one line are served one after
  // the other by the L2, and resolve_kernel allocates from both pools once per fragment (with the counters side by
  // side, and a statistics increment per fragment on the same line, those atomics were 30 % of the kernel's time)
  counters_.alloc(48);
  {
    unsigned long long init[48] = {};
    init[0] = ix_.n_index_tids;
    KB_CK(cudaMemcpyAsync(counters_.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    KB_CK(cudaStreamSynchronize(st));   // init is a stack array
  }
  error_.alloc(1);
  error_.zero(st);

  dd_.pool = pool_.p; dd_.pool_top = counters_.p + 0; dd_.pool_cap = pool_cap;
  dd_.dslots = dslots_.p; dd_.dmask = ix_.dict_cap - 1;
  dd_.count = count_.p; dd_.first = first_.p;
  dd_.m2 = m2_.p; dd_.m2_mask = m2_cap - 1;
  dd_.mn_key = mn_key_.p; dd_.mn_val = mn_val_.p; dd_.mn_mask = mn_cap - 1;
  dd_.tpool = tpool_.p; dd_.tpool_top = counters_.p + 16; dd_.tpool_cap = tpool_cap;
  dd_.error = error_.p; dd_.stats = counters_.p + 32;

  // batch staging
  const uint32_t max_frag = opt_.max_batch_reads;
  if (bws_->d_handles.n < max_frag) bws_->d_handles.alloc(max_frag);
  if (bws_->d_tl.n < max_frag) bws_->d_tl.alloc(max_frag);
  if (bws_->d_qcount.n < 1) bws_->d_qcount.alloc(1);
  if (bws_->d_qentries.n < (size_t)max_frag * KB_Q_STRIDE) bws_->d_qentries.alloc((size_t)max_frag * KB_Q_STRIDE);
  // rare path: fragments with more than KB_MAX_E distinct EC sets (spill area per resident lane + wide queue)
  if (bws_->d_qbig_count.n < 1) bws_->d_qbig_count.alloc(1);
  if (bws_->d_qbig.n < (size_t)KB_QBIG_CAP * KB_QBIG_STRIDE) bws_->d_qbig.alloc((size_t)KB_QBIG_CAP * KB_QBIG_STRIDE);
  {
    int sms = 0, tpsm = 0;
    KB_CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix_.device));
    KB_CK(cudaDeviceGetAttribute(&tpsm, cudaDevAttrMaxThreadsPerMultiProcessor, ix_.device));
    const size_t lanes = (size_t)sms * (size_t)tpsm;
    if (bws_->d_spill.n < lanes * KB_SPILL) bws_->d_spill.alloc(lanes * KB_SPILL);
  }
  // resolve-kernel scratch: 2 x max_set_len words per lane group, at most ~1 GiB in total
  {
    const char* e = getenv("KB_RESOLVE_G");
    const int g = e ? atoi(e) : 32;
    resolve_group_ = (g == 4 || g == 8 || g == 16 || g == 32) ? (uint32_t)g : 32u;
  }
  const uint64_t stride = std::max<uint64_t>(64, 2ull * ix_.max_set_len);
  uint64_t warps = (1ull << 28) / stride;
  // the kernel is latency-bound: fill the SMs (32 warps each), every warp split into 32 / group lane groups
  warps = std::min<uint64_t>((uint64_t)device_sm_count() * 32 * (32 / resolve_group_), std::max<uint64_t>(64, warps));
  n_resolve_warps_ = (uint32_t)(warps / 16 * 16);
  if (bws_->d_scratch.n < (size_t)n_resolve_warps_ * stride) bws_->d_scratch.alloc((size_t)n_resolve_warps_ * stride);
  scratch_stride_ = (uint32_t)stride;
  // EM workspace: sized once per index for the EC tables runs on it normally end with (twice the index's own sets),
  // so that the timed EM tail of a run does not allocate; it still grows on demand
  if (!opt_.bus) reserve_em(2 * (size_t)nE + (1u << 16), 4 * (size_t)ix_.n_index_tids + (1u << 20));
  KB_CK(cudaStreamSynchronize(st));
}

void Quant::reserve_em(size_t n_ecs, size_t nnz) {
  KB_CK(cudaSetDevice(ix_.device));
  EmWs& w = *emws_;
  const uint32_t T = ix_.flat.num_targets();
  const size_t n1 = n_ecs + 1;
  auto g32 = [](DBuf<uint32_t>& b, size_t need) { if (b.n < need) b.alloc(need); };
  auto gd = [](DBuf<double>& b, size_t need) { if (b.n < need) b.alloc(need); };
  if (w.used.n < ix_.dict_cap) w.used.alloc(ix_.dict_cap);
  if (w.scal.n < 8) w.scal.alloc(8);
  if (w.key_in.n < n1) { w.key_in.alloc(n1); w.key_out.alloc(n1); }
  g32(w.idx_in, n1); g32(w.order, n1); g32(w.handle, n1); g32(w.count, n1); g32(w.len, n1);
  g32(w.multi_len, n1); g32(w.is_multi, n1); g32(w.ec_off, n1); g32(w.m_off, n1); g32(w.multi_index, n1);
  g32(w.minkey, n1); g32(w.ckey, n1); g32(w.cval, n1); g32(w.ckey_out, n1); g32(w.rlen, n1 + 1);
  const size_t tmp_need = emprep_sort_bytes((uint32_t)n1, (uint32_t)std::max<size_t>(nnz, (size_t)T + 1));
  if (w.tmp.n < tmp_need) w.tmp.alloc(tmp_need);
  g32(w.ec_tid, std::max<size_t>(1, nnz)); g32(w.multi_ec, n1); g32(w.m_rowoff, n1 + 1);
  const size_t nz = std::max<size_t>(1, nnz);
  g32(w.m_tid, nz); g32(w.m_row, nz); g32(w.m_iota, nz); g32(w.sortv, nz); g32(w.t_midx, nz);
  if (w.k64_in.n < nz) { w.k64_in.alloc(nz); w.k64_out.alloc(nz); }
  if (w.bar.n < 1) w.bar.alloc(1);
  g32(w.cnt_row, n1); gd(w.single_cnt, T);
  gd(w.m_w, nz); gd(w.t_w, nz);
  g32(w.t_deg, (size_t)T + 1); g32(w.t_off, (size_t)T + 1);
  if (w.t_single.n < T) w.t_single.alloc(T);
  gd(w.eff, T); gd(w.alpha, T); gd(w.norm, n1);
  if (w.emi.n < 8) w.emi.alloc(8);
  if (w.chcount.n < 2) w.chcount.alloc(2);
}

Quant::~Quant() {
  if (stream_) cudaStreamSynchronize(stream_);
  if (own_ws_) { delete emws_; delete bws_; } else { ix_.ws_in_use = false; }
  if (h_off_pinned_) cudaFreeHost(h_off_pinned_);
  for (auto ev : events_) cudaEventDestroy(ev);
  for (int i = 0; i < 2; ++i) {
    if (ev_copied_[i]) cudaEventDestroy(ev_copied_[i]);
    if (ev_done_[i]) cudaEventDestroy(ev_done_[i]);
  }
  if (copy_stream_) cudaStreamDestroy(copy_stream_);
  if (stream_ && own_stream_) cudaStreamDestroy(stream_);
}

void Quant::set_stream(cudaStream_t st) {
  KB_CK(cudaStreamSynchronize(stream_));
  if (stream_ && own_stream_) cudaStreamDestroy(stream_);
  stream_ = st;
  own_stream_ = false;
  apply_l2_window();
}

// Kernels launched on the run's stream treat the presence filter as persisting in L2; everything else streams.
void Quant::apply_l2_window() {
  if (!ix_.filter.p || ix_.l2_persist_bytes == 0) return;
  int max_win = 0;
  cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, ix_.device);
  const size_t bytes = std::min<size_t>(ix_.filter.n * 4, (size_t)std::max(0, max_win));
  if (bytes == 0) return;
  cudaStreamAttrValue v{};
  v.accessPolicyWindow.base_ptr = (void*)ix_.filter.p;
  v.accessPolicyWindow.num_bytes = bytes;
  v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)ix_.l2_persist_bytes / (double)bytes);
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(stream_, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
}

Quant::Timings Quant::timings() {
  KB_CK(cudaStreamSynchronize(stream_));
  for (size_t i = 0; i + 3 < events_.size(); i += 4) {
    float a = 0, b = 0, c = 0;
    KB_CK(cudaEventElapsedTime(&c, events_[i], events_[i + 1]));
    KB_CK(cudaEventElapsedTime(&a, events_[i + 1], events_[i + 2]));
    KB_CK(cudaEventElapsedTime(&b, events_[i + 2], events_[i + 3]));
    tacc_.pack_ms += c;
    tacc_.match_ms += a;
    tacc_.resolve_ms += b;
    ++tacc_.match_launches;
    ++tacc_.resolve_launches;
  }
  for (auto ev : events_) cudaEventDestroy(ev);
  events_.clear();
  return tacc_;
}

void Quant::sync() { KB_CK(cudaStreamSynchronize(stream_)); }

void Quant::check_device_errors() {
  int herr = 0;
  error_.download(&herr, 1, 0, stream_);
  KB_CK(cudaStreamSynchronize(stream_));
  if (herr == 0) return;
  std::string m = "kallisto_b200: device-side failure:";
  if (herr & KB_DEVERR_POOL_FULL) m += " set pool exhausted;";
  if (herr & KB_DEVERR_DICT_FULL) m += " set dictionary full;";
  if (herr & KB_DEVERR_MEMO_FULL) m += " memo table full;";
  if (herr & KB_DEVERR_TPOOL_FULL) m += " tuple pool exhausted;";
  if (herr & KB_DEVERR_E_OVERFLOW) m += " a fragment hit more than 128 distinct EC sets, or more than 65536 fragments of a batch hit more than 16;";
  throw Error(m);
}

void Quant::run_batch(const uint8_t* d_bases, const uint32_t* d_off, uint32_t n_reads, uint32_t fixed_len,
                      uint32_t max_read_len, const uint8_t* d_bases2, const uint32_t* d_off2) {
  const uint32_t n_frag = opt_.paired ? n_reads / 2 : n_reads;
  if (opt_.paired && (n_reads & 1)) throw Error("kallisto_b200: odd number of reads in a paired batch");
  if (n_frag > bws_->d_handles.n) throw Error("kallisto_b200: batch larger than max_batch_reads");
  if (n_frag == 0) return;
  ecs_valid_ = false;
  dev_stats_valid_ = false;
  dev_problem_valid_ = false;
  BatchArgs ba{};
  ba.bases = d_bases;
  ba.off = d_off;
  ba.bases2 = d_bases2;
  ba.off2 = d_off2;
  ba.fixed_len = fixed_len;
  ba.n_frag = n_frag;
  ba.paired = opt_.paired;
  ba.strand_mode = opt_.strand_mode;
  ba.frag_base = have_frag_base_ ? frag_base_ : n_frag_total_;
  have_frag_base_ = false;
  ba.handle_out = bws_->d_handles.p;
  const bool want_fld = opt_.paired && opt_.collect_fld && tlencount_ < 10000;   // ProcessReads.cpp:981-1017
  ba.tl_out = want_fld ? bws_->d_tl.p : nullptr;
  ba.q_count = bws_->d_qcount.p;
  ba.q_entries = bws_->d_qentries.p;
  ba.spill = bws_->d_spill.p;
  ba.qbig_count = bws_->d_qbig_count.p;
  ba.qbig_entries = bws_->d_qbig.p;
  ba.qbig_cap = KB_QBIG_CAP;
  ba.nb = std::max<uint32_t>(1, (max_read_len + 31) / 32);
  ba.pstride = (3 * ba.nb + 7) & ~7u;
  {
    const size_t need = (size_t)n_reads * ba.pstride;
    if ((uint64_t)n_reads * ba.nb >= (1ull << 32))      // pack_kernel / dlist_scan_kernel index (read, word) with 32 bits
      throw Error("kallisto_b200: batch too large for its longest read (reads x ceil(max length / 32) must be < 2^32)");
    if (bws_->d_packed.n < need) bws_->d_packed.alloc(std::max(need, (size_t)opt_.max_batch_reads * 2 * 16));
  }
  ba.packed = bws_->d_packed.p;
  ba.empty_ec = ix_.empty_ec;
  ba.refill_min = opt_.refill_min;
  ba.skip = cur_skip_;
  ba.skip_w = nullptr;
  if (ix_.dev.dfk) {
    // D-list: fragments holding a distinguishing flanking k-mer are marked by dlist_scan_kernel (after packing)
    if (cur_skip_) {
      ba.skip_w = const_cast<uint8_t*>(cur_skip_);          // bus: the scan adds to the bad-barcode marks
    } else {
      if (bws_->d_skip.n < n_frag) bws_->d_skip.alloc(std::max<size_t>(n_frag, opt_.max_batch_reads));
      KB_CK(cudaMemsetAsync(bws_->d_skip.p, 0, n_frag, stream_));
      ba.skip = ba.skip_w = bws_->d_skip.p;
    }
  }
  ba.fp_fl = opt_.fp_fl;
  ba.start = cur_start_;
  ba.start2 = cur_start2_;
  ba.notag = cur_notag_;
  ba.alt_start = cur_alt_start_;
  ba.alt_start2 = cur_alt_start2_;
  ResolveArgs ra{};
  ra.scratch = bws_->d_scratch.p;
  ra.scratch_stride = scratch_stride_;
  ra.n_warps = n_resolve_warps_;
  ra.group = resolve_group_;

  int tpb = opt_.threads_per_block;
  const size_t per_thread = (size_t)4 * (KB_MAX_E + 3 + 4 * ba.nb);   // match_kernel's shared memory per lane
  while (tpb > 32 && per_thread * tpb > 200 * 1024) tpb >>= 1;
  if (per_thread * tpb > 200 * 1024) throw Error("kallisto_b200: read too long for the short-read kernel");
  cudaEvent_t* ev = nullptr;
  if (timing_) {
    const size_t base = events_.size();
    events_.resize(base + 4);
    for (int i = 0; i < 4; ++i) KB_CK(cudaEventCreate(&events_[base + i]));
    ev = events_.data() + base;
  }
  launch_pseudoalign(ix_.dev, dd_, ba, ra, tpb, stream_, ev);
  KB_CK(cudaGetLastError());
  n_kernel_launches += 3 + (ba.skip_w ? 1 : 0);   // pack_kernel, [dlist_scan_kernel,] match_kernel, resolve_kernel
  if (want_fld) {
    launch_fld_finalize(dd_, ba, stream_);
    ++n_kernel_launches;
    h_tl_.resize(n_frag);
    bws_->d_tl.download(h_tl_.data(), n_frag, 0, stream_);
    KB_CK(cudaStreamSynchronize(stream_));
    // first (10000 - tlencount) qualifying fragments of this batch, in read order
    int goal = 10000 - (int)tlencount_;
    uint32_t local = 0;
    for (uint32_t i = 0; i < n_frag && goal > 0; ++i) {
      const uint16_t tl = h_tl_[i];
      if (tl > 0) { ++flens_[tl]; tl_list_.push_back(tl); --goal; ++local; }
    }
    tlencount_ += local;
  }
  n_frag_total_ += n_frag;
}

void Quant::pseudoalign_device(const uint8_t* d_bases, const uint32_t* d_off, uint32_t n_reads, uint32_t fixed_len,
                               uint32_t max_read_len) {
  KB_CK(cudaSetDevice(ix_.device));
  run_batch(d_bases, d_off, n_reads, fixed_len, max_read_len);
}

void Quant::pseudoalign_host(const char* bases, const uint32_t* off, uint32_t n_reads, uint32_t fixed_len,
                             int32_t* handles_out) {
  KB_CK(cudaSetDevice(ix_.device));
  if (n_reads == 0) return;
  uint64_t n_bases;
  uint32_t maxlen = fixed_len;
  if (off) {
    n_bases = off[n_reads];
    maxlen = 0;
    for (uint32_t i = 0; i < n_reads; ++i) maxlen = std::max(maxlen, off[i + 1] - off[i]);
    if (off[0] != 0) throw Error("kallisto_b200: offsets must start at 0");
  } else {
    n_bases = (uint64_t)n_reads * fixed_len;
  }
  // stage into buffer `s`; the copy runs on its own stream so that it overlaps the previous batch's kernels
  const int s = stage_idx_;
  stage_idx_ ^= 1;
  DBuf<uint8_t>& db = bws_->stage_b[s][0];
  DBuf<uint32_t>& dofs = bws_->stage_o[s][0];
  KB_CK(cudaStreamWaitEvent(copy_stream_, ev_done_[s], 0));      // kernels that last read this buffer
  if (db.n < n_bases + 16) { KB_CK(cudaStreamSynchronize(stream_)); db.alloc(std::max<uint64_t>(n_bases + 16, opt_.max_batch_bases)); }
  KB_CK(cudaMemcpyAsync(db.p, bases, n_bases, cudaMemcpyHostToDevice, copy_stream_));
  if (off) {
    if (dofs.n < (size_t)n_reads + 1) { KB_CK(cudaStreamSynchronize(stream_)); dofs.alloc(std::max<size_t>((size_t)n_reads + 1, (size_t)opt_.max_batch_reads * 2 + 1)); }
    KB_CK(cudaMemcpyAsync(dofs.p, off, ((size_t)n_reads + 1) * 4, cudaMemcpyHostToDevice, copy_stream_));
  }
  KB_CK(cudaEventRecord(ev_copied_[s], copy_stream_));
  KB_CK(cudaStreamWaitEvent(stream_, ev_copied_[s], 0));
  run_batch(db.p, off ? dofs.p : nullptr, n_reads, fixed_len, maxlen);
  KB_CK(cudaEventRecord(ev_done_[s], stream_));
  if (handles_out) {
    const uint32_t n_frag = opt_.paired ? n_reads / 2 : n_reads;
    bws_->d_handles.download(handles_out, n_frag, 0, stream_);
    KB_CK(cudaStreamSynchronize(stream_));
  } else {
    // the caller may reuse its buffers once the copy is done; the kernels keep running
    KB_CK(cudaEventSynchronize(ev_copied_[s]));
  }
}

void Quant::pseudoalign_host_pe(const char* bases1, const uint32_t* off1, const char* bases2, const uint32_t* off2,
                                uint32_t n_pairs, uint32_t fixed_len, int32_t* handles_out) {
  KB_CK(cudaSetDevice(ix_.device));
  if (n_pairs == 0) return;
  if (!opt_.paired) throw Error("kallisto_b200: per-mate buffers need a paired run");
  if ((off1 == nullptr) != (off2 == nullptr)) throw Error("kallisto_b200: give offsets for both mates or for neither");
  uint64_t nb1, nb2;
  uint32_t maxlen = fixed_len;
  if (off1) {
    nb1 = off1[n_pairs];
    nb2 = off2[n_pairs];
    maxlen = 0;
    for (uint32_t i = 0; i < n_pairs; ++i) maxlen = std::max(maxlen, std::max(off1[i + 1] - off1[i], off2[i + 1] - off2[i]));
  } else {
    nb1 = nb2 = (uint64_t)n_pairs * fixed_len;
  }
  const int s = stage_idx_;
  stage_idx_ ^= 1;
  DBuf<uint8_t>& b1 = bws_->stage_b[s][0];
  DBuf<uint8_t>& b2 = bws_->stage_b[s][1];
  DBuf<uint32_t>& o1 = bws_->stage_o[s][0];
  DBuf<uint32_t>& o2 = bws_->stage_o[s][1];
  KB_CK(cudaStreamWaitEvent(copy_stream_, ev_done_[s], 0));
  if (b1.n < nb1 + 16) { KB_CK(cudaStreamSynchronize(stream_)); b1.alloc(std::max<uint64_t>(nb1 + 16, opt_.max_batch_bases / 2 + 16)); }
  if (b2.n < nb2 + 16) { KB_CK(cudaStreamSynchronize(stream_)); b2.alloc(std::max<uint64_t>(nb2 + 16, opt_.max_batch_bases / 2 + 16)); }
  KB_CK(cudaMemcpyAsync(b1.p, bases1, nb1, cudaMemcpyHostToDevice, copy_stream_));
  KB_CK(cudaMemcpyAsync(b2.p, bases2, nb2, cudaMemcpyHostToDevice, copy_stream_));
  if (off1) {
    const size_t no = (size_t)n_pairs + 1;
    if (o1.n < no) { KB_CK(cudaStreamSynchronize(stream_)); o1.alloc(std::max<size_t>(no, (size_t)opt_.max_batch_reads + 1)); }
    if (o2.n < no) { KB_CK(cudaStreamSynchronize(stream_)); o2.alloc(std::max<size_t>(no, (size_t)opt_.max_batch_reads + 1)); }
    KB_CK(cudaMemcpyAsync(o1.p, off1, no * 4, cudaMemcpyHostToDevice, copy_stream_));
    KB_CK(cudaMemcpyAsync(o2.p, off2, no * 4, cudaMemcpyHostToDevice, copy_stream_));
  }
  KB_CK(cudaEventRecord(ev_copied_[s], copy_stream_));
  KB_CK(cudaStreamWaitEvent(stream_, ev_copied_[s], 0));
  run_batch(b1.p, off1 ? o1.p : nullptr, 2 * n_pairs, fixed_len, maxlen, b2.p, off1 ? o2.p : nullptr);
  KB_CK(cudaEventRecord(ev_done_[s], stream_));
  if (handles_out) {
    bws_->d_handles.download(handles_out, n_pairs, 0, stream_);
    KB_CK(cudaStreamSynchronize(stream_));
  } else {
    KB_CK(cudaEventSynchronize(ev_copied_[s]));
  }
}

void Quant::bus_batch_host(const char* const* bases, const uint32_t* const* offs, uint32_t n_sets, BusRecord* records_out,
                           uint32_t* n_records_out) {
  KB_CK(cudaSetDevice(ix_.device));
  if (!opt_.bus) throw Error("kallisto_b200: not a bus run");
  if (n_records_out) *n_records_out = 0;
  if (n_sets == 0) return;
  const BusSpec& sp = opt_.bus_spec;
  cudaStream_t st = stream_;
  const uint8_t* db[4] = {nullptr, nullptr, nullptr, nullptr};
  const uint32_t* dofs[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int k = 0; k < sp.nfiles; ++k) {
    if (!bases[k] || !offs[k]) throw Error("kallisto_b200: bus batch needs bases and offsets for every file of the technology");
    const uint64_t nbz = offs[k][n_sets];
    if (bus_b_[k].n < nbz + 16) bus_b_[k].alloc(std::max<uint64_t>(nbz + 16, opt_.max_batch_bases / 2 + 16));
    if (bus_o_[k].n < (size_t)n_sets + 1) bus_o_[k].alloc(std::max<size_t>((size_t)n_sets + 1, (size_t)opt_.max_batch_reads + 1));
    KB_CK(cudaMemcpyAsync(bus_b_[k].p, bases[k], nbz, cudaMemcpyHostToDevice, st));
    KB_CK(cudaMemcpyAsync(bus_o_[k].p, offs[k], ((size_t)n_sets + 1) * 4, cudaMemcpyHostToDevice, st));
    db[k] = bus_b_[k].p;
    dofs[k] = bus_o_[k].p;
  }
  // longest cDNA read of the batch (sizes the packed-read layout)
  uint32_t maxlen = 0;
  const uint32_t* so = offs[sp.seq_file];
  for (uint32_t i = 0; i < n_sets; ++i) maxlen = std::max(maxlen, so[i + 1] - so[i]);
  if (sp.paired) {
    so = offs[sp.seq2_file];
    for (uint32_t i = 0; i < n_sets; ++i) maxlen = std::max(maxlen, so[i + 1] - so[i]);
  }
  const uint32_t n_rec = bus_core(db, dofs, n_sets, maxlen);
  if (n_rec && records_out) {
    bus_rec_.download(records_out, n_rec, 0, st);
    KB_CK(cudaStreamSynchronize(st));
  }
  if (n_records_out) *n_records_out = n_rec;
}

// Same with the files of the batch already resident in device memory; the records stay on the device
// (bus_records_device(), valid until the next batch).
uint32_t Quant::bus_batch_device(const uint8_t* const* d_bases, const uint32_t* const* d_offs, uint32_t n_sets,
                                 uint32_t max_seq_len) {
  KB_CK(cudaSetDevice(ix_.device));
  if (!opt_.bus) throw Error("kallisto_b200: not a bus run");
  if (n_sets == 0) return 0;
  for (int k = 0; k < opt_.bus_spec.nfiles; ++k)
    if (!d_bases[k] || !d_offs[k]) throw Error("kallisto_b200: bus batch needs bases and offsets for every file of the technology");
  return bus_core(d_bases, d_offs, n_sets, max_seq_len);
}

uint32_t Quant::bus_core(const uint8_t* const* db, const uint32_t* const* dofs, uint32_t n_sets, uint32_t maxlen) {
  const BusSpec& sp = opt_.bus_spec;
  cudaStream_t st = stream_;
  BusArgs a{};
  for (int k = 0; k < sp.nfiles; ++k) { a.bases[k] = db[k]; a.off[k] = dofs[k]; }
  const size_t n1 = (size_t)n_sets + 1;
  auto grow = [&](auto& b, size_t need) { if (b.n < need) b.alloc(std::max<size_t>(need, (size_t)opt_.max_batch_reads + 1)); };
  grow(bus_bc_, n1); grow(bus_umi_, n1); grow(bus_flags_, n1); grow(bus_skip_, n1);
  if (sp.tag_len) grow(bus_notag_, n1);
  grow(bus_isnew_, n1); grow(bus_newrank_, n1); grow(bus_ismapped_, n1); grow(bus_rank_, n1); grow(bus_rec_, n1);
  if (bus_hist_.n < 66) { bus_hist_.alloc(66); bus_hist_.zero(st); }
  if (bus_nvalid_.n < 1) bus_nvalid_.alloc(1);
  if (bus_idof_.n < ix_.dict_cap) { bus_idof_.alloc(ix_.dict_cap); launch_fill_i32(bus_idof_.p, ix_.dict_cap, -1, st); }
  const size_t tb = bus_scan_bytes(n_sets);
  if (bus_tmp_.n < tb) bus_tmp_.alloc(std::max(tb, bus_scan_bytes(opt_.max_batch_reads)));
  bus_nvalid_.zero(st);
  a.n_sets = n_sets;
  a.set_base = n_frag_total_ - bus_sample_base_;      // --num: read numbers restart with every sample (one reader per batch)
  a.spec = sp;
  a.barcode = (uint64_t*)bus_bc_.p; a.umi = (uint64_t*)bus_umi_.p; a.flags = bus_flags_.p; a.skip = bus_skip_.p;
  a.notag = sp.tag_len ? bus_notag_.p : nullptr;
  a.bc_hist = bus_hist_.p; a.umi_hist = bus_hist_.p + 33; a.n_valid = bus_nvalid_.p;
  launch_bus_fields(a, st);
  KB_CK(cudaGetLastError());
  // the cDNA read(s): single-read or paired pseudoalignment with the strand filter of the technology
  uint32_t min_start = (uint32_t)(sp.paired ? std::min(sp.seq_start, sp.seq2_start) : sp.seq_start);
  if (sp.tag_len) {
    // a read set without the tag has no UMI: the sequence read that shares the UMI's file starts where the tag would
    // have started (src/ProcessReads.cpp:1547,1553-1554), the other one where the technology says
    const int at = sp.umi_a[0] - sp.tag_len;
    cur_notag_ = bus_notag_.p;
    cur_alt_start_ = (uint32_t)(sp.umi_f[0] == sp.seq_file ? at : sp.seq_start);
    cur_alt_start2_ = (uint32_t)(sp.paired && sp.umi_f[0] == sp.seq2_file ? at : sp.seq2_start);
    min_start = std::min(min_start, sp.paired ? std::min(cur_alt_start_, cur_alt_start2_) : cur_alt_start_);
  }
  maxlen = maxlen > min_start ? maxlen - min_start : 1;
  const uint64_t base = n_frag_total_;
  cur_skip_ = bus_skip_.p;
  cur_start_ = (uint32_t)sp.seq_start;
  if (sp.paired) {
    // two sequence reads (busopt.paired, src/ProcessReads.cpp:1550-1567,1646-1650): the pair goes through the same
    // match x 2 / intersectKmers / strand filter / mapPair path as `quant` (one buffer per mate)
    cur_start2_ = (uint32_t)sp.seq2_start;
    run_batch(db[sp.seq_file], dofs[sp.seq_file], 2 * n_sets, 0, maxlen, db[sp.seq2_file], dofs[sp.seq2_file]);
  } else {
    run_batch(db[sp.seq_file], dofs[sp.seq_file], n_sets, 0, maxlen);
  }
  cur_skip_ = nullptr;
  cur_start_ = 0;
  cur_start2_ = 0;
  cur_notag_ = nullptr;
  cur_alt_start_ = cur_alt_start2_ = 0;
  launch_bus_records(dd_, bws_->d_handles.p, n_sets, base, bus_next_id_, bus_idof_.p, bus_isnew_.p, bus_newrank_.p,
                     bus_ismapped_.p, bus_rank_.p, (const uint64_t*)bus_bc_.p, (const uint64_t*)bus_umi_.p, bus_flags_.p,
                     bus_rec_.p, bus_tmp_.p, bus_tmp_.n, st);
  KB_CK(cudaGetLastError());
  n_kernel_launches += 4;      // bus_fields, bus_newflag, bus_newid, bus_records (CUB scans not counted)
  uint32_t n_new = 0, n_rec = 0;
  unsigned long long n_valid = 0;
  bus_newrank_.download(&n_new, 1, n_sets, st);
  bus_rank_.download(&n_rec, 1, n_sets, st);
  bus_nvalid_.download(&n_valid, 1, 0, st);
  KB_CK(cudaStreamSynchronize(st));
  bus_next_id_ += n_new;
  bus_valid_total_ += n_valid;
  return n_rec;
}

void Quant::bus_begin_sample(uint64_t barcode) {
  if (!opt_.bus) throw Error("kallisto_b200: not a bus run");
  if (opt_.bus_spec.n_bc != 0) throw Error("kallisto_b200: sample barcodes need a technology without a barcode read");
  opt_.bus_spec.fake_bc = barcode;
  bus_sample_base_ = n_frag_total_;
  // per-sample fragment-length histogram and quota (batchFlens[id] / tlencounts[id], src/ProcessReads.cpp:486-493)
  std::fill(flens_.begin(), flens_.end(), 0u);
  tl_list_.clear();
  tlencount_ = 0;
}

void Quant::bus_lengths(uint32_t* bc_hist, uint32_t* umi_hist) {
  uint32_t h[66] = {0};
  if (bus_hist_.n >= 66) {
    bus_hist_.download(h, 66, 0, stream_);
    KB_CK(cudaStreamSynchronize(stream_));
  }
  memcpy(bc_hist, h, 33 * 4);
  memcpy(umi_hist, h + 33, 33 * 4);
}

void Quant::set_flens(const uint32_t* f) {
  flens_.assign(f, f + 1000);
  tl_list_.clear();
  tlencount_ = 0;
  for (int i = 0; i < 1000; ++i) tlencount_ += flens_[i];
}

namespace {
__global__ void gather_used_kernel(DevDict dd, const uint32_t* used, uint32_t n, uint32_t* cnt, unsigned long long* first,
                                   unsigned long long* word) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t h = used[i];
  cnt[i] = dd.count[h];
  first[i] = dd.first[h];
  word[i] = dd.dslots[h];
}
}  // namespace

const EcTable& Quant::finalize_ecs() {
  if (ecs_valid_) return ecs_;
  KB_CK(cudaSetDevice(ix_.device));
  check_device_errors();
  DBuf<uint32_t> d_used, d_n, d_cnt;
  DBuf<unsigned long long> d_first, d_word;
  d_used.alloc(ix_.dict_cap);
  d_n.alloc(1);
  launch_collect_used(dd_, d_used.p, d_n.p, stream_);
  uint32_t n_used = 0;
  d_n.download(&n_used, 1, 0, stream_);
  unsigned long long pool_top = 0;
  KB_CK(cudaMemcpyAsync(&pool_top, dd_.pool_top, 8, cudaMemcpyDeviceToHost, stream_));
  KB_CK(cudaStreamSynchronize(stream_));
  std::vector<uint32_t> used(n_used), cnt(n_used);
  std::vector<unsigned long long> first(n_used), word(n_used);
  if (n_used) {
    d_cnt.alloc(n_used);
    d_first.alloc(n_used);
    d_word.alloc(n_used);
    gather_used_kernel<<<(n_used + 255) / 256, 256, 0, stream_>>>(dd_, d_used.p, n_used, d_cnt.p, d_first.p, d_word.p);
    d_used.download(used.data(), n_used, 0, stream_);
    d_cnt.download(cnt.data(), n_used, 0, stream_);
    d_first.download(first.data(), n_used, 0, stream_);
    d_word.download(word.data(), n_used, 0, stream_);
  }
  // sets discovered at run time
  std::vector<uint32_t> dyn;
  if (pool_top > ix_.n_index_tids) {
    dyn.resize(pool_top - ix_.n_index_tids);
    pool_.download(dyn.data(), dyn.size(), ix_.n_index_tids, stream_);
  }
  KB_CK(cudaStreamSynchronize(stream_));
  std::vector<uint32_t> order(n_used);
  std::iota(order.begin(), order.end(), 0u);
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return first[a] < first[b]; });
  ecs_ = EcTable();
  ecs_.off.reserve(n_used + 1);
  ecs_.off.push_back(0);
  for (uint32_t oi = 0; oi < n_used; ++oi) {
    const uint32_t i = order[oi];
    const uint32_t off = (uint32_t)word[i];
    const uint32_t len = (uint32_t)((word[i] >> 32) & 0xFFFFFFu);
    const uint32_t* src = off < ix_.n_index_tids ? ix_.flat.ec_tid.data() + off : dyn.data() + (off - ix_.n_index_tids);
    ecs_.tid.insert(ecs_.tid.end(), src, src + len);
    ecs_.off.push_back(ecs_.tid.size());
    ecs_.count.push_back(cnt[i]);
    ecs_.handle.push_back((int32_t)used[i]);
  }
  ecs_valid_ = true;
  return ecs_;
}

Stats Quant::stats() {
  Stats s;
  s.n_processed = n_frag_total_;   // bus: read sets with a bad barcode/UMI are skipped but still counted as processed (ProcessReads.cpp:1372)
  if (dev_stats_valid_) {   // computed on the device by run_em_device: no EC table on the host needed
    s.n_pseudoaligned = dev_pseudoaligned_;
    s.n_unique = dev_unique_;
  } else {
    const EcTable& e = finalize_ecs();
    for (uint32_t i = 0; i < e.n(); ++i) {
      s.n_pseudoaligned += e.count[i];
      if (e.off[i + 1] - e.off[i] == 1) s.n_unique += e.count[i];
    }
  }
  unsigned long long st[4];
  KB_CK(cudaMemcpy(st, dd_.stats, sizeof(st), cudaMemcpyDeviceToHost));
  s.n_probes = st[0];
  s.n_resolved = st[1];
  s.n_memo_hits = st[2];
  s.n_slot_visits = st[3];
  return s;
}

// MinCollector::compute_mean_frag_lens_trunc (src/MinCollector.cpp:629-651) when fld_mean == 0,
// MinCollector::init_mean_fl_trunc + trunc_gaussian_fld (src/MinCollector.cpp:583-627,
// src/weights.cpp:248-296) otherwise.
std::vector<double> mean_fl_trunc_of(const uint32_t* flens, double fld_mean, double fld_sd) {
  const int MAXF = 1000;
  std::vector<double> out(MAXF, 0.0);
  if (fld_mean == 0.0) {
    std::vector<int> counts(MAXF, 0);
    std::vector<double> mass(MAXF, 0.0);
    counts[0] = (int)flens[0];
    for (size_t i = 1; i < (size_t)MAXF; ++i) {
      mass[i] = static_cast<double>(flens[i] * i) + mass[i - 1];
      counts[i] = (int)flens[i] + counts[i - 1];
      if (counts[i] > 0) out[i] = mass[i] / static_cast<double>(counts[i]);
    }
  } else {
    // trunc_gaussian_fld(0, MAX_FRAG_LEN, mean, sd), src/weights.cpp:248-271
    std::vector<double> mean_fl(MAXF, 0.0);
    double total_mass = 0.0, total_density = 0.0;
    for (size_t i = 0; i < (size_t)MAXF; ++i) {
      double x = static_cast<double>(0 + i);
      x = (x - fld_mean) / fld_sd;
      const double cur_density = std::exp(-0.5 * x * x) / fld_sd;
      total_mass += cur_density * i;
      total_density += cur_density;
      if (total_mass > 0) mean_fl[i] = total_mass / total_density;
    }
    out = mean_fl;
  }
  return out;
}

std::vector<double> Quant::mean_fl_trunc(double fld_mean, double fld_sd) const {
  return mean_fl_trunc_of(flens_.data(), fld_mean, fld_sd);
}

namespace {
int em_tpb() {
  if (const char* s = getenv("KB_EM_TPB")) return std::max(32, std::min(1024, atoi(s)));   // tuning knob
  return 1024;     // one block per SM: fewest participants in the grid barrier (20.8 vs 21.6 us per round with 4 x 256)
}
struct EmHost {
  std::vector<uint32_t> multi_ec, m_off, m_tid, t_off, t_midx;
  std::vector<double> m_w, t_w, eff;
  std::vector<int32_t> t_single;
};

// get_frag_len_means + calc_eff_lens + calc_weights (src/weights.cpp:7-28,58-79,220-246)
void em_setup(const FlatIndex& f, const EcTable& ecs, const std::vector<double>& fl_trunc, EmHost& h) {
  const uint32_t T = f.num_targets();
  h.eff.resize(T);
  const double marginal = fl_trunc[999];
  for (uint32_t t = 0; t < T; ++t) {
    const double mean = f.target_len[t] >= 1000 ? marginal : fl_trunc[f.target_len[t]];
    const double len = static_cast<double>(f.target_len[t]);
    double e = len - mean + 1;
    if (e < 1.0) e = len;
    h.eff[t] = e;
  }
  h.t_single.assign(T, -1);
  h.m_off.push_back(0);
  std::vector<uint32_t> deg(T + 1, 0);
  for (uint32_t e = 0; e < ecs.n(); ++e) {
    const uint64_t b = ecs.off[e], n = ecs.off[e + 1] - b;
    if (n == 1) {
      h.t_single[ecs.tid[b]] = (int32_t)e;
      continue;
    }
    h.multi_ec.push_back(e);
    for (uint64_t j = 0; j < n; ++j) {
      const uint32_t t = ecs.tid[b + j];
      h.m_tid.push_back(t);
      h.m_w.push_back(static_cast<double>(ecs.count[e]) / h.eff[t]);
      ++deg[t + 1];
    }
    h.m_off.push_back((uint32_t)h.m_tid.size());
  }
  h.t_off.assign(T + 1, 0);
  for (uint32_t t = 0; t < T; ++t) h.t_off[t + 1] = h.t_off[t] + deg[t + 1];
  h.t_midx.resize(h.m_tid.size());
  h.t_w.resize(h.m_tid.size());
  std::vector<uint32_t> fill(h.t_off.begin(), h.t_off.end() - 1);
  for (uint32_t r = 0; r < h.multi_ec.size(); ++r)
    for (uint32_t j = h.m_off[r]; j < h.m_off[r + 1]; ++j) {
      const uint32_t t = h.m_tid[j];
      h.t_midx[fill[t]] = r;
      h.t_w[fill[t]] = h.m_w[j];
      ++fill[t];
    }
}

struct EmDevice {
  DBuf<uint32_t> multi_ec, m_off, m_tid, t_off, t_midx, counts;
  DBuf<double> m_w, t_w, alpha, norm;
  DBuf<int32_t> t_single;
  DBuf<int> rounds, fstate;
  DBuf<unsigned int> chcount, bar;
  DBuf<uint32_t> cnt_row;
  DBuf<double> single_cnt;
};

void em_upload(const EmHost& h, uint32_t n_ec, uint32_t T, int nb, EmDevice& d, EmProblem& p, cudaStream_t st) {
  const size_t nm = h.multi_ec.size(), nnz = h.m_tid.size();
  d.multi_ec.alloc(std::max<size_t>(1, nm)); d.multi_ec.upload(h.multi_ec.data(), nm, st);
  d.m_off.upload(h.m_off.data(), h.m_off.size(), st);
  d.m_tid.alloc(std::max<size_t>(1, nnz)); d.m_tid.upload(h.m_tid.data(), nnz, st);
  d.m_w.alloc(std::max<size_t>(1, nnz)); d.m_w.upload(h.m_w.data(), nnz, st);
  d.t_off.upload(h.t_off.data(), h.t_off.size(), st);
  d.t_midx.alloc(std::max<size_t>(1, nnz)); d.t_midx.upload(h.t_midx.data(), nnz, st);
  d.t_w.alloc(std::max<size_t>(1, nnz)); d.t_w.upload(h.t_w.data(), nnz, st);
  d.t_single.upload(h.t_single.data(), h.t_single.size(), st);
  d.counts.alloc(std::max<size_t>(1, (size_t)nb * n_ec));
  d.alpha.alloc((size_t)nb * T);
  d.norm.alloc(std::max<size_t>(1, (size_t)nb * nm));
  d.rounds.alloc(nb); d.rounds.zero(st);
  d.bar.alloc(1);
  d.cnt_row.alloc(std::max<size_t>(1, (size_t)nb * nm));
  d.single_cnt.alloc((size_t)nb * T);
  d.fstate.alloc(nb); d.fstate.zero(st);
  d.chcount.alloc((size_t)nb * 2); d.chcount.zero(st);
  p = EmProblem();
  p.n_ec = n_ec; p.n_targets = T; p.n_multi = (uint32_t)nm;
  p.multi_ec = d.multi_ec.p; p.m_off = d.m_off.p; p.m_tid = d.m_tid.p; p.m_w = d.m_w.p;
  p.t_off = d.t_off.p; p.t_midx = d.t_midx.p; p.t_w = d.t_w.p; p.t_single = d.t_single.p;
  p.nb = nb; p.counts = d.counts.p; p.alpha = d.alpha.p; p.norm = d.norm.p;
  p.rounds = d.rounds.p; p.bar = d.bar.p; p.chcount = d.chcount.p; p.fstate = d.fstate.p;
  p.cnt_row = d.cnt_row.p; p.single_cnt = d.single_cnt.p;
}

void em_fetch(const EmProblem& p, EmDevice& d, int nb, uint32_t T, std::vector<double>& alpha, std::vector<int>& rounds,
              cudaStream_t st) {
  alpha.resize((size_t)nb * T);
  rounds.resize(nb);
  std::vector<int> state(nb);
  d.alpha.download(alpha.data(), alpha.size(), 0, st);
  d.rounds.download(rounds.data(), nb, 0, st);
  d.fstate.download(state.data(), nb, 0, st);
  KB_CK(cudaStreamSynchronize(st));
  for (int b = 0; b < nb; ++b)
    if (state[b] == 3)   // stop detected on the last allowed iteration: zero small alphas (EMAlgorithm.h:213-216)
      for (uint32_t t = 0; t < T; ++t)
        if (alpha[(size_t)b * T + t] < 1e-7 / 10.0) alpha[(size_t)b * T + t] = 0.0;
}
}  // namespace

EmResult Quant::run_em(const EcTable& ecs, const std::vector<double>& fl_trunc, int max_iter, int min_rounds) {
  KB_CK(cudaSetDevice(ix_.device));
  const FlatIndex& f = ix_.flat;
  const uint32_t T = f.num_targets();
  EmHost h;
  em_setup(f, ecs, fl_trunc, h);
  EmDevice d;
  EmProblem p;
  em_upload(h, ecs.n(), T, 1, d, p, stream_);
  d.counts.upload(ecs.count.data(), ecs.n(), stream_);
  std::vector<double> a0(T, 1.0 / T);   // uniform start (EMAlgorithm.h:38)
  d.alpha.upload(a0.data(), T, stream_);
  p.max_iter = max_iter;
  p.min_rounds = min_rounds;
  cudaEvent_t e0, e1;
  KB_CK(cudaEventCreate(&e0));
  KB_CK(cudaEventCreate(&e1));
  KB_CK(cudaEventRecord(e0, stream_));
  launch_em(p, 256, stream_);
  KB_CK(cudaGetLastError());
  KB_CK(cudaEventRecord(e1, stream_));
  EmResult r;
  std::vector<int> rounds;
  em_fetch(p, d, 1, T, r.alpha, rounds, stream_);
  float ms = 0;
  KB_CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  r.seconds = ms * 1e-3;
  last_em_seconds = r.seconds;
  r.rounds = rounds[0];
  r.eff_lens = h.eff;
  return r;
}

// The whole tail of `kallisto quant` without the EC table ever visiting the host: EC ids by first
// occurrence, CSR/CSC + weights on the device (kernels_emprep.cu), then em_kernel.
EmResult Quant::run_em_device(const std::vector<double>& fl_trunc, int max_iter, int min_rounds) {
  KB_CK(cudaSetDevice(ix_.device));
  check_device_errors();
  // the EM works out of L2: give it the part that the k-mer presence filter held as persisting lines during pseudoalignment
  if (ix_.l2_persist_bytes) cudaCtxResetPersistingL2Cache();
  const FlatIndex& f = ix_.flat;
  const uint32_t T = f.num_targets();
  cudaStream_t st = stream_;
  EmWs& w = *emws_;
  // effective lengths (T values, host arithmetic identical to calc_eff_lens)
  EmResult res;
  res.eff_lens.resize(T);
  {
    const double marginal = fl_trunc[999];
    for (uint32_t t = 0; t < T; ++t) {
      const double mean = f.target_len[t] >= 1000 ? marginal : fl_trunc[f.target_len[t]];
      const double len = static_cast<double>(f.target_len[t]);
      double e = len - mean + 1;
      if (e < 1.0) e = len;
      res.eff_lens[t] = e;
    }
  }
  const bool trace = getenv("KB_EM_TRACE") != nullptr;     // host wall clock of the set-up steps on stderr
  double t_last = now_s();
  auto mark = [&](const char* what) {
    if (!trace) return;
    const double t = now_s();
    fprintf(stderr, "[em-trace] %s: %.3f ms\n", what, (t - t_last) * 1e3);
    t_last = t;
  };
  cudaEvent_t e0, e1, e2;
  KB_CK(cudaEventCreate(&e0));
  KB_CK(cudaEventCreate(&e1));
  KB_CK(cudaEventCreate(&e2));
  KB_CK(cudaEventRecord(e0, st));
  mark("eff lens + events");
  // ---- phase 1: used handles, sorted by first occurrence; lengths and offsets
  if (w.used.n < ix_.dict_cap) w.used.alloc(ix_.dict_cap);
  if (w.scal.n < 8) w.scal.alloc(8);
  launch_collect_used(dd_, w.used.p, w.scal.p, st);
  uint32_t n = 0;
  w.scal.download(&n, 1, 0, st);
  KB_CK(cudaStreamSynchronize(st));
  mark("drain stream + collect used handles");
  res.alpha.assign(T, 0.0);
  if (n == 0) {
    res.rounds = 0;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    return res;
  }
  const size_t n1 = (size_t)n + 1;
  auto grow32 = [](DBuf<uint32_t>& b, size_t need) { if (b.n < need) b.alloc(need + need / 4); };
  auto growd = [](DBuf<double>& b, size_t need) { if (b.n < need) b.alloc(need + need / 4); };
  if (w.key_in.n < n1) { w.key_in.alloc(n1 + n1 / 4); w.key_out.alloc(n1 + n1 / 4); }
  grow32(w.idx_in, n1); grow32(w.order, n1); grow32(w.handle, n1); grow32(w.count, n1); grow32(w.len, n1);
  grow32(w.multi_len, n1); grow32(w.is_multi, n1); grow32(w.ec_off, n1); grow32(w.m_off, n1); grow32(w.multi_index, n1);
  grow32(w.minkey, n1); grow32(w.ckey, n1); grow32(w.cval, n1); grow32(w.ckey_out, n1); grow32(w.rlen, n1 + 1);
  const uint32_t nnz_guess = std::max<uint32_t>(n * 4, 1u << 20);
  size_t tmp_need = emprep_sort_bytes(n + 1, std::max<uint32_t>(nnz_guess, T + 1));
  if (w.tmp.n < tmp_need) w.tmp.alloc(tmp_need);
  mark("phase-1 buffers");
  // the scans run over n + 1 items: the extra item must be zero
  KB_CK(cudaMemsetAsync(w.len.p + n, 0, 4, st));
  KB_CK(cudaMemsetAsync(w.multi_len.p + n, 0, 4, st));
  KB_CK(cudaMemsetAsync(w.is_multi.p + n, 0, 4, st));
  emprep_sort_by_first(dd_, w.used.p, n, w.key_in.p, w.key_out.p, w.idx_in.p, w.order.p, w.tmp.p, w.tmp.n, st);
  EmPrep ep{};
  ep.n_ec = n; ep.n_targets = T;
  ep.handle = w.handle.p; ep.count = w.count.p; ep.len = w.len.p; ep.ec_off = w.ec_off.p; ep.m_off = w.m_off.p;
  ep.multi_index = w.multi_index.p; ep.minkey = w.minkey.p;
  emprep_meta(dd_, w.used.p, w.order.p, n, ep, w.multi_len.p, w.is_multi.p, w.tmp.p, w.tmp.n, st);
  uint32_t tot[3] = {0, 0, 0};
  w.ec_off.download(&tot[0], 1, n, st);
  w.m_off.download(&tot[1], 1, n, st);
  w.multi_index.download(&tot[2], 1, n, st);
  KB_CK(cudaStreamSynchronize(st));
  mark("sort by first occurrence + scans");
  const uint32_t nnz_all = tot[0], nnz = tot[1], n_multi = tot[2];
  // ---- phase 2: EC table, CSR, weights, CSC
  tmp_need = emprep_sort_bytes(n + 1, std::max<uint32_t>(nnz, T + 1));
  if (w.tmp.n < tmp_need) w.tmp.alloc(tmp_need);
  grow32(w.ec_tid, std::max<uint32_t>(1, nnz_all)); grow32(w.multi_ec, (size_t)n_multi + 1); grow32(w.m_rowoff, (size_t)n_multi + 2);
  const size_t nz = std::max<uint32_t>(1, nnz);
  grow32(w.m_tid, nz); grow32(w.m_row, nz); grow32(w.m_iota, nz); grow32(w.sortv, nz); grow32(w.t_midx, nz);
  if (w.k64_in.n < nz) { w.k64_in.alloc(nz + nz / 4); w.k64_out.alloc(nz + nz / 4); }
  if (w.bar.n < 1) w.bar.alloc(1);
  growd(w.m_w, nz); growd(w.t_w, nz);
  grow32(w.t_deg, (size_t)T + 1); grow32(w.t_off, (size_t)T + 1);
  if (w.t_single.n < T) w.t_single.alloc(T);
  growd(w.eff, T); growd(w.alpha, T); growd(w.norm, (size_t)n_multi + 1);
  grow32(w.cnt_row, (size_t)n_multi + 1); growd(w.single_cnt, T);
  mark("phase-2 buffers");
  KB_CK(cudaMemsetAsync(w.t_deg.p, 0, ((size_t)T + 1) * 4, st));
  launch_fill_i32(w.t_single.p, T, -1, st);
  KB_CK(cudaMemsetAsync(w.m_rowoff.p, 0, 4, st));   // n_multi == 0: offsets [0]
  w.eff.upload(res.eff_lens.data(), T, st);
  ep.n_multi = n_multi;
  ep.ec_tid = w.ec_tid.p; ep.multi_ec = w.multi_ec.p; ep.m_rowoff = w.m_rowoff.p; ep.m_tid = w.m_tid.p; ep.m_w = w.m_w.p;
  ep.m_row = w.m_row.p; ep.m_iota = w.m_iota.p; ep.t_deg = w.t_deg.p; ep.t_off = w.t_off.p; ep.t_midx = w.t_midx.p;
  ep.t_w = w.t_w.p; ep.t_single = w.t_single.p; ep.eff = w.eff.p; ep.k64_in = w.k64_in.p;
  emprep_rows(ep, w.is_multi.p, w.ckey.p, w.cval.p, w.ckey_out.p, w.rlen.p, w.tmp.p, w.tmp.n, st);
  emprep_fill(dd_, ep, nnz, w.k64_out.p, w.sortv.p, w.tmp.p, w.tmp.n, (unsigned long long*)w.key_in.p, st);
  KB_CK(cudaGetLastError());
  // ---- EM
  if (w.emi.n < 8) w.emi.alloc(8);
  if (w.chcount.n < 2) w.chcount.alloc(2);
  w.emi.zero(st);
  w.chcount.zero(st);
  launch_fill_f64(w.alpha.p, T, 1.0 / T, st);   // uniform start (EMAlgorithm.h:38)
  EmProblem p{};
  p.n_ec = n; p.n_targets = T; p.n_multi = n_multi;
  p.multi_ec = w.multi_ec.p; p.m_off = w.m_rowoff.p; p.m_tid = w.m_tid.p; p.m_w = w.m_w.p;
  p.t_off = w.t_off.p; p.t_midx = w.t_midx.p; p.t_w = w.t_w.p; p.t_single = w.t_single.p;
  p.nb = 1; p.counts = w.count.p; p.alpha = w.alpha.p; p.norm = w.norm.p;
  p.rounds = w.emi.p; p.bar = w.bar.p; p.fstate = w.emi.p + 3; p.chcount = w.chcount.p;
  p.cnt_row = w.cnt_row.p; p.single_cnt = w.single_cnt.p;
  p.max_iter = max_iter; p.min_rounds = min_rounds;
  mark("fill launches + uploads");
  // collect_used, gather_used, ec_meta, multi_compact, row_len, ec_fill, csc_fill, stats, fill_i32, fill_f64,
  // em_gather_counts + em_kernel
  n_kernel_launches += 11 + 1;
  KB_CK(cudaEventRecord(e1, st));
  launch_em(p, em_tpb(), st);
  KB_CK(cudaGetLastError());
  KB_CK(cudaEventRecord(e2, st));
  int emi[4] = {0, 0, 0, 0};
  unsigned long long s2[2] = {0, 0};
  w.alpha.download(res.alpha.data(), T, 0, st);
  w.emi.download(emi, 4, 0, st);
  w.key_in.download(s2, 2, 0, st);
  KB_CK(cudaStreamSynchronize(st));
  mark("EM kernel + results");
  if (emi[3] == 3)
    for (uint32_t t = 0; t < T; ++t)
      if (res.alpha[t] < 1e-7 / 10.0) res.alpha[t] = 0.0;
  float ms_prep = 0, ms_em = 0;
  KB_CK(cudaEventElapsedTime(&ms_prep, e0, e1));
  KB_CK(cudaEventElapsedTime(&ms_em, e1, e2));
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  res.rounds = emi[0];
  res.seconds = ms_em * 1e-3;
  last_em_seconds = res.seconds;
  last_prep_seconds = ms_prep * 1e-3;
  dev_stats_valid_ = true;
  dev_problem_valid_ = true;
  dev_n_multi_ = n_multi;
  dev_n_ecs_ = n; dev_nnz_ = nnz_all; dev_pseudoaligned_ = s2[0]; dev_unique_ = s2[1];
  return res;
}

std::vector<int> Quant::run_bootstrap_device(const std::vector<double>& fl_trunc, uint64_t seed, int B, std::vector<double>& alpha_out,
                                             std::vector<uint32_t>* samples_out, double* ms_out) {
  KB_CK(cudaSetDevice(ix_.device));
  std::vector<int> rounds;
  if (B <= 0) { alpha_out.clear(); return rounds; }
  if (!dev_problem_valid_) run_em_device(fl_trunc);      // builds the matrices (and runs the main EM once)
  const uint32_t T = ix_.flat.num_targets();
  alpha_out.assign((size_t)B * T, 0.0);
  rounds.assign(B, 0);
  if (!dev_problem_valid_) return rounds;                // nothing pseudoaligned
  if (ix_.l2_persist_bytes) cudaCtxResetPersistingL2Cache();
  cudaStream_t st = stream_;
  EmWs& w = *emws_;
  const uint32_t nE = (uint32_t)dev_n_ecs_, n_multi = dev_n_multi_;
  // seeds (src/main.cpp:2746-2752)
  std::mt19937_64 rnd;
  rnd.seed(seed);
  std::vector<uint32_t> x0(B);
  for (int b = 0; b < B; ++b) {
    const uint64_t s = rnd();
    uint32_t x = (uint32_t)(s % 2147483647ull);      // minstd_rand0 seeding (libstdc++ linear_congruential_engine::seed)
    if (x == 0) x = 1;
    x0[b] = x;
  }
  // std::discrete_distribution<int>(counts): normalised probabilities and their partial sums, in the host's own
  // sequential double arithmetic (a parallel scan would round differently)
  std::vector<uint32_t> cnt(nE);
  w.count.download(cnt.data(), nE, 0, st);
  KB_CK(cudaStreamSynchronize(st));
  uint64_t N = 0;
  for (uint32_t e = 0; e < nE; ++e) N += cnt[e];
  const int n_draws = (int)N;   // Multinomial::n_ is an int
  if (w.bs_counts.n < (size_t)B * nE) w.bs_counts.alloc((size_t)B * nE);
  cudaEvent_t e0, e1, e2;
  KB_CK(cudaEventCreate(&e0)); KB_CK(cudaEventCreate(&e1)); KB_CK(cudaEventCreate(&e2));
  KB_CK(cudaEventRecord(e0, st));
  if (nE >= 2) {
    std::vector<double> prob(cnt.begin(), cnt.end());
    const double sum = std::accumulate(prob.begin(), prob.end(), 0.0);
    for (auto& v : prob) v /= sum;
    std::vector<double> cp(nE);
    std::partial_sum(prob.begin(), prob.end(), cp.begin());
    cp[nE - 1] = 1.0;
    if (w.bs_cp.n < nE) w.bs_cp.alloc(nE);
    if (w.bs_x0.n < (size_t)B) w.bs_x0.alloc(B);
    w.bs_cp.upload(cp.data(), nE, st);
    w.bs_x0.upload(x0.data(), B, st);
    ResampleArgs ra{};
    ra.cp = w.bs_cp.p; ra.n_ec = nE; ra.n_draws = n_draws > 0 ? (uint64_t)n_draws : 0; ra.nb = B; ra.x0 = w.bs_x0.p;
    ra.samp = w.bs_counts.p;
    launch_resample(ra, st);
    KB_CK(cudaGetLastError());
    KB_CK(cudaStreamSynchronize(st));     // cp / x0 are local vectors
    ++n_kernel_launches;
  } else {
    // a single class: discrete_distribution returns 0 without consuming the engine
    std::vector<uint32_t> s1((size_t)B * nE, 0);
    for (int b = 0; b < B && nE == 1; ++b) s1[b] = n_draws > 0 ? (uint32_t)n_draws : 0;
    w.bs_counts.upload(s1.data(), s1.size(), st);
    KB_CK(cudaStreamSynchronize(st));
  }
  KB_CK(cudaEventRecord(e1, st));
  if (samples_out) {
    samples_out->resize((size_t)B * nE);
    w.bs_counts.download(samples_out->data(), samples_out->size(), 0, st);
    KB_CK(cudaStreamSynchronize(st));
  }
  // the B problems, `chunk` at a time: alpha + norm + counts of a chunk are meant to stay in L2 next to the shared matrices
  int chunk = B;
  {
    const size_t per = ((size_t)T + n_multi) * 8 + (size_t)nE * 4;
    const size_t budget = 64u << 20;
    chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)B, budget / std::max<size_t>(1, per)));
    if (const char* s = getenv("KB_BS_CHUNK")) { const int v = atoi(s); if (v > 0) chunk = std::min(B, v); }   // tuning knob
  }
  if (w.bs_alpha.n < (size_t)chunk * T) w.bs_alpha.alloc((size_t)chunk * T);
  if (w.bs_norm.n < (size_t)chunk * std::max<uint32_t>(1, n_multi)) w.bs_norm.alloc((size_t)chunk * std::max<uint32_t>(1, n_multi));
  if (w.cnt_row.n < (size_t)chunk * std::max<uint32_t>(1, n_multi)) w.cnt_row.alloc((size_t)chunk * std::max<uint32_t>(1, n_multi));
  if (w.single_cnt.n < (size_t)chunk * T) w.single_cnt.alloc((size_t)chunk * T);
  if (w.bs_emi.n < (size_t)2 * chunk) w.bs_emi.alloc((size_t)2 * chunk);
  if (w.bs_ch.n < (size_t)2 * chunk) w.bs_ch.alloc((size_t)2 * chunk);
  std::vector<int> emi((size_t)2 * chunk);
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = std::min(chunk, B - b0);
    launch_fill_f64(w.bs_alpha.p, (uint64_t)nb * T, 1.0 / T, st);
    w.bs_emi.zero(st);
    w.bs_ch.zero(st);
    EmProblem p{};
    p.n_ec = nE; p.n_targets = T; p.n_multi = n_multi;
    p.multi_ec = w.multi_ec.p; p.m_off = w.m_rowoff.p; p.m_tid = w.m_tid.p; p.m_w = w.m_w.p;
    p.t_off = w.t_off.p; p.t_midx = w.t_midx.p; p.t_w = w.t_w.p; p.t_single = w.t_single.p;
    p.nb = nb; p.counts = w.bs_counts.p + (size_t)b0 * nE; p.alpha = w.bs_alpha.p; p.norm = w.bs_norm.p;
    p.rounds = w.bs_emi.p; p.fstate = w.bs_emi.p + chunk; p.bar = w.bar.p; p.chcount = w.bs_ch.p;
    p.cnt_row = w.cnt_row.p; p.single_cnt = w.single_cnt.p;
    p.max_iter = 10000; p.min_rounds = 50;
    launch_em(p, em_tpb(), st);
    KB_CK(cudaGetLastError());
    n_kernel_launches += 2;
    w.bs_alpha.download(alpha_out.data() + (size_t)b0 * T, (size_t)nb * T, 0, st);
    w.bs_emi.download(emi.data(), (size_t)2 * chunk, 0, st);
    KB_CK(cudaStreamSynchronize(st));
    for (int b = 0; b < nb; ++b) {
      rounds[b0 + b] = emi[b];
      if (emi[chunk + b] == 3)   // stop detected on the last allowed iteration: zero small alphas (EMAlgorithm.h:213-216)
        for (uint32_t t = 0; t < T; ++t)
          if (alpha_out[(size_t)(b0 + b) * T + t] < 1e-7 / 10.0) alpha_out[(size_t)(b0 + b) * T + t] = 0.0;
    }
  }
  KB_CK(cudaEventRecord(e2, st));
  KB_CK(cudaEventSynchronize(e2));
  if (ms_out) {
    float a = 0, b = 0;
    KB_CK(cudaEventElapsedTime(&a, e0, e1));
    KB_CK(cudaEventElapsedTime(&b, e1, e2));
    ms_out[0] = a; ms_out[1] = b;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  return rounds;
}

void Quant::export_prepare(uint32_t* n_sets, uint32_t* n_entries) {
  KB_CK(cudaSetDevice(ix_.device));
  check_device_errors();
  cudaStream_t st = stream_;
  EmWs& w = *emws_;
  if (w.used.n < ix_.dict_cap) w.used.alloc(ix_.dict_cap);
  if (w.scal.n < 8) w.scal.alloc(8);
  launch_collect_used(dd_, w.used.p, w.scal.p, st);
  uint32_t n = 0;
  w.scal.download(&n, 1, 0, st);
  KB_CK(cudaStreamSynchronize(st));
  exp_n_ = n;
  exp_nnz_ = 0;
  if (n) {
    const size_t n1 = (size_t)n + 1;
    auto grow32 = [](DBuf<uint32_t>& b, size_t need) { if (b.n < need) b.alloc(need + need / 4); };
    if (w.key_in.n < n1) { w.key_in.alloc(n1 + n1 / 4); w.key_out.alloc(n1 + n1 / 4); }
    grow32(w.idx_in, n1); grow32(w.order, n1); grow32(w.handle, n1); grow32(w.count, n1); grow32(w.len, n1);
    grow32(w.multi_len, n1); grow32(w.is_multi, n1); grow32(w.ec_off, n1); grow32(w.m_off, n1); grow32(w.multi_index, n1);
    grow32(w.minkey, n1);
    const size_t tmp_need = emprep_sort_bytes(n + 1, 1u << 20);
    if (w.tmp.n < tmp_need) w.tmp.alloc(tmp_need);
    KB_CK(cudaMemsetAsync(w.len.p + n, 0, 4, st));
    KB_CK(cudaMemsetAsync(w.multi_len.p + n, 0, 4, st));
    KB_CK(cudaMemsetAsync(w.is_multi.p + n, 0, 4, st));
    emprep_sort_by_first(dd_, w.used.p, n, w.key_in.p, w.key_out.p, w.idx_in.p, w.order.p, w.tmp.p, w.tmp.n, st);
    EmPrep ep{};
    ep.n_ec = n; ep.n_targets = ix_.flat.num_targets();
    ep.handle = w.handle.p; ep.count = w.count.p; ep.len = w.len.p; ep.ec_off = w.ec_off.p; ep.m_off = w.m_off.p;
    ep.multi_index = w.multi_index.p; ep.minkey = w.minkey.p;
    emprep_meta(dd_, w.used.p, w.order.p, n, ep, w.multi_len.p, w.is_multi.p, w.tmp.p, w.tmp.n, st);
    w.ec_off.download(&exp_nnz_, 1, n, st);
    KB_CK(cudaStreamSynchronize(st));
    grow32(w.ec_tid, std::max<uint32_t>(1, exp_nnz_));
    ep.ec_tid = w.ec_tid.p;
    emprep_fill_table(dd_, ep, st);
    KB_CK(cudaGetLastError());
  }
  *n_sets = exp_n_;
  *n_entries = exp_nnz_;
}

void Quant::export_copy(uint32_t* d_off, uint32_t* d_tids, uint32_t* d_counts, unsigned long long* d_first) {
  KB_CK(cudaSetDevice(ix_.device));
  EmWs& w = *emws_;
  cudaStream_t st = stream_;
  if (exp_n_) {
    KB_CK(cudaMemcpyAsync(d_off, w.ec_off.p, ((size_t)exp_n_ + 1) * 4, cudaMemcpyDeviceToDevice, st));
    KB_CK(cudaMemcpyAsync(d_tids, w.ec_tid.p, (size_t)exp_nnz_ * 4, cudaMemcpyDeviceToDevice, st));
    KB_CK(cudaMemcpyAsync(d_counts, w.count.p, (size_t)exp_n_ * 4, cudaMemcpyDeviceToDevice, st));
    KB_CK(cudaMemcpyAsync(d_first, w.key_out.p, (size_t)exp_n_ * 8, cudaMemcpyDeviceToDevice, st));
  } else {
    KB_CK(cudaMemsetAsync(d_off, 0, 4, st));
  }
  KB_CK(cudaStreamSynchronize(st));
}

uint64_t Quant::merge_local(const std::vector<Quant*>& others, uint64_t first_stride) {
  uint64_t total = n_frag_total_;
  size_t sum_n = 0, sum_nnz = 0;
  std::vector<std::pair<uint32_t, uint32_t>> sizes;
  for (Quant* o : others) {
    uint32_t n = 0, nnz = 0;
    o->export_prepare(&n, &nnz);                 // numbers o's ECs on ITS device (synchronises o's stream)
    sizes.push_back({n, nnz});
    sum_n += n;
    sum_nnz += nnz;
    total += o->n_frag_total_;
  }
  KB_CK(cudaSetDevice(ix_.device));
  check_device_errors();
  if (lm_off_.n < sum_n + others.size() + 1) lm_off_.alloc(sum_n + others.size() + 1);
  if (lm_counts_.n < sum_n + 1) lm_counts_.alloc(sum_n + 1);
  if (lm_first_.n < sum_n + 1) lm_first_.alloc(sum_n + 1);
  if (lm_tids_.n < sum_nnz + 1) lm_tids_.alloc(sum_nnz + 1);
  std::vector<ImportSeg> segs;
  size_t o_off = 0, o_n = 0, o_nnz = 0;
  for (size_t i = 0; i < others.size(); ++i) {
    Quant* o = others[i];
    const uint32_t n = sizes[i].first, nnz = sizes[i].second;
    if (n) {
      EmWs& w = *o->emws_;
      const int src = o->ix_.device, dst = ix_.device;
      KB_CK(cudaMemcpyPeerAsync(lm_off_.p + o_off, dst, w.ec_off.p, src, ((size_t)n + 1) * 4, stream_));
      KB_CK(cudaMemcpyPeerAsync(lm_tids_.p + o_nnz, dst, w.ec_tid.p, src, (size_t)nnz * 4, stream_));
      KB_CK(cudaMemcpyPeerAsync(lm_counts_.p + o_n, dst, w.count.p, src, (size_t)n * 4, stream_));
      KB_CK(cudaMemcpyPeerAsync(lm_first_.p + o_n, dst, w.key_out.p, src, (size_t)n * 8, stream_));
      ImportSeg sg;
      sg.n_sets = n; sg.off = lm_off_.p + o_off; sg.tids = lm_tids_.p + o_nnz; sg.counts = lm_counts_.p + o_n; sg.first = lm_first_.p + o_n;
      segs.push_back(sg);
      if (first_stride) throw Error("kallisto_b200: merge_local expects runs fed with global fragment indices");
    }
    o_off += (size_t)n + 1; o_n += n; o_nnz += nnz;
  }
  ecs_valid_ = false;
  dev_stats_valid_ = false;
  dev_problem_valid_ = false;
  for (size_t i = 0; i < segs.size(); i += KB_IMPORT_SEGS) {
    launch_import_segments(dd_, segs.data() + i, (int)std::min<size_t>(KB_IMPORT_SEGS, segs.size() - i), stream_);
    KB_CK(cudaGetLastError());
    ++n_kernel_launches;
  }
  KB_CK(cudaStreamSynchronize(stream_));
  n_frag_total_ = total;
  return total;
}

void Quant::import_sets_device(uint32_t n_sets, const uint32_t* d_off, const uint32_t* d_tids, const uint32_t* d_counts,
                               const unsigned long long* d_first, unsigned long long first_offset) {
  KB_CK(cudaSetDevice(ix_.device));
  ecs_valid_ = false;
  dev_stats_valid_ = false;
  dev_problem_valid_ = false;
  launch_import_sets(dd_, n_sets, d_off, d_tids, d_counts, d_first, first_offset, stream_);
  KB_CK(cudaGetLastError());
  KB_CK(cudaStreamSynchronize(stream_));
}

std::vector<int> Quant::run_bootstrap(const EcTable& ecs, const std::vector<double>& fl_trunc, uint64_t seed, int B,
                                      std::vector<double>& alpha_out, std::vector<uint32_t>* samples_out) {
  KB_CK(cudaSetDevice(ix_.device));
  const FlatIndex& f = ix_.flat;
  const uint32_t T = f.num_targets();
  const uint32_t nE = ecs.n();
  std::vector<int> rounds;
  if (B <= 0) { alpha_out.clear(); return rounds; }
  // seeds (src/main.cpp:2746-2752)
  std::mt19937_64 rnd;
  rnd.seed(seed);
  std::vector<uint32_t> x0(B);
  for (int b = 0; b < B; ++b) {
    const uint64_t s = rnd();
    uint32_t x = (uint32_t)(s % 2147483647ull);      // minstd_rand0 seeding (libstdc++ linear_congruential_engine::seed)
    if (x == 0) x = 1;
    x0[b] = x;
  }
  // std::discrete_distribution<int>(counts): normalised probabilities and their partial sums
  std::vector<double> cp;
  uint64_t N = 0;
  for (uint32_t e = 0; e < nE; ++e) N += ecs.count[e];
  const int n_draws = (int)N;   // Multinomial::n_ is an int
  if (nE >= 2) {
    std::vector<double> prob(ecs.count.begin(), ecs.count.end());
    const double sum = std::accumulate(prob.begin(), prob.end(), 0.0);
    for (auto& v : prob) v /= sum;
    cp.resize(nE);
    std::partial_sum(prob.begin(), prob.end(), cp.begin());
    cp[nE - 1] = 1.0;
  }
  EmHost h;
  em_setup(f, ecs, fl_trunc, h);
  EmDevice d;
  EmProblem p;
  em_upload(h, nE, T, B, d, p, stream_);
  if (nE >= 2) {
    DBuf<double> d_cp;
    DBuf<uint32_t> d_x0;
    d_cp.upload(cp.data(), nE, stream_);
    d_x0.upload(x0.data(), B, stream_);
    ResampleArgs ra{};
    ra.cp = d_cp.p; ra.n_ec = nE; ra.n_draws = n_draws > 0 ? (uint64_t)n_draws : 0; ra.nb = B; ra.x0 = d_x0.p;
    ra.samp = d.counts.p;
    launch_resample(ra, stream_);
    KB_CK(cudaGetLastError());
    KB_CK(cudaStreamSynchronize(stream_));
  } else {
    // a single class: discrete_distribution returns 0 without consuming the engine
    std::vector<uint32_t> s((size_t)B * nE, 0);
    for (int b = 0; b < B && nE == 1; ++b) s[b] = n_draws > 0 ? (uint32_t)n_draws : 0;
    d.counts.upload(s.data(), s.size(), stream_);
    KB_CK(cudaStreamSynchronize(stream_));
  }
  if (samples_out) {
    samples_out->resize((size_t)B * nE);
    d.counts.download(samples_out->data(), samples_out->size(), 0, stream_);
  }
  std::vector<double> a0((size_t)B * T, 1.0 / T);
  d.alpha.upload(a0.data(), a0.size(), stream_);
  p.max_iter = 10000;
  p.min_rounds = 50;
  launch_em(p, 256, stream_);
  KB_CK(cudaGetLastError());
  em_fetch(p, d, B, T, alpha_out, rounds, stream_);
  return rounds;
}

std::vector<int> tcc_run(Index& ix, const TccInput& in, std::vector<double>& alpha_out) {
  KB_CK(cudaSetDevice(ix.device));
  const uint32_t T = ix.flat.num_targets(), nE = in.n_ecs, S = in.n_samples;
  alpha_out.assign((size_t)S * T, 0.0);
  std::vector<int> rounds(S, 0);
  if (S == 0) return rounds;
  // structure of the problem on the host, from the EC table (EC ids = line numbers of matrix.ec)
  std::vector<uint32_t> multi_ec, m_off{0}, m_tid, m_ec, t_off(T + 1, 0), t_midx, t_ec, t_tid;
  std::vector<int32_t> t_single(T, -1);
  for (uint32_t e = 0; e < nE; ++e) {
    const uint64_t b = in.ec_off[e], n = in.ec_off[e + 1] - b;
    for (uint64_t j = 0; j < n; ++j)
      if (in.tids[b + j] >= T) throw Error("kallisto_b200: equivalence class file has a transcript id out of range");
    if (n == 1) { t_single[in.tids[b]] = (int32_t)e; continue; }
    if (n == 0) continue;
    multi_ec.push_back(e);
    for (uint64_t j = 0; j < n; ++j) {
      m_tid.push_back(in.tids[b + j]);
      m_ec.push_back(e);
      ++t_off[in.tids[b + j] + 1];
    }
    m_off.push_back((uint32_t)m_tid.size());
  }
  const uint64_t nnz = m_tid.size();
  const uint32_t n_multi = (uint32_t)multi_ec.size();
  for (uint32_t t = 0; t < T; ++t) t_off[t + 1] += t_off[t];
  t_midx.resize(nnz); t_ec.resize(nnz); t_tid.resize(nnz);
  {
    std::vector<uint32_t> fill(t_off.begin(), t_off.end() - 1);
    for (uint32_t r = 0; r < n_multi; ++r)
      for (uint32_t j = m_off[r]; j < m_off[r + 1]; ++j) {
        const uint32_t t = m_tid[j], at = fill[t]++;
        t_midx[at] = r; t_ec[at] = multi_ec[r]; t_tid[at] = t;
      }
  }
  for (uint64_t i = 0; i < in.row_off[S]; ++i)
    if (in.ec_ids[i] >= nE) throw Error("kallisto_b200: TCC file refers to an equivalence class that is not in the EC file");
  cudaStream_t st = nullptr;
  KB_CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  DBuf<uint32_t> d_multi_ec, d_m_off, d_m_tid, d_m_ec, d_t_off, d_t_midx, d_t_ec, d_t_tid, d_ecid, d_val, d_counts;
  DBuf<int32_t> d_single;
  DBuf<unsigned long long> d_rowoff;
  DBuf<double> d_eff, d_mw, d_tw, d_alpha, d_norm, d_single_cnt;
  DBuf<uint32_t> d_cnt_row;
  DBuf<int> d_emi;
  DBuf<unsigned> d_ch, d_bar;
  auto up32 = [&](DBuf<uint32_t>& d, const std::vector<uint32_t>& h) { d.alloc(std::max<size_t>(1, h.size())); d.upload(h.data(), h.size(), st); };
  up32(d_multi_ec, multi_ec); up32(d_m_off, m_off); up32(d_m_tid, m_tid); up32(d_m_ec, m_ec);
  up32(d_t_off, t_off); up32(d_t_midx, t_midx); up32(d_t_ec, t_ec); up32(d_t_tid, t_tid);
  d_single.upload(t_single.data(), T, st);
  d_ecid.alloc(std::max<uint64_t>(1, in.row_off[S])); d_ecid.upload(in.ec_ids, in.row_off[S], st);
  d_val.alloc(std::max<uint64_t>(1, in.row_off[S])); d_val.upload(in.counts, in.row_off[S], st);
  // samples per chunk: weights dominate (16 bytes per entry and sample); ~2 GB of work space
  const size_t per = (size_t)nnz * 16 + (size_t)nE * 4 + ((size_t)T + n_multi) * 8;
  int chunk = (int)std::max<size_t>(1, std::min<size_t>(S, ((size_t)2 << 30) / std::max<size_t>(1, per)));
  if (const char* s = getenv("KB_TCC_CHUNK")) { const int v = atoi(s); if (v > 0) chunk = std::min<int>((int)S, v); }   // tests
  d_counts.alloc((size_t)chunk * std::max<uint32_t>(1, nE));
  d_mw.alloc(std::max<size_t>(1, (size_t)chunk * nnz)); d_tw.alloc(std::max<size_t>(1, (size_t)chunk * nnz));
  d_alpha.alloc((size_t)chunk * T); d_norm.alloc(std::max<size_t>(1, (size_t)chunk * n_multi));
  d_cnt_row.alloc(std::max<size_t>(1, (size_t)chunk * n_multi)); d_single_cnt.alloc((size_t)chunk * T);
  d_eff.alloc(in.per_sample_eff ? (size_t)chunk * T : (size_t)T);
  if (!in.per_sample_eff) d_eff.upload(in.eff_lens, T, st);
  d_emi.alloc((size_t)2 * chunk); d_ch.alloc((size_t)2 * chunk); d_bar.alloc(1);
  d_rowoff.alloc((size_t)chunk + 1);
  std::vector<unsigned long long> ro((size_t)chunk + 1);
  std::vector<int> emi((size_t)2 * chunk);
  for (uint32_t s0 = 0; s0 < S; s0 += (uint32_t)chunk) {
    const int nb = (int)std::min<uint32_t>((uint32_t)chunk, S - s0);
    for (int b = 0; b <= nb; ++b) ro[b] = in.row_off[s0 + b];
    d_rowoff.upload(ro.data(), (size_t)nb + 1, st);
    if (in.per_sample_eff) d_eff.upload(in.eff_lens + (size_t)s0 * T, (size_t)nb * T, st);
    TccFill f{};
    f.n_ec = nE; f.n_targets = T; f.nb = (uint32_t)nb; f.row_off = d_rowoff.p; f.ec_ids = d_ecid.p; f.vals = d_val.p;
    f.counts = d_counts.p; f.nnz = nnz; f.m_ec = d_m_ec.p; f.m_tid = d_m_tid.p; f.t_ec = d_t_ec.p; f.t_tid = d_t_tid.p;
    f.eff = d_eff.p; f.eff_stride = in.per_sample_eff ? T : 0; f.m_w = d_mw.p; f.t_w = d_tw.p;
    launch_tcc_fill(f, st);
    KB_CK(cudaGetLastError());
    launch_fill_f64(d_alpha.p, (uint64_t)nb * T, 1.0 / T, st);
    d_emi.zero(st);
    d_ch.zero(st);
    EmProblem p{};
    p.n_ec = nE; p.n_targets = T; p.n_multi = n_multi;
    p.multi_ec = d_multi_ec.p; p.m_off = d_m_off.p; p.m_tid = d_m_tid.p; p.m_w = d_mw.p;
    p.t_off = d_t_off.p; p.t_midx = d_t_midx.p; p.t_w = d_tw.p; p.t_single = d_single.p;
    p.nb = nb; p.counts = d_counts.p; p.alpha = d_alpha.p; p.norm = d_norm.p;
    p.rounds = d_emi.p; p.fstate = d_emi.p + chunk; p.bar = d_bar.p; p.chcount = d_ch.p;
    p.cnt_row = d_cnt_row.p; p.single_cnt = d_single_cnt.p;
    p.max_iter = 10000; p.min_rounds = 50; p.w_stride = nnz;
    launch_em(p, em_tpb(), st);
    KB_CK(cudaGetLastError());
    d_alpha.download(alpha_out.data() + (size_t)s0 * T, (size_t)nb * T, 0, st);
    d_emi.download(emi.data(), (size_t)2 * chunk, 0, st);
    KB_CK(cudaStreamSynchronize(st));
    for (int b = 0; b < nb; ++b) {
      rounds[s0 + b] = emi[b];
      if (emi[chunk + b] == 3)
        for (uint32_t t = 0; t < T; ++t)
          if (alpha_out[(size_t)(s0 + b) * T + t] < 1e-7 / 10.0) alpha_out[(size_t)(s0 + b) * T + t] = 0.0;
    }
  }
  cudaStreamDestroy(st);
  return rounds;
}

template struct DBuf<uint8_t>;
template struct DBuf<uint16_t>;
template struct DBuf<uint32_t>;
template struct DBuf<int32_t>;
template struct DBuf<uint64_t>;
template struct DBuf<unsigned long long>;
template struct DBuf<double>;
template struct DBuf<KmerSlot>;
template struct DBuf<Memo2Entry>;
template struct DBuf<BusRecord>;
template struct DBuf<uint4>;


}  // namespace kb
