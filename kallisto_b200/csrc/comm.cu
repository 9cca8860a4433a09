// Multi-GPU exchange in C++ over NCCL: the multi-rank form of the merge MasterProcessor::update performs
// under writer_lock (src/ProcessReads.cpp:424-483).
//
// Reads shard across ranks (one kb::Quant per GPU, index replicated); the only exchange is at the end of
// pseudoalignment.  Every rank numbers its equivalence classes on the device (Quant::export_prepare), a
// 4-word meta record per rank is all-gathered (the only host synchronisation), ranks != root send their
// tables (CSR offsets, transcript ids, counts, first-occurrence keys, the ordered fragment-length samples)
// straight to the root with grouped ncclSend/ncclRecv -- unpadded, nobody else receives them -- and the
// root folds all of them into its set dictionary BY CONTENT with ONE launch of import_segments_kernel
// (warp per incoming set).  EC ids are discovered independently on every rank, so a dense all-reduce of
// count vectors is not possible before this merge: the content-keyed merge is the reduction.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already loaded NCCL (torch) shares
// that copy, and the library has no link-time dependency for single-GPU users.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "engine.hpp"

namespace kb {

namespace {

inline void ck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw Error(std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}
#define KB_CK(x) ck((x), #x)

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) return;
    auto sym = [&](const char* s) { return dlsym(api.lib, s); };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.Send || !api.Recv ||
      !api.GroupStart || !api.GroupEnd)
    throw Error("kallisto_b200: NCCL (libnccl.so.2) could not be loaded: multi-GPU runs need it");
  return api;
}

inline void nck(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) {
    const char* m = nccl().GetErrorString ? nccl().GetErrorString(r) : "?";
    throw Error(std::string("NCCL error in ") + what + ": " + m);
  }
}
#define KB_NCK(x) nck((x), #x)

__global__ void add_u64_kernel(unsigned long long* p, uint32_t n, unsigned long long v) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] += v;
}

constexpr uint32_t kTlCap = 10000;   // fragment-length samples a run collects at most (ProcessReads.cpp:985-1004)

}  // namespace

struct CommImpl {
  ncclComm_t comm = nullptr;
  bool owned = true;
  // root: receive area (grow-only); every rank: meta words
  DBuf<unsigned long long> d_meta;      // n_ranks x 4
  unsigned long long* h_meta = nullptr; // pinned
  DBuf<uint32_t> r_off, r_tids, r_counts;
  DBuf<unsigned long long> r_first;
  DBuf<uint16_t> r_tl, s_tl;
  uint16_t* h_tl = nullptr;             // pinned: n_ranks x kTlCap
};

void Comm::unique_id(void* out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  KB_NCK(nccl().GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
}

static void comm_buffers(CommImpl& c, int n_ranks) {
  c.d_meta.alloc((size_t)n_ranks * 4);
  KB_CK(cudaMallocHost((void**)&c.h_meta, (size_t)n_ranks * 4 * sizeof(unsigned long long)));
  KB_CK(cudaMallocHost((void**)&c.h_tl, (size_t)n_ranks * kTlCap * sizeof(uint16_t)));
  c.s_tl.alloc(kTlCap);
}

Comm::Comm(int n_ranks_, int rank_, const void* id128, int device_) : n_ranks(n_ranks_), rank(rank_), device(device_) {
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) throw Error("kallisto_b200: bad communicator shape");
  KB_CK(cudaSetDevice(device));
  impl_ = new CommImpl();
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  KB_NCK(nccl().CommInitRank(&impl_->comm, n_ranks, id, rank));
  comm_buffers(*impl_, n_ranks);
}

Comm::Comm(void* nccl_comm, int n_ranks_, int rank_, int device_, bool take_ownership)
    : n_ranks(n_ranks_), rank(rank_), device(device_) {
  if (!nccl_comm || n_ranks < 1 || rank < 0 || rank >= n_ranks) throw Error("kallisto_b200: bad communicator");
  nccl();
  KB_CK(cudaSetDevice(device));
  impl_ = new CommImpl();
  impl_->comm = (ncclComm_t)nccl_comm;
  impl_->owned = take_ownership;
  comm_buffers(*impl_, n_ranks);
}

std::vector<Comm*> Comm::init_all(const std::vector<int>& devices) {
  NcclApi& api = nccl();
  if (!api.CommInitAll) throw Error("kallisto_b200: ncclCommInitAll missing");
  std::vector<ncclComm_t> cs(devices.size());
  KB_NCK(api.CommInitAll(cs.data(), (int)devices.size(), devices.data()));
  std::vector<Comm*> out;
  for (size_t i = 0; i < devices.size(); ++i) out.push_back(new Comm((void*)cs[i], (int)devices.size(), (int)i, devices[i], true));
  return out;
}

Comm::~Comm() {
  if (!impl_) return;
  cudaSetDevice(device);
  if (impl_->comm && impl_->owned) nccl().CommDestroy(impl_->comm);
  if (impl_->h_meta) cudaFreeHost(impl_->h_meta);
  if (impl_->h_tl) cudaFreeHost(impl_->h_tl);
  delete impl_;
}

void Comm::reserve(size_t n_sets, size_t n_entries) {
  KB_CK(cudaSetDevice(device));
  CommImpl& c = *impl_;
  if (rank != 0 || n_ranks == 1) return;
  const size_t peers = (size_t)n_ranks - 1;
  if (c.r_off.n < peers * (n_sets + 1)) c.r_off.alloc(peers * (n_sets + 1));
  if (c.r_counts.n < peers * n_sets) c.r_counts.alloc(peers * n_sets);
  if (c.r_first.n < peers * n_sets) c.r_first.alloc(peers * n_sets);
  if (c.r_tids.n < peers * n_entries) c.r_tids.alloc(peers * n_entries);
  if (c.r_tl.n < peers * kTlCap) c.r_tl.alloc(peers * kTlCap);
}

// Collective over the communicator.  first_stride orders the ranks' fragment indices (rank r's first-occurrence keys
// are offset by r * first_stride: slices read one after the other); 0 when the runs were given global fragment
// indices already (Quant::set_frag_base).
uint64_t Quant::merge_to_root(Comm& cm, uint64_t first_stride) {
  KB_CK(cudaSetDevice(ix_.device));
  if (cm.device != ix_.device) throw Error("kallisto_b200: communicator and run live on different devices");
  NcclApi& api = nccl();
  CommImpl& c = *cm.impl_;
  cudaStream_t st = stream_;
  const int N = cm.n_ranks, me = cm.rank;
  uint32_t n = 0, nnz = 0;
  if (me != 0) export_prepare(&n, &nnz);          // table in emws_: ec_off, ec_tid, count, key_out (first)
  else check_device_errors();
  EmWs& w = *emws_;
  const uint32_t my_tl = (uint32_t)std::min<size_t>(tl_list_.size(), kTlCap);
  // ---- meta: {sets, entries, fragments processed, fragment-length samples} of every rank
  unsigned long long mine[4] = {n, nnz, n_frag_total_, my_tl};
  KB_CK(cudaMemcpyAsync(c.d_meta.p + (size_t)me * 4, mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  KB_CK(cudaStreamSynchronize(st));               // `mine` is a stack array
  if (N > 1) KB_NCK(api.AllGather(c.d_meta.p + (size_t)me * 4, c.d_meta.p, 4, ncclUint64, c.comm, st));
  KB_CK(cudaMemcpyAsync(c.h_meta, c.d_meta.p, (size_t)N * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  KB_CK(cudaStreamSynchronize(st));
  uint64_t total = 0;
  for (int r = 0; r < N; ++r) total += c.h_meta[r * 4 + 2];
  if (N == 1) return total;
  if (me != 0) {
    // ---- sender: first-occurrence keys carry the rank order, then one grouped send of the table
    if (n && first_stride) add_u64_kernel<<<(n + 255) / 256, 256, 0, st>>>(w.key_out.p, n, (unsigned long long)me * first_stride);
    if (my_tl) KB_CK(cudaMemcpyAsync(c.s_tl.p, tl_list_.data(), (size_t)my_tl * 2, cudaMemcpyHostToDevice, st));
    KB_NCK(api.GroupStart());
    if (n) {
      KB_NCK(api.Send(w.ec_off.p, (size_t)n + 1, ncclUint32, 0, c.comm, st));
      KB_NCK(api.Send(w.ec_tid.p, nnz, ncclUint32, 0, c.comm, st));
      KB_NCK(api.Send(w.count.p, n, ncclUint32, 0, c.comm, st));
      KB_NCK(api.Send(w.key_out.p, n, ncclUint64, 0, c.comm, st));
    }
    if (my_tl) KB_NCK(api.Send(c.s_tl.p, (size_t)my_tl * 2, ncclUint8, 0, c.comm, st));
    KB_NCK(api.GroupEnd());
    KB_CK(cudaStreamSynchronize(st));             // tl_list_ / table buffers may be reused by the caller afterwards
    return total;
  }
  // ---- root: receive every table into one area, then ONE import launch over all of them
  size_t sum_n = 0, sum_nnz = 0, max_n = 0, max_nnz = 0;
  for (int r = 1; r < N; ++r) {
    sum_n += c.h_meta[r * 4];
    sum_nnz += c.h_meta[r * 4 + 1];
    max_n = std::max<size_t>(max_n, c.h_meta[r * 4]);
    max_nnz = std::max<size_t>(max_nnz, c.h_meta[r * 4 + 1]);
  }
  if (c.r_off.n < sum_n + (size_t)N) c.r_off.alloc(sum_n + (size_t)N + sum_n / 4);
  if (c.r_counts.n < sum_n) c.r_counts.alloc(sum_n + sum_n / 4);
  if (c.r_first.n < sum_n) c.r_first.alloc(sum_n + sum_n / 4);
  if (c.r_tids.n < sum_nnz) c.r_tids.alloc(sum_nnz + sum_nnz / 4);
  if (c.r_tl.n < (size_t)(N - 1) * kTlCap) c.r_tl.alloc((size_t)(N - 1) * kTlCap);
  std::vector<ImportSeg> segs;
  KB_NCK(api.GroupStart());
  {
    size_t o_off = 0, o_n = 0, o_nnz = 0;
    for (int r = 1; r < N; ++r) {
      const size_t rn = c.h_meta[r * 4], rnnz = c.h_meta[r * 4 + 1], rtl = c.h_meta[r * 4 + 3];
      if (rn) {
        KB_NCK(api.Recv(c.r_off.p + o_off, rn + 1, ncclUint32, r, c.comm, st));
        KB_NCK(api.Recv(c.r_tids.p + o_nnz, rnnz, ncclUint32, r, c.comm, st));
        KB_NCK(api.Recv(c.r_counts.p + o_n, rn, ncclUint32, r, c.comm, st));
        KB_NCK(api.Recv(c.r_first.p + o_n, rn, ncclUint64, r, c.comm, st));
        ImportSeg s;
        s.n_sets = (uint32_t)rn;
        s.off = c.r_off.p + o_off; s.tids = c.r_tids.p + o_nnz; s.counts = c.r_counts.p + o_n; s.first = c.r_first.p + o_n;
        segs.push_back(s);
      }
      if (rtl) KB_NCK(api.Recv(c.r_tl.p + (size_t)(r - 1) * kTlCap, rtl * 2, ncclUint8, r, c.comm, st));
      o_off += rn + 1; o_n += rn; o_nnz += rnnz;
    }
  }
  KB_NCK(api.GroupEnd());
  ecs_valid_ = false;
  dev_stats_valid_ = false;
  dev_problem_valid_ = false;
  for (size_t i = 0; i < segs.size(); i += KB_IMPORT_SEGS) {
    launch_import_segments(dd_, segs.data() + i, (int)std::min<size_t>(KB_IMPORT_SEGS, segs.size() - i), st);
    KB_CK(cudaGetLastError());
    ++n_kernel_launches;
  }
  // fragment-length distribution: the first 10000 qualifying pairs in read order = this rank's, then rank 1's, ...
  // (src/ProcessReads.cpp:985-1004,1174-1181); only needed when the root's own slice did not fill the quota
  if (opt_.paired && opt_.collect_fld && tlencount_ < kTlCap) {
    for (int r = 1; r < N; ++r) {
      const size_t rtl = c.h_meta[r * 4 + 3];
      if (rtl) KB_CK(cudaMemcpyAsync(c.h_tl + (size_t)r * kTlCap, c.r_tl.p + (size_t)(r - 1) * kTlCap, rtl * 2, cudaMemcpyDeviceToHost, st));
    }
    KB_CK(cudaStreamSynchronize(st));
    for (int r = 1; r < N && tlencount_ < kTlCap; ++r) {
      const size_t rtl = c.h_meta[r * 4 + 3];
      for (size_t i = 0; i < rtl && tlencount_ < kTlCap; ++i) {
        const uint16_t tl = c.h_tl[(size_t)r * kTlCap + i];
        ++flens_[tl];
        tl_list_.push_back(tl);
        ++tlencount_;
      }
    }
  }
  n_frag_total_ = total;
  return total;
}

}  // namespace kb
