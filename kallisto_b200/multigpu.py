"""Multi-GPU exchange of equivalence-class tables over torch.distributed -- TEST TRANSPORT.

The product path is the C++/NCCL merge of csrc/comm.cu (kb_quant_merge_nccl: what `kallisto_b200 quant --devices`
and bench.py call).  This module keeps the same export -> transport -> import-by-content scheme on top of
torch.distributed so that the exchange logic can be exercised on a CPU box over gloo (tests/test_multigpu_host.py)
and the export / import kernels on one GPU (tests/test_gpu_multi.py).

Reads shard naturally: every rank pseudoaligns its own contiguous slice of the input against a
replicated index and ends up with its own set dictionary.  There is exactly one exchange step, at
the end: every rank numbers its equivalence classes, the tables (CSR of transcript ids, counts,
first-occurrence indices) are all-gathered (NCCL over NVLink; a few MB per rank), and rank 0 folds
the other ranks' tables into its dictionary BY CONTENT on the GPU (import_sets_kernel) before the
single EM.  A dense all-reduce of count vectors is not possible before that merge, because EC ids
are discovered independently on every rank; the content-keyed merge is the reduction.

The exchange itself is backend-agnostic (it only uses all_gather on flat tensors), so the CPU test
suite runs it over gloo with world_size 2.
"""
import torch
import torch.distributed as dist

RANK_STRIDE = 1 << 40   # rank r's fragment indices are offset by r * RANK_STRIDE: ranks own consecutive slices


def all_gather_tables(off, tids, counts, first, group=None):
    """Every rank passes its table as flat tensors (int32 off[n+1], int32 tids[m], int32 counts[n],
    int64 first[n]); every rank gets the list of all ranks' tables back (padded all_gather)."""
    world = dist.get_world_size(group)
    dev = off.device
    sizes = torch.tensor([counts.numel(), tids.numel()], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    ns = [int(s[0]) for s in all_sizes]
    ms = [int(s[1]) for s in all_sizes]
    nmax, mmax = max(ns), max(ms)

    def gather(t, length, dtype):
        pad = torch.zeros(length, dtype=dtype, device=dev)
        pad[: t.numel()] = t
        out = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(out, pad, group=group)
        return out

    g_off = gather(off, nmax + 1, torch.int32)
    g_tid = gather(tids, max(1, mmax), torch.int32)
    g_cnt = gather(counts, max(1, nmax), torch.int32)
    g_fst = gather(first, max(1, nmax), torch.int64)
    return [dict(n=ns[r], m=ms[r], off=g_off[r][: ns[r] + 1], tids=g_tid[r][: ms[r]], counts=g_cnt[r][: ns[r]],
                 first=g_fst[r][: ns[r]]) for r in range(world)]


def export_table(mc, device):
    """Number this rank's ECs on the device and return them as flat torch tensors on `device`."""
    n, m = mc.export_prepare()
    off = torch.zeros(n + 1, dtype=torch.int32, device=device)
    tids = torch.zeros(max(1, m), dtype=torch.int32, device=device)
    counts = torch.zeros(max(1, n), dtype=torch.int32, device=device)
    first = torch.zeros(max(1, n), dtype=torch.int64, device=device)
    mc.export_device(off.data_ptr(), tids.data_ptr(), counts.data_ptr(), first.data_ptr())
    return off, tids[:m], counts[:n], first[:n]


def merge_on_rank0(mc, n_processed_local, device, group=None):
    """Collective.  After it, rank 0's run holds the equivalence classes of all ranks (ids ordered as if
    the ranks' slices had been read one after the other); returns the global number of fragments."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    off, tids, counts, first = export_table(mc, device)
    tables = all_gather_tables(off, tids, counts, first, group)
    nproc = torch.tensor([n_processed_local], dtype=torch.int64, device=device)
    all_np = [torch.zeros_like(nproc) for _ in range(world)]
    dist.all_gather(all_np, nproc, group=group)
    if rank == 0:
        for r in range(1, world):
            t = tables[r]
            if t["n"]:
                mc.import_device(t["n"], t["off"].data_ptr(), t["tids"].data_ptr(), t["counts"].data_ptr(),
                                 t["first"].data_ptr(), r * RANK_STRIDE, int(all_np[r][0]))
    return int(sum(int(x[0]) for x in all_np))
