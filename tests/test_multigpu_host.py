"""CPU, world_size 2 and 4 over gloo: the EC-table exchange of kallisto_b200/multigpu.py (the only
collective step of the multi-GPU path).  The GPU-side merge kernel is covered by tests/test_gpu_multi.py;
here the protocol itself (sizes, padding, ordering) is checked against a single-process merge."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kallisto_b200 import multigpu


def make_table(seed, n):
    rng = np.random.default_rng(seed)
    sets = []
    for i in range(n):
        k = int(rng.integers(1, 6))
        sets.append(tuple(sorted(rng.choice(50, size=k, replace=False).tolist())))
    sets = list(dict.fromkeys(sets))
    counts = rng.integers(1, 100, len(sets))
    first = np.sort(rng.choice(10 ** 6, len(sets), replace=False))
    off = np.zeros(len(sets) + 1, np.int32)
    tids = []
    for i, s in enumerate(sets):
        tids.extend(s)
        off[i + 1] = len(tids)
    return dict(off=off, tids=np.array(tids, np.int32), counts=counts.astype(np.int32), first=first.astype(np.int64), sets=sets)


def merged_reference(tables):
    """content-keyed merge, ranks in order: what rank 0 must hold afterwards"""
    acc = {}
    for r, t in enumerate(tables):
        for i, s in enumerate(t["sets"]):
            f = int(t["first"][i]) + r * multigpu.RANK_STRIDE
            c = int(t["counts"][i])
            if s in acc:
                acc[s] = (acc[s][0] + c, min(acc[s][1], f))
            else:
                acc[s] = (c, f)
    return acc


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = make_table(100 + rank, 40 + 25 * rank)
    tabs = multigpu.all_gather_tables(torch.from_numpy(t["off"]), torch.from_numpy(t["tids"]), torch.from_numpy(t["counts"]),
                                      torch.from_numpy(t["first"]))
    # every rank reconstructs every table exactly
    ok = True
    acc = {}
    for r in range(world):
        ref = make_table(100 + r, 40 + 25 * r)
        g = tabs[r]
        ok &= g["n"] == len(ref["sets"]) and g["m"] == len(ref["tids"])
        ok &= np.array_equal(g["off"].numpy(), ref["off"]) and np.array_equal(g["tids"].numpy(), ref["tids"])
        ok &= np.array_equal(g["counts"].numpy(), ref["counts"]) and np.array_equal(g["first"].numpy(), ref["first"])
        off = g["off"].numpy()
        for i in range(g["n"]):
            s = tuple(g["tids"].numpy()[off[i]:off[i + 1]].tolist())
            f = int(g["first"][i]) + r * multigpu.RANK_STRIDE
            c = int(g["counts"][i])
            acc[s] = (acc[s][0] + c, min(acc[s][1], f)) if s in acc else (c, f)
    ok &= acc == merged_reference([make_table(100 + r, 40 + 25 * r) for r in range(world)])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_table_exchange(world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
