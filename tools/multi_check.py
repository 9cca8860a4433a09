#!/usr/bin/env python
"""N-rank check of the NCCL merge (csrc/comm.cu), run under torchrun on N GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multi_check.py

Every rank pseudoaligns its contiguous slice of the synth_small reads, kb_quant_merge_nccl folds the slices into
rank 0, and rank 0 compares with ONE run over all reads: EC sets in first-occurrence order, counts, the
fragment-length histogram (completed in rank order) and the EM, bit for bit.  Also runs the command line with
--devices 0..N-1 against --device 0 (abundance.tsv must be byte-identical)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import kallisto_b200 as K  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import util  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    torch.cuda.set_device(lr)
    uid = [K.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = K.Comm(world, rank, uid[0], lr)
    ok = True
    for name in ("synth_small", "config1"):
        ds = util.dataset(name)
        n = len(ds["s1"])
        # ragged slices; rank 0's is tiny so that the fragment-length samples have to be completed from the other ranks
        cuts = [0, 300] + [300 + (n - 300) * i // (world - 1) for i in range(1, world)] if world > 1 else [0, n]
        lo, hi = cuts[rank], cuts[rank + 1]
        ix = K.KmerIndex(ds["index"], device=lr)
        mc = K.MinCollector(ix, paired=True)
        if hi > lo:
            mc.process_buffer(*O.to_batch(ds["s1"][lo:hi], ds["s2"][lo:hi]), want_handles=False)
        total = mc.merge_nccl(comm)
        if rank == 0:
            whole = K.MinCollector(ix, paired=True)
            whole.process_buffer(*util.batch(ds, True), want_handles=False)
            wo, wt, wc, _ = whole.ec_table()
            ao, at, ac, _ = mc.ec_table()
            same_sets = util.ec_sets(ao, at) == util.ec_sets(wo, wt)
            same_counts = np.array_equal(ac, wc)
            same_fl = np.array_equal(mc.flens, whole.flens)
            ra, rw = mc.run_em(), whole.run_em()
            same_em = ra["rounds"] == rw["rounds"] and np.array_equal(ra["est_counts"], rw["est_counts"])
            st = mc.finalize()
            print("[multi_check] %s world=%d: total=%d (want %d) sets=%s counts=%s flens=%s em=%s n_processed=%d" % (
                name, world, total, n, same_sets, same_counts, same_fl, same_em, st["n_processed"]), flush=True)
            ok = ok and same_sets and same_counts and same_fl and same_em and total == n and st["n_processed"] == n
            whole.close()
        mc.close()
        ix.close()
    dist.barrier()
    if rank == 0:
        # command line: --devices 0..N-1 vs one device
        ds = util.dataset("synth_small")
        exe = os.path.join(ROOT, "kallisto_b200", "kallisto_b200")
        with tempfile.TemporaryDirectory() as td:
            outs = []
            for tag, dv in (("one", ["--device", "0"]), ("all", ["--devices", ",".join(str(i) for i in range(world))])):
                out = os.path.join(td, tag)
                env = dict(os.environ, KB_CLI_BATCH_READS="1500")
                r = subprocess.run([exe, "quant", "-i", ds["index"], "-o", out, "--plaintext", "-b", "2", "-t", "4"] + dv +
                                   [os.path.join(ds["dir"], "reads_1.fastq.gz"), os.path.join(ds["dir"], "reads_2.fastq.gz")],
                                   capture_output=True, text=True, env=env)
                if r.returncode != 0:
                    print("[multi_check] CLI %s failed: %s" % (tag, r.stderr[-800:]), flush=True)
                    ok = False
                    break
                outs.append([open(os.path.join(out, f)).read() for f in ("abundance.tsv", "bs_abundance_0.tsv", "bs_abundance_1.tsv")])
            if len(outs) == 2:
                ref = open(os.path.join(ds["dir"], "ref_quant_paired", "abundance.tsv")).read()
                same = outs[0] == outs[1] and outs[0][0] == ref
                print("[multi_check] CLI --devices 0..%d == --device 0 == reference: %s" % (world - 1, same), flush=True)
                ok = ok and same
        print("[multi_check] %s" % ("ALL OK" if ok else "FAILED"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
