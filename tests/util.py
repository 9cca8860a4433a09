"""Shared helpers for the parity tests."""
import os

import numpy as np

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

MODES = {
    # name: (paired, strand, golden npz)
    "paired": (True, 0, "ecs_paired.npz"),
    "paired_fr": (True, 1, "ecs_paired_fr.npz"),
    "paired_rf": (True, 2, "ecs_paired_rf.npz"),
    "single": (False, 0, "ecs_single.npz"),
    "single_fr": (False, 1, "ecs_single_fr.npz"),
    "single_rf": (False, 2, "ecs_single_rf.npz"),
}

_cache = {}


def dataset(name):
    """-> dict(index path, s1, s2 lists of read sequences)"""
    if name not in _cache:
        d = os.path.join(GOLDEN, name)
        _cache[name] = dict(dir=d, index=os.path.join(d, "transcripts.kidx"),
                            s1=O.read_fastq(os.path.join(d, "reads_1.fastq.gz")),
                            s2=O.read_fastq(os.path.join(d, "reads_2.fastq.gz")))
    return _cache[name]


def batch(ds, paired):
    return O.to_batch(ds["s1"], ds["s2"] if paired else None)


def golden_ecs(ds, mode):
    return np.load(os.path.join(ds["dir"], MODES[mode][2]))


def ec_sets(off, tids):
    return [tuple(int(x) for x in tids[int(off[i]):int(off[i + 1])]) for i in range(len(off) - 1)]


def read_abundance(path):
    names, lens, eff, est, tpm = [], [], [], [], []
    with open(path) as f:
        next(f)
        for line in f:
            a = line.rstrip("\n").split("\t")
            names.append(a[0]); lens.append(int(a[1])); eff.append(float(a[2])); est.append(float(a[3])); tpm.append(float(a[4]))
    return names, np.array(lens), np.array(eff), np.array(est), np.array(tpm)


def handles_to_ids(frag_handles, ec_handles):
    """Translate per-fragment device handles into EC ids (order of first occurrence)."""
    m = {int(h): i for i, h in enumerate(ec_handles)}
    return np.array([m[int(h)] if h >= 0 else -1 for h in frag_handles], np.int32)
