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


# ---- paired / sample-per-file BUS runs (tests/golden/buspaired) -------------------------------------------------------
# The inputs are derived from synth_small's reads by this one function, used by tests/golden/make_golden.py (which runs
# the unmodified reference on them) and by the tests (which run this build on them), so only the outputs are stored.
BUSPAIRED_CASES = {
    # name: (arguments after `bus -i IDX -o OUT -t T`, input files by key)
    "bulk_paired": (["-x", "bulk", "--paired"], ["a_1", "a_2", "b_1", "b_2"]),
    "bulk_paired_num_fr": (["-x", "bulk", "--paired", "--num", "--fr-stranded"], ["a_1", "a_2", "b_1", "b_2"]),
    "bulk_single": (["-x", "bulk"], ["a_1", "b_1", "a_2"]),
    "smartseq2_paired": (["-x", "smartseq2", "--paired"], ["i_1", "i_2", "s_1", "s_2"]),
    "smartseq2_single_rf": (["-x", "smartseq2", "--rf-stranded"], ["i_1", "i_2", "s_1"]),
    "stormlike": (["-x", "-1,-1,-1:1,0,8:0,0,0,1,14,0", "--paired", "--rf-stranded"], ["s_1", "u_2"]),
    "smartseq3": (["-x", "smartseq3"], ["i_1", "i_2", "t_1", "s_2"]),
    "tag_single_fr": (["-x", "0,0,8:1,0,19:1,22,0", "--tag", "ATTGCGCAATG", "--fr-stranded"], ["i_1", "t_1"]),
}
SMARTSEQ3_TAG = b"ATTGCGCAATG"
# `bus --batch FILE` (src/main.cpp:1108-1180): lines "id file1 file2"; lines with the same id share a sample barcode
BATCHFILE_LINES = [("#id", "file1", "file2"), ("cellA", "a_1", "a_2"), ("cellB", "b_1", "b_2"), ("cellA", "s_1", "s_2")]


def write_batch_file(path, inputs):
    with open(path, "w") as f:
        for i, k1, k2 in BATCHFILE_LINES:
            f.write("%s %s %s\n" % (i, inputs.get(k1, k1), inputs.get(k2, k2)))
        f.write("\n")
    return path


def buspaired_inputs(dst):
    """Writes the FASTQ files of the buspaired cases into `dst` -> {key: path}.
    a_*/b_*: synth_small's pairs 0..11999 / 12000..19999 (two samples of `bus -x bulk`);
    s_*: pairs 0..7999; i_1/i_2: index reads for them (SMARTSEQ2: the barcode is both index reads, whole: 6 cells,
    some reads with an N, some 6 instead of 8 nt, three empty ones -- an empty barcode piece skips the read set,
    src/ProcessReads.cpp:1592-1598); u_2: the second mates with a 14-nt prefix whose first 8 nt are the UMI
    (the STORM-seq layout, src/main.cpp:1358-1365), some with an N in the UMI, some shorter than the UMI."""
    ds = dataset("synth_small")
    s1, s2 = ds["s1"], ds["s2"]
    rng = np.random.default_rng(20260923)
    A = np.frombuffer(b"ACGT", np.uint8)

    def write(key, seqs):
        p = os.path.join(dst, key + ".fq")
        with open(p, "wb") as f:
            for i, s in enumerate(seqs):
                if not isinstance(s, bytes):
                    s = s.encode()
                f.write(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)))
        return p

    out = {}
    out["a_1"] = write("a_1", s1[:12000]); out["a_2"] = write("a_2", s2[:12000])
    out["b_1"] = write("b_1", s1[12000:]); out["b_2"] = write("b_2", s2[12000:])
    n = 8000
    out["s_1"] = write("s_1", s1[:n]); out["s_2"] = write("s_2", s2[:n])
    cells = A[rng.integers(0, 4, (6, 2, 8))]
    which = rng.integers(0, 6, n)
    i1 = [bytes(cells[c, 0]) for c in which]
    i2 = [bytes(cells[c, 1]) for c in which]
    for j in rng.choice(n, 80, replace=False):
        b = bytearray(i1[j]); b[int(rng.integers(0, 8))] = ord("N"); i1[j] = bytes(b)
    for j in rng.choice(n, 40, replace=False):
        i2[j] = i2[j][:6]
    for j in rng.choice(n, 3, replace=False):
        i1[j] = b""
    out["i_1"] = write("i_1", i1); out["i_2"] = write("i_2", i2)
    pre = A[rng.integers(0, 4, (n, 14))]
    u2 = []
    for j in range(n):
        m = s2[j] if isinstance(s2[j], bytes) else s2[j].encode()
        u2.append(bytes(pre[j]) + m)
    for j in rng.choice(n, 60, replace=False):
        b = bytearray(u2[j]); b[int(rng.integers(0, 8))] = ord("N"); u2[j] = bytes(b)
    for j in rng.choice(n, 5, replace=False):
        u2[j] = u2[j][:int(rng.integers(1, 8))]
    out["u_2"] = write("u_2", u2)
    # t_1: first mates of a SMARTSEQ3 run: 40 % UMI reads = tag (11 nt) + UMI (8) + GGG + cDNA, a few of them with one
    # wrong letter in the tag (still a UMI read: one mismatch is allowed, src/ProcessReads.cpp:1517) or with two Ns (an
    # internal read then, like the plain 60 %)
    kinds = rng.random(n)
    t1 = []
    for j in range(n):
        r = s1[j] if isinstance(s1[j], bytes) else s1[j].encode()
        if kinds[j] < 0.4:
            tag = bytearray(SMARTSEQ3_TAG)
            if kinds[j] < 0.05:
                q = int(rng.integers(0, 11))
                tag[q] = ord("C") if tag[q] != ord("C") else ord("A")
            elif kinds[j] < 0.08:
                for q in rng.choice(11, 2, replace=False):
                    tag[int(q)] = ord("N")
            r = bytes(tag) + bytes(A[rng.integers(0, 4, 8)]) + b"GGG" + r
        t1.append(r)
    for j in rng.choice(n, 4, replace=False):
        t1[j] = t1[j][:int(rng.integers(5, 19))]          # shorter than tag + UMI: the set is skipped
    out["t_1"] = write("t_1", t1)
    return out
