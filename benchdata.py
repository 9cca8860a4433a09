"""Seeded synthetic transcriptomes and reads (bench + test infrastructure, not product code).

There is no network and no GENCODE FASTA in this environment, so BASELINE config 2 ("human
GENCODE-v44 txome") is stood in for by a transcriptome with GENCODE-like statistics, as
SURVEY.md section 8(d) prescribes: genes made of 5-15 exons of log-normal length (median 150 nt),
isoforms = ordered exon subsets (>= 2 exons) + a 3' UTR, lengths clipped to [200, 100000].
Reads: 2 x L bp, fragment length ~ N(200, 30^2) truncated to [L, 999], transcript drawn
proportionally to Gamma(0.5) abundance x effective length, mate 2 reverse-complemented, each fragment
flipped with p = 0.5 (unstranded), 0.5 % substitutions, 0.1 % of reads carry one N, 5 % of pairs
are random sequence.  Everything is a pure function of the seed.
"""
import os

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    COMP[a] = b


class Transcriptome:
    def __init__(self, seqs, names):
        self.seqs = seqs            # list of uint8 arrays (ASCII)
        self.names = names
        self.lens = np.array([len(s) for s in seqs], np.int64)
        self.concat = np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)
        self.starts = np.zeros(len(seqs) + 1, np.int64)
        np.cumsum(self.lens, out=self.starts[1:])

    def write_fasta(self, path, width=0):
        with open(path, "wb") as f:
            for n, s in zip(self.names, self.seqs):
                f.write(b">" + n.encode() + b"\n")
                f.write(s.tobytes())
                f.write(b"\n")


def make_transcriptome(n_genes, seed=44, iso_mean=4.06, max_len=100000):
    """~iso_mean isoforms per gene (GENCODE v44: 252k transcripts / 62k genes)."""
    rng = np.random.default_rng(seed)
    seqs, names = [], []
    for g in range(n_genes):
        n_ex = int(rng.integers(5, 16))
        ex_len = np.clip(np.exp(rng.normal(np.log(150.0), 0.6, n_ex)).astype(np.int64), 30, 5000)
        exons = [ACGT[rng.integers(0, 4, int(l))] for l in ex_len]
        utr = ACGT[rng.integers(0, 4, int(np.clip(np.exp(rng.normal(np.log(400.0), 0.7)), 50, 5000)))]
        n_iso = 1 + int(rng.poisson(iso_mean - 1.0))
        seen = set()
        for i in range(n_iso):
            for _ in range(8):
                keep = rng.random(n_ex) < 0.7
                if keep.sum() >= 2 and keep.tobytes() not in seen:
                    break
            else:
                continue
            if keep.sum() < 2 or keep.tobytes() in seen:
                continue
            seen.add(keep.tobytes())
            s = np.concatenate([e for e, k in zip(exons, keep) if k] + [utr])
            if len(s) < 200:
                continue
            seqs.append(s[:max_len])
            names.append("SYNT%06d.%d" % (g, i + 1))
    return Transcriptome(seqs, names)


def simulate_pairs(tx, n_pairs, read_len=100, seed=20240601, err=0.005, n_frac=0.001, random_frac=0.05,
                   fl_mean=200.0, fl_sd=30.0, abundance_seed=7, stranded=False):
    """-> (r1, r2) uint8 arrays of shape (n_pairs, read_len), ASCII."""
    rng = np.random.default_rng(seed)
    T = len(tx.lens)
    ab = np.random.default_rng(abundance_seed).gamma(0.5, 1.0, T)
    w = ab * np.maximum(tx.lens - fl_mean + 1, 1.0)
    w /= w.sum()
    t = rng.choice(T, size=n_pairs, p=w)
    fl = np.clip(np.rint(rng.normal(fl_mean, fl_sd, n_pairs)), read_len, 999).astype(np.int64)
    fl = np.minimum(fl, tx.lens[t])
    ok = fl >= read_len
    start = (rng.random(n_pairs) * (tx.lens[t] - fl + 1)).astype(np.int64)
    base = tx.starts[t] + start
    ar = np.arange(read_len)
    r1 = np.empty((n_pairs, read_len), np.uint8)
    r2 = np.empty((n_pairs, read_len), np.uint8)
    idx1 = np.where(ok, base, 0)[:, None] + ar[None, :]
    r1[:] = tx.concat[np.minimum(idx1, len(tx.concat) - 1)]
    idx2 = np.where(ok, base + fl - 1, read_len - 1)[:, None] - ar[None, :]
    r2[:] = COMP[tx.concat[np.clip(idx2, 0, len(tx.concat) - 1)]]
    if not stranded:
        flip = rng.random(n_pairs) < 0.5
        r1[flip], r2[flip] = r2[flip], r1[flip].copy()
    # random (unmappable) pairs
    rnd = (rng.random(n_pairs) < random_frac) | ~ok
    nr = int(rnd.sum())
    r1[rnd] = ACGT[rng.integers(0, 4, (nr, read_len))]
    r2[rnd] = ACGT[rng.integers(0, 4, (nr, read_len))]
    # substitutions
    for r in (r1, r2):
        m = rng.random(r.shape) < err
        sub = ACGT[rng.integers(0, 4, int(m.sum()))]
        r[m] = sub
        nn = rng.random(n_pairs) < n_frac
        r[nn, rng.integers(0, read_len, int(nn.sum()))] = ord("N")
    return r1, r2


def write_fastq(path, reads, tag, first_index=0):
    """Uncompressed FASTQ with @r<i>/<tag> names and constant quality 'I'."""
    n, L = reads.shape
    qual = b"I" * L
    with open(path, "wb") as f:
        CH = 200000
        for c0 in range(0, n, CH):
            c1 = min(n, c0 + CH)
            parts = []
            blk = reads[c0:c1]
            for i in range(c1 - c0):
                parts.append(b"@r%d/%d\n" % (first_index + c0 + i, tag))
                parts.append(blk[i].tobytes())
                parts.append(b"\n+\n")
                parts.append(qual)
                parts.append(b"\n")
            f.write(b"".join(parts))


def interleave(r1, r2):
    """-> bases (uint8, flat) for kb_pseudoalign_batch with fixed_len = read_len."""
    n, L = r1.shape
    out = np.empty((n, 2, L), np.uint8)
    out[:, 0, :] = r1
    out[:, 1, :] = r2
    return out.reshape(-1)
