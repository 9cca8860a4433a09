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


# ---------------------------------------------------------------------------------------------
# Same read model on the GPU with torch ops (bench plumbing: inputs must already be resident in
# HBM when the timed region starts, and 30 M pairs are too slow to simulate with numpy).
# ---------------------------------------------------------------------------------------------
def _abundance_weights(lens, fl_mean=200.0, abundance_seed=7):
    ab = np.random.default_rng(abundance_seed).gamma(0.5, 1.0, len(lens))
    w = ab * np.maximum(lens - fl_mean + 1, 1.0)
    return w / w.sum()


class TorchSimulator:
    def __init__(self, concat, lens, device, read_len=100, err=0.005, n_frac=0.001, random_frac=0.05, fl_mean=200.0,
                 fl_sd=30.0):
        import torch
        self.torch = torch
        self.dev = device
        self.L = read_len
        self.err, self.n_frac, self.random_frac, self.fl_mean, self.fl_sd = err, n_frac, random_frac, fl_mean, fl_sd
        self.concat = torch.from_numpy(np.ascontiguousarray(concat)).to(device)
        lens = np.asarray(lens, np.int64)
        starts = np.zeros(len(lens), np.int64)
        np.cumsum(lens[:-1], out=starts[1:])
        self.lens = torch.from_numpy(lens).to(device)
        self.starts = torch.from_numpy(starts).to(device)
        w = _abundance_weights(lens, fl_mean)
        self.cdf = torch.from_numpy(np.cumsum(w)).to(device)
        self.comp = torch.from_numpy(COMP).to(device)
        self.acgt = torch.from_numpy(ACGT.copy()).to(device)

    def pairs(self, n_pairs, seed):
        """-> uint8 tensor (n_pairs, 2, L) on the device: mates interleaved, ready for
        kb_pseudoalign_batch(_device) with fixed_len = L."""
        torch = self.torch
        g = torch.Generator(device=self.dev)
        g.manual_seed(int(seed))
        L = self.L
        u = torch.rand(n_pairs, generator=g, device=self.dev, dtype=torch.float64)
        t = torch.searchsorted(self.cdf, u).clamp_(max=len(self.lens) - 1)
        fl = torch.normal(self.fl_mean, self.fl_sd, (n_pairs,), generator=g, device=self.dev).round_().clamp_(L, 999).long()
        tl = self.lens[t]
        fl = torch.minimum(fl, tl)
        ok = fl >= L
        start = (torch.rand(n_pairs, generator=g, device=self.dev, dtype=torch.float64) * (tl - fl + 1).double()).long()
        base = self.starts[t] + start
        ar = torch.arange(L, device=self.dev)
        nmax = self.concat.numel() - 1
        out = torch.empty((n_pairs, 2, L), dtype=torch.uint8, device=self.dev)
        CH = 1 << 20
        for c0 in range(0, n_pairs, CH):
            c1 = min(n_pairs, c0 + CH)
            b = torch.where(ok[c0:c1], base[c0:c1], torch.zeros_like(base[c0:c1]))
            f = torch.where(ok[c0:c1], fl[c0:c1], torch.full_like(fl[c0:c1], L))
            r1 = self.concat[(b[:, None] + ar[None, :]).clamp_(max=nmax)]
            r2 = self.comp[self.concat[((b + f - 1)[:, None] - ar[None, :]).clamp_(0, nmax)].long()]
            flip = torch.rand(c1 - c0, generator=g, device=self.dev) < 0.5
            out[c0:c1, 0] = torch.where(flip[:, None], r2, r1)
            out[c0:c1, 1] = torch.where(flip[:, None], r1, r2)
            rnd = (torch.rand(c1 - c0, generator=g, device=self.dev) < self.random_frac) | ~ok[c0:c1]
            rnd_bases = self.acgt[torch.randint(0, 4, (c1 - c0, 2, L), generator=g, device=self.dev)]
            out[c0:c1] = torch.where(rnd[:, None, None], rnd_bases, out[c0:c1])
            m = torch.rand((c1 - c0, 2, L), generator=g, device=self.dev) < self.err
            sub = self.acgt[torch.randint(0, 4, (c1 - c0, 2, L), generator=g, device=self.dev)]
            out[c0:c1] = torch.where(m, sub, out[c0:c1])
            nn = torch.rand((c1 - c0, 2), generator=g, device=self.dev) < self.n_frac
            pos = torch.randint(0, L, (c1 - c0, 2), generator=g, device=self.dev)
            nmask = torch.zeros((c1 - c0, 2, L), dtype=torch.bool, device=self.dev)
            nmask.scatter_(2, pos[:, :, None], nn[:, :, None])
            out[c0:c1] = torch.where(nmask, torch.full_like(out[c0:c1], ord("N")), out[c0:c1])
        return out


def write_fastq_fast(path, reads, tag, append=False):
    """Vectorised FASTQ writer: fixed-width names @r%09d/<tag>, constant quality 'I'."""
    n, L = reads.shape
    name_w = 1 + 1 + 9 + 2   # '@' 'r' digits '/' tag
    row = name_w + 1 + L + 1 + 2 + L + 1
    with open(path, "ab" if append else "wb") as f:
        CH = 500000
        for c0 in range(0, n, CH):
            c1 = min(n, c0 + CH)
            m = c1 - c0
            buf = np.empty((m, row), np.uint8)
            buf[:, 0] = ord("@")
            buf[:, 1] = ord("r")
            idx = np.arange(c0, c1, dtype=np.int64)
            for d in range(9):
                buf[:, 2 + d] = ord("0") + (idx // 10 ** (8 - d)) % 10
            buf[:, 11] = ord("/")
            buf[:, 12] = ord("0") + tag
            buf[:, 13] = ord("\n")
            buf[:, 14:14 + L] = reads[c0:c1]
            o = 14 + L
            buf[:, o] = ord("\n")
            buf[:, o + 1] = ord("+")
            buf[:, o + 2] = ord("\n")
            buf[:, o + 3:o + 3 + L] = ord("I")
            buf[:, o + 3 + L] = ord("\n")
            f.write(buf.tobytes())


def fastq_image(reads, tag, first_index):
    """FASTQ text of `reads` ((n, L) uint8 torch tensor, any device) as an (n, row) uint8 tensor on the same device:
    fixed-width names @r%09d/<tag>, constant quality 'I' -- the layout of write_fastq_fast, built with torch ops
    so that 10^7 reads take a second on the GPU."""
    import torch
    n, L = reads.shape
    dev = reads.device
    row = 14 + L + 3 + L + 1
    buf = torch.empty((n, row), dtype=torch.uint8, device=dev)
    buf[:, 0] = ord("@")
    buf[:, 1] = ord("r")
    idx = torch.arange(first_index, first_index + n, dtype=torch.int64, device=dev)
    for d in range(9):
        buf[:, 2 + d] = (48 + (idx // 10 ** (8 - d)) % 10).to(torch.uint8)
    buf[:, 11] = ord("/")
    buf[:, 12] = ord("0") + tag
    buf[:, 13] = ord("\n")
    buf[:, 14:14 + L] = reads
    o = 14 + L
    buf[:, o] = ord("\n")
    buf[:, o + 1] = ord("+")
    buf[:, o + 2] = ord("\n")
    buf[:, o + 3:o + 3 + L] = ord("I")
    buf[:, o + 3 + L] = ord("\n")
    return buf


class TorchSimulator10x:
    """BASELINE config 3 reads: R1 = 16-nt barcode from a seeded whitelist of 6000 cells + 12-nt
    random UMI; R2 = 91-nt cDNA from the last 400 nt of a transcript, sense strand (kept by the technology's
    default --fr-stranded), transcript drawn from the same abundance model as the quant reads; 0.5 % substitutions,
    0.1 % of reads carry one N, 5 % random sequence."""

    def __init__(self, concat, lens, device, cdna_len=91, n_cells=6000, err=0.005, n_frac=0.001, random_frac=0.05):
        import torch
        self.torch = torch
        self.dev = device
        self.L = cdna_len
        self.err, self.n_frac, self.random_frac = err, n_frac, random_frac
        self.concat = torch.from_numpy(np.ascontiguousarray(concat)).to(device)
        lens = np.asarray(lens, np.int64)
        starts = np.zeros(len(lens), np.int64)
        np.cumsum(lens[:-1], out=starts[1:])
        self.lens = torch.from_numpy(lens).to(device)
        self.starts = torch.from_numpy(starts).to(device)
        self.cdf = torch.from_numpy(np.cumsum(_abundance_weights(lens, 200.0))).to(device)
        self.acgt = torch.from_numpy(ACGT.copy()).to(device)
        wl = np.random.default_rng(6000).integers(0, 4, (n_cells, 16))
        self.whitelist = torch.from_numpy(ACGT[wl]).to(device)

    def sets(self, n, seed):
        """-> (r1 (n, 28) uint8, r2 (n, L) uint8) on the device."""
        torch = self.torch
        g = torch.Generator(device=self.dev)
        g.manual_seed(int(seed))
        L = self.L
        u = torch.rand(n, generator=g, device=self.dev, dtype=torch.float64)
        t = torch.searchsorted(self.cdf, u).clamp_(max=len(self.lens) - 1)
        tl = self.lens[t]
        lo = (tl - 400).clamp_(min=0)
        span = (tl - L - lo + 1).clamp_(min=1)
        start = lo + (torch.rand(n, generator=g, device=self.dev, dtype=torch.float64) * span.double()).long()
        base = self.starts[t] + start
        ar = torch.arange(L, device=self.dev)
        r2 = torch.empty((n, L), dtype=torch.uint8, device=self.dev)
        CH = 1 << 21
        nmax = self.concat.numel() - 1
        for c0 in range(0, n, CH):
            c1 = min(n, c0 + CH)
            x = self.concat[(base[c0:c1, None] + ar[None, :]).clamp_(max=nmax)]
            rnd = torch.rand(c1 - c0, generator=g, device=self.dev) < self.random_frac
            x = torch.where(rnd[:, None], self.acgt[torch.randint(0, 4, (c1 - c0, L), generator=g, device=self.dev)], x)
            m = torch.rand((c1 - c0, L), generator=g, device=self.dev) < self.err
            x = torch.where(m, self.acgt[torch.randint(0, 4, (c1 - c0, L), generator=g, device=self.dev)], x)
            nn = torch.rand(c1 - c0, generator=g, device=self.dev) < self.n_frac
            pos = torch.randint(0, L, (c1 - c0,), generator=g, device=self.dev)
            nmask = torch.zeros((c1 - c0, L), dtype=torch.bool, device=self.dev)
            nmask.scatter_(1, pos[:, None], nn[:, None])
            r2[c0:c1] = torch.where(nmask, torch.full_like(x, ord("N")), x)
        bc = self.whitelist[torch.randint(0, self.whitelist.shape[0], (n,), generator=g, device=self.dev)]
        umi = self.acgt[torch.randint(0, 4, (n, 12), generator=g, device=self.dev)]
        return torch.cat([bc, umi], dim=1).contiguous(), r2
