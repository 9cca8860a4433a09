"""CPU: the library's FASTA/FASTQ reader (kallisto_b200/csrc/fastx.hpp, kseq_read grammar)."""
import gzip
import os

import pytest

import kallisto_b200 as K
from oracle import oracle as O
from tests import util


def fnv(seqs):
    h = 1469598103934665603
    M = (1 << 64) - 1
    for s in seqs:
        for c in s:
            h = ((h ^ c) * 1099511628211) & M
        h = ((h ^ 0xFF) * 1099511628211) & M
    return h


def check(path, seqs):
    n, b, h = K.fastx_summary(path)
    assert n == len(seqs)
    assert b == sum(len(s) for s in seqs)
    assert h == fnv(seqs)


def test_bundled_gz_fastq():
    d = util.dataset("config1")
    check(os.path.join(d["dir"], "reads_1.fastq.gz"), d["s1"])


def test_formats(tmp_path):
    seqs = [b"ACGTACGTAC", b"", b"NNNNACGT", b"acgtnACGT", b"G" * 300]
    # plain 4-line FASTQ, last record without trailing newline
    p = tmp_path / "a.fq"
    p.write_bytes(b"".join(b"@r%d some comment\n%s\n+\n%s\n" % (i, s, b"I" * len(s)) for i, s in enumerate(seqs))[:-1])
    check(str(p), seqs)
    # CRLF line ends (kseq keeps the lone CR of an empty CRLF sequence line, src/kseq.h ks_getuntil2:
    # "str->l > 1"; so does this reader -- use non-empty sequences here)
    p = tmp_path / "b.fq"
    ne = [s for s in seqs if s]
    p.write_bytes(b"".join(b"@r%d\r\n%s\r\n+\r\n%s\r\n" % (i, s, b"I" * len(s)) for i, s in enumerate(ne)))
    check(str(p), ne)
    # multi-line FASTA, gzip, '@' inside a quality string is not a header
    p = tmp_path / "c.fa.gz"
    with gzip.open(p, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">t%d\n" % i)
            for j in range(0, len(s), 7):
                f.write(s[j:j + 7] + b"\n")
    check(str(p), seqs)
    p = tmp_path / "d.fq"
    p.write_bytes(b"@r0\nACGT\n+\n@@@@\n@r1\nTTTT\n+r1\n@III\n")
    check(str(p), [b"ACGT", b"TTTT"])
    # multi-line FASTQ (sequence and quality wrapped)
    p = tmp_path / "e.fq"
    p.write_bytes(b"@r0\nACGT\nACGT\n+\nIIII\nIIII\n@r1\nGG\n+\nII\n")
    check(str(p), [b"ACGTACGT", b"GG"])


def test_missing_file():
    with pytest.raises(K.KallistoB200Error):
        K.fastx_summary("/nonexistent.fq")
