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


# ---- parallel ingest path (ParallelFastx): always the sequential parse, whatever the segment cuts hit ----
def _same(path, threads=(2, 3, 8)):
    want = K.fastx_summary(path, 1)
    for t in threads:
        assert K.fastx_summary(path, t) == want, (path, t)
    return want


def test_parallel_matches_sequential(tmp_path, monkeypatch):
    import random
    rnd = random.Random(7)
    recs = []
    for i in range(3000):
        L = rnd.choice([0, 1, 5, 30, 50, 75, 100, 151])
        s = bytes(rnd.choice(b"ACGTNacgt") for _ in range(L))
        q = bytes(rnd.choice(b"@+>I#5F!") for _ in range(L))     # quality lines that look like headers / separators
        recs.append((s, q))
    p = tmp_path / "p.fq"
    p.write_bytes(b"".join(b"@r%d/1\n%s\n+\n%s\n" % (i, s, q) for i, (s, q) in enumerate(recs)))
    seqs = [s for s, _ in recs]
    for window in ("1000", "4096", "65536", "100000000"):        # bytes per thread and window: many cuts ... one
        monkeypatch.setenv("KB_FASTX_WINDOW", window)
        n, b, h = _same(str(p))
        assert n == len(seqs) and b == sum(map(len, seqs)) and h == fnv(seqs)


def test_parallel_odd_layouts(tmp_path, monkeypatch):
    monkeypatch.setenv("KB_FASTX_WINDOW", "64")
    # multi-line FASTQ: guessed segment starts are frequently wrong and must be rejected by the proof
    body = b"".join(b"@r%d\nACGTACGTAC\nGGGGG\n+\n@IIIIIIIII\n+IIII\n" % i for i in range(400))
    p = tmp_path / "ml.fq"
    p.write_bytes(body)
    n, b, h = _same(str(p))
    assert n == 400 and b == 400 * 15
    # FASTA (no '+' lines at all: no segment start can be guessed, the window is parsed by one thread)
    p = tmp_path / "x.fa"
    p.write_bytes(b"".join(b">t%d\nACGTAC\nGT\n" % i for i in range(500)))
    n, b, h = _same(str(p))
    assert n == 500 and b == 500 * 8
    # CRLF, no trailing newline, empty file, garbage before the first header
    p = tmp_path / "crlf.fq"
    p.write_bytes(b"".join(b"@r%d\r\nACGTA\r\n+\r\nIIIII\r\n" % i for i in range(300))[:-2])
    assert _same(str(p))[0] == 300
    p = tmp_path / "empty.fq"
    p.write_bytes(b"")
    assert _same(str(p)) == (0, 0, fnv([]))
    p = tmp_path / "junk.fq"
    p.write_bytes(b"junk line\n\n" + b"".join(b"@r%d\nAC\n+\nII\n" % i for i in range(100)))
    assert _same(str(p))[0] == 100


def test_parallel_gz_falls_back_to_zlib(tmp_path):
    d = util.dataset("config1")
    f = os.path.join(d["dir"], "reads_1.fastq.gz")
    assert K.fastx_summary(f, 8) == K.fastx_summary(f, 1)


def test_garbage_input_never_crashes(tmp_path):
    """Random bytes, records cut mid-way, a truncated gzip stream: both readers finish (with records or an error)."""
    import random
    import subprocess
    import sys
    rnd = random.Random(11)
    good = b"".join(b"@r%d\nACGTNACGT%s\n+\nIIIIIIIII%s\n" % (i, b"A" * (i % 7), b"I" * (i % 7)) for i in range(300))
    cases = {"rand": bytes(rnd.randrange(256) for _ in range(20000)),
             "ats": b"@" * 5000, "plus": b"@x\n" + b"+\n" * 3000, "nl": b"\n" * 4000,
             "cut1": good[:len(good) // 2 + 3], "cut2": good[:len(good) - 5], "nul": good.replace(b"A", b"\0", 50)}
    for i in range(8):
        b = bytearray(good)
        for _ in range(20):
            b[rnd.randrange(len(b))] = rnd.randrange(256)
        cases["flip%d" % i] = bytes(b)
    gz = gzip.compress(good)
    cases["trunc.gz"] = gz[:len(gz) // 2]
    for name, blob in cases.items():
        (tmp_path / name).write_bytes(blob)
    child = ("import sys, os\nsys.path.insert(0, sys.argv[1])\nimport kallisto_b200 as K\n"
             "os.environ['KB_FASTX_WINDOW'] = '257'\n"
             "for fn in sorted(os.listdir(sys.argv[2])):\n"
             "    for t in (1, 4):\n"
             "        try:\n"
             "            r = K.fastx_summary(os.path.join(sys.argv[2], fn), t)\n"
             "            print(fn, t, 'OK', r[0], r[1], r[2], flush=True)\n"
             "        except K.KallistoB200Error:\n"
             "            print(fn, t, 'ERR', flush=True)\n")
    r = subprocess.run([sys.executable, "-c", child, util.ROOT, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, "a reader crashed: rc %d\n%s\n%s" % (r.returncode, r.stdout[-400:], r.stderr[-400:])
    out = {}
    for line in r.stdout.splitlines():
        a = line.split()
        out.setdefault(a[0], {})[a[1]] = a[2:]
    assert set(out) == set(cases)
    for name, res in out.items():
        assert res["1"] == res["4"], (name, res)          # the parallel reader is the sequential parse, also on garbage


def test_crlf_multiline_quality_counts_like_kseq(tmp_path, monkeypatch):
    """ks_getuntil2 (src/kseq.h:137) drops the trailing CR of every line, quality lines included; counting it would
    end a multi-line quality block early and turn a leftover quality line that starts with '@' into a phantom record.
    Records: 20 bases on 10 CRLF lines, quality on 10 CRLF lines of which the 8th..10th start with '@'."""
    import re
    import subprocess
    recs = []
    want = []
    for i in range(50):
        seq = ("ACGTTGCAAC" * 2)[i % 7:][:20].ljust(20, "A")
        want.append(seq.encode())
        q = ["II"] * 7 + ["@I", "@+", "@>"]
        recs.append("@r%d\r\n" % i + "".join(seq[j:j + 2] + "\r\n" for j in range(0, 20, 2)) + "+\r\n" + "".join(x + "\r\n" for x in q))
    p = tmp_path / "crlf.fq"
    p.write_bytes("".join(recs).encode())
    n, nb, h1 = K.fastx_summary(str(p))
    assert (n, nb) == (50, 1000)
    monkeypatch.setenv("KB_FASTX_WINDOW", "300")
    assert K.fastx_summary(str(p), threads=4) == (n, nb, h1)
    if O.have_ref():   # the reference's own reader agrees on the record count
        idx = os.path.join(util.GOLDEN, "config1", "transcripts.kidx")
        r = subprocess.run([O.REF_BIN, "quant", "-i", idx, "-o", str(tmp_path / "o"), "--single", "-l", "200", "-s", "20", str(p)],
                           capture_output=True, text=True)
        m = re.search(r"processed ([0-9,]+) reads", r.stderr)
        assert m and int(m.group(1).replace(",", "")) == 50


def test_wrong_length_quality_ends_the_file_like_kseq(tmp_path, monkeypatch):
    """kseq_read returns -2 for a record whose quality string has another length than its sequence, and
    FastqSequenceReader::fetchSequences ends the file there (src/ProcessReads.cpp:3178-3182): the record and everything
    after it are not processed."""
    import re
    import subprocess
    good = "".join("@r%d\nACGTACGTACGTACGTACGTACGTACGTACGTACGT\n+\n%s\n" % (i, "I" * 36) for i in range(300))
    bad = "@bad\nACGTACGTAC\n+\nIIII\n"
    p = tmp_path / "trunc.fq"
    p.write_bytes((good + bad + good).encode())
    n, nb, h1 = K.fastx_summary(str(p))
    assert n == 300 and nb == 300 * 36
    for w in ("200", "3000", "100000"):
        monkeypatch.setenv("KB_FASTX_WINDOW", w)
        assert K.fastx_summary(str(p), threads=4) == (n, nb, h1)
    if O.have_ref():
        idx = os.path.join(util.GOLDEN, "config1", "transcripts.kidx")
        r = subprocess.run([O.REF_BIN, "quant", "-i", idx, "-o", str(tmp_path / "o"), "--single", "-l", "200", "-s", "20", str(p)],
                           capture_output=True, text=True)
        m = re.search(r"processed ([0-9,]+) reads", r.stderr)
        assert m and int(m.group(1).replace(",", "")) == 300
