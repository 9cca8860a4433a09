"""CPU: the command-line front end's gzip decoder (csrc/fast_inflate.hpp) against zlib -- every block type, compression
level and strategy, multi-member files, optional header fields, sync-flushed streams, maximum-distance and
maximum-length matches; damaged files must fail with an error (CRC-32 / ISIZE of every member are verified), never crash."""
import os
import random
import struct
import subprocess
import sys
import zlib

import pytest

import kallisto_b200 as K
from tests import util


def gz_member(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, memlevel=8):
    c = zlib.compressobj(level, zlib.DEFLATED, 31, memlevel, strategy)
    return c.compress(data) + c.flush()


def corpus():
    rnd = random.Random(5)
    fastq = b"".join(b"@r%09d/1\n%s\n+\n%s\n" % (i, bytes(rnd.choice(b"ACGT") for _ in range(100)),
                                                bytes(rnd.choice(b"IIIIIIHG5#") for _ in range(100))) for i in range(2500))
    datas = {"empty": b"", "one": b"A", "fastq": fastq, "zeros": b"\0" * 300000,
             "rand": bytes(rnd.randrange(256) for _ in range(150000)),
             "text": b"the quick brown fox jumps over the lazy dog\n" * 4000, "ramp": bytes(i & 255 for i in range(100000)),
             "far": bytes(rnd.randrange(256) for _ in range(32768)) * 5, "len258": b"ab" * 150000}
    cases = {}
    for name, data in datas.items():
        for lvl in (0, 1, 6, 9):
            cases["%s_l%d" % (name, lvl)] = (gz_member(data, lvl), data)
        cases[name + "_fixed"] = (gz_member(data, 6, zlib.Z_FIXED), data)
        cases[name + "_huff"] = (gz_member(data, 6, zlib.Z_HUFFMAN_ONLY), data)
        cases[name + "_rle"] = (gz_member(data, 6, zlib.Z_RLE), data)
        cases[name + "_mem1"] = (gz_member(data, 9, zlib.Z_DEFAULT_STRATEGY, 1), data)
    cases["multi"] = (gz_member(fastq[:50000], 1) + gz_member(b"", 6) + gz_member(fastq[50000:], 9), fastq)
    hdr = bytes([0x1f, 0x8b, 8, 4 | 8 | 16 | 2, 0, 0, 0, 0, 0, 3]) + struct.pack("<H", 5) + b"EXTRA" + b"name.fq\0" + b"a comment\0"
    hdr += struct.pack("<H", zlib.crc32(hdr) & 0xFFFF)
    raw = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = raw.compress(fastq) + raw.flush()
    cases["hdrfields"] = (hdr + body + struct.pack("<II", zlib.crc32(fastq), len(fastq) & 0xFFFFFFFF), fastq)
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    parts = []
    for i in range(0, len(fastq), 7000):
        parts.append(c.compress(fastq[i:i + 7000]))
        parts.append(c.flush(zlib.Z_SYNC_FLUSH if i % 14000 else zlib.Z_FULL_FLUSH))
    parts.append(c.flush())
    cases["syncflush"] = (b"".join(parts), fastq)
    cases["trailing_garbage"] = (gz_member(fastq, 6) + b"\0\0\0garbage", fastq)
    big = os.urandom(1 << 20) + fastq * 20          # larger than the decoder's 4 MiB output chunk
    cases["big_mixed"] = (gz_member(big, 1), big)
    return cases


def test_decoder_matches_zlib(tmp_path):
    for name, (blob, data) in corpus().items():
        p = tmp_path / (name + ".gz")
        p.write_bytes(blob)
        assert K.gz_summary(str(p)) == (len(data), zlib.crc32(data)), name


def test_fastq_reader_through_gz_equals_zlib_path(tmp_path, monkeypatch):
    import gzip
    rnd = random.Random(9)
    recs = [bytes(rnd.choice(b"ACGTN") for _ in range(rnd.choice([0, 30, 75, 151]))) for _ in range(20000)]
    raw = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)) for i, s in enumerate(recs))
    p = tmp_path / "r.fq.gz"
    p.write_bytes(gzip.compress(raw[:len(raw) // 2], 1) + gzip.compress(raw[len(raw) // 2:], 9))
    fast = K.fastx_summary(str(p))
    monkeypatch.setenv("KB_FASTGZ", "0")              # the reference's way: zlib gzread
    assert K.fastx_summary(str(p)) == fast
    assert fast[0] == len(recs) and fast[1] == sum(map(len, recs))


def test_damaged_gzip_fails_cleanly(tmp_path):
    rnd = random.Random(3)
    cases = corpus()
    blobs = [cases[k][0] for k in ("fastq_l1", "fastq_l9", "fastq_fixed", "rand_l6", "multi", "len258_l9")]
    n = 0
    for i in range(400):
        b = bytearray(rnd.choice(blobs))
        mode = i % 4
        if mode == 0:
            for _ in range(rnd.randrange(1, 6)):
                b[rnd.randrange(len(b))] = rnd.randrange(256)
        elif mode == 1:
            b = b[:rnd.randrange(len(b))]
        elif mode == 2:
            q = rnd.randrange(len(b))
            b[q:q] = bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 9)))
        else:
            b[rnd.randrange(10, len(b))] ^= 1 << rnd.randrange(8)
        (tmp_path / ("c%03d.gz" % i)).write_bytes(bytes(b))
        n += 1
    child = ("import sys, os\nsys.path.insert(0, sys.argv[1])\nimport kallisto_b200 as K\n"
             "for fn in sorted(os.listdir(sys.argv[2])):\n"
             "    try:\n        K.gz_summary(os.path.join(sys.argv[2], fn)); print('OK', flush=True)\n"
             "    except K.KallistoB200Error:\n        print('ERR', flush=True)\n")
    r = subprocess.run([sys.executable, "-c", child, util.ROOT, str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, "the decoder crashed: rc %d\n%s" % (r.returncode, r.stderr[-400:])
    lines = r.stdout.split()
    assert len(lines) == n
    assert lines.count("ERR") > n * 0.9          # a flipped bit that still yields the same CRC-checked stream is vanishingly rare
