"""CPU: the v13 index parser (csrc/index_v13.cpp, host only) on damaged files -- truncated anywhere, random byte
flips, absurd length fields.  It has to answer with an error (or parse, when the damage hit a section the loader
skips or plain sequence bytes), never crash or hang: KmerIndex::load has the same duty towards its users
(src/KmerIndex.cpp:1330-1559 exits with a message on a bad header)."""
import os
import subprocess
import sys

import numpy as np

from tests import util

CHILD = r"""
import sys, os
sys.path.insert(0, sys.argv[1])
import kallisto_b200 as K
for fn in sorted(os.listdir(sys.argv[2])):
    try:
        K.inspect_index(os.path.join(sys.argv[2], fn))
        print(fn, "OK", flush=True)
    except K.KallistoB200Error as e:
        print(fn, "ERR", flush=True)
"""


def test_damaged_index_files(tmp_path):
    data = open(os.path.join(util.GOLDEN, "config1", "transcripts.kidx"), "rb").read()
    rng = np.random.default_rng(5)
    cases = {}
    for cut in [0, 7, 8, 15, 16, 100] + [int(x) for x in rng.integers(0, len(data), 20)]:
        cases["trunc%07d" % cut] = data[:cut]
    for i in range(30):
        b = bytearray(data)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        cases["flip%02d" % i] = bytes(b)
    for i in range(10):
        b = bytearray(data)
        pos = int(rng.integers(0, len(b) - 8))
        b[pos:pos + 8] = (0xFFFFFFFFFFFFFFF0).to_bytes(8, "little")
        cases["huge%02d" % i] = bytes(b)
    for name, blob in cases.items():
        (tmp_path / name).write_bytes(blob)
    r = subprocess.run([sys.executable, "-c", CHILD, util.ROOT, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, "the parser crashed: rc %d\n%s\n%s" % (r.returncode, r.stdout[-500:], r.stderr[-500:])
    seen = dict(line.split()[:2] for line in r.stdout.splitlines() if line.strip())
    assert set(seen) == set(cases)
    assert all(v == "ERR" for k, v in seen.items() if k.startswith("trunc"))
