#!/usr/bin/env python
"""Closes the one pin this image cannot: abundance.h5 is written by csrc/h5_writer.hpp without libhdf5 and checked by an
own reader (tests/h5mini.py), because neither libhdf5 nor h5py exists here.  Where h5py IS available, run

    python tools/check_h5_with_h5py.py out/abundance.h5

It opens the file with the real library and compares every dataset (dtype, shape, values, chunking, deflate level) with
what tests/h5mini.py reads and with the layout H5Writer produces (src/H5Writer.cpp:4-71, src/h5utils.h:42-91).  Exit code
0 = libhdf5 agrees; 2 = h5py is not installed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import h5mini  # noqa: E402


def main():
    try:
        import h5py
    except ImportError:
        print("h5py is not installed: nothing checked")
        return 2
    path = sys.argv[1]
    mine = h5mini.read(path)
    bad = 0
    with h5py.File(path, "r") as f:
        def walk(g, m, prefix):
            nonlocal bad
            if sorted(g.keys()) != sorted(m.keys()):
                print("members differ under", prefix or "/", sorted(g.keys()), sorted(m.keys()))
                bad += 1
            for k in g:
                if isinstance(g[k], h5py.Group):
                    walk(g[k], m[k], prefix + "/" + k)
                    continue
                d = g[k]
                v = d[()]
                want = m[k]
                if isinstance(want, list):                      # fixed-length strings
                    got = [x.decode() if isinstance(x, bytes) else str(x) for x in v.tolist()]
                    ok = got == want and d.dtype.kind == "S"
                else:
                    ok = v.dtype == want.dtype and np.array_equal(v, want)
                ok = ok and d.chunks == d.shape and d.compression == "gzip" and d.compression_opts == 6
                print(("ok   " if ok else "DIFF ") + prefix + "/" + k, d.dtype, d.shape)
                bad += 0 if ok else 1
        walk(f, mine, "")
    print("libhdf5 agrees with tests/h5mini.py" if not bad else "%d differences" % bad)
    return 0 if not bad else 1


if __name__ == "__main__":
    sys.exit(main())
