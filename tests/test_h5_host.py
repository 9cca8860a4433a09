"""CPU: abundance.h5 as written by csrc/h5_writer.hpp (no libhdf5 in this build; H5Writer of the reference:
src/H5Writer.cpp:4-71, src/h5utils.h:42-91) read back by the independent, strict reader tests/h5mini.py: every dataset the
reference writes is there, with the reference's types (int32 / double / fixed-length NUL-terminated strings), one
deflate-compressed chunk each, and holds exactly the numbers of abundance.tsv's run."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import h5mini, util
from tests.test_cli_host_pipeline import build

CSRC = os.path.join(util.ROOT, "kallisto_b200", "csrc")
pytestmark = pytest.mark.skipif(not shutil.which("g++"), reason="no g++")


@pytest.mark.parametrize("n,B", [(1, 0), (1000, 3), (5000, 300)])
def test_writer_round_trip(tmp_path, n, B):
    exe = str(tmp_path / "h5_driver")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + CSRC, "-o", exe, os.path.join(util.ROOT, "tests", "stub", "h5_driver.cpp"), "-lz", "-pthread"])
    path = str(tmp_path / "t.h5")
    subprocess.check_call([exe, path, str(n), str(B)])
    f = h5mini.File(path)
    r = f.root
    assert f.deflate_level == 6
    assert sorted(r) == (["aux", "bootstrap", "est_counts"] if B else ["aux", "est_counts"])
    est = np.array([i * 0.5 + 1.0 / (i + 1) for i in range(n)])
    assert r["est_counts"].dtype == np.dtype("<f8") and np.array_equal(r["est_counts"], est)          # bit-exact doubles
    assert r["aux"]["ids"] == ["ENST%d%s|gene" % (i, "x" * (i % 40)) for i in range(n)]
    assert r["aux"]["lengths"].dtype == np.dtype("<i4") and np.array_equal(r["aux"]["lengths"], 200 + 7 * np.arange(n))
    assert list(r["aux"]["num_bootstrap"]) == [B]
    if B:
        assert sorted(r["bootstrap"]) == sorted("bs%d" % b for b in range(B))
        for b in (0, B // 2, B - 1):
            assert np.array_equal(r["bootstrap"]["bs%d" % b], est * (b + 1))
        assert f.int_k >= (B + 7) // 8 / 2                      # all symbol table nodes of /bootstrap under one B-tree node


def test_quant_writes_abundance_h5_like_an_hdf5_build(tmp_path):
    """Through the command line (host side against the stub library): without --plaintext abundance.h5 appears next to
    abundance.tsv and holds the same run (src/main.cpp:2693-2702); bootstraps go into /bootstrap instead of
    bs_abundance_*.tsv (:2732-2776); with --plaintext there is no HDF5 file."""
    exe = build(str(tmp_path / "stub"))
    ds = os.path.join(util.GOLDEN, "synth_small")
    files = [os.path.join(ds, "reads_%d.fastq.gz" % m) for m in (1, 2)]
    idx = os.path.join(ds, "transcripts.kidx")
    env = dict(os.environ, KB_CLI_CLEANUP="1")
    out = tmp_path / "o"
    r = subprocess.run([exe, "quant", "-i", idx, "-o", str(out), "-b", "11", "-t", "2"] + files, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-500:]
    assert sorted(os.listdir(out)) == ["abundance.h5", "abundance.tsv", "run_info.json"]
    h = h5mini.read(str(out / "abundance.h5"))
    names, lens, eff, est, tpm = util.read_abundance(str(out / "abundance.tsv"))
    assert sorted(h) == ["aux", "bootstrap", "est_counts"]
    assert sorted(h["aux"]) == sorted(["num_bootstrap", "num_processed", "fld", "bias_observed", "bias_normalized", "kallisto_version",
                                       "index_version", "call", "start_time", "ids", "eff_lengths", "lengths"])
    assert h["aux"]["ids"] == names and np.array_equal(h["aux"]["lengths"], lens)
    np.testing.assert_allclose(h["est_counts"], est, rtol=1e-5)         # the text file has 6 significant digits
    np.testing.assert_allclose(h["aux"]["eff_lengths"], eff, rtol=1e-5)
    assert list(h["aux"]["num_bootstrap"]) == [11] and list(h["aux"]["num_processed"]) == [20000] and list(h["aux"]["index_version"]) == [13]
    assert h["aux"]["kallisto_version"] == ["0.51.1"] and h["aux"]["call"][0].startswith(exe + " quant -i ")
    assert len(h["aux"]["fld"]) == 1000 and h["aux"]["fld"][200] == 10          # the stub's histogram
    assert len(h["aux"]["bias_observed"]) == 4096 and set(h["aux"]["bias_observed"]) == {1} and set(h["aux"]["bias_normalized"]) == {1.0}
    assert sorted(h["bootstrap"]) == sorted("bs%d" % b for b in range(11))
    import json
    assert json.load(open(out / "run_info.json"))["n_bootstraps"] == 11
    # -l / -s: the stored distribution is the truncated Gaussian of trunc_gaussian_counts (src/weights.cpp:273-296)
    out2 = tmp_path / "o2"
    r = subprocess.run([exe, "quant", "-i", idx, "-o", str(out2), "--single", "-l", "200", "-s", "20", files[0]], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-500:]
    fld = h5mini.read(str(out2 / "abundance.h5"))["aux"]["fld"]
    x = (np.arange(1000) - 200.0) / 20.0
    dens = np.exp(-0.5 * x * x) / 20.0
    assert np.array_equal(fld, np.round(dens * 10000 / dens.sum()).astype(np.int32)) and 9990 <= fld.sum() <= 10010
    out3 = tmp_path / "o3"
    r = subprocess.run([exe, "quant", "-i", idx, "-o", str(out3), "--plaintext", "-b", "2"] + files, capture_output=True, text=True, env=env)
    assert r.returncode == 0 and sorted(os.listdir(out3)) == ["abundance.tsv", "bs_abundance_0.tsv", "bs_abundance_1.tsv", "run_info.json"]


def test_h5dump_round_trip(tmp_path):
    """`kallisto_b200 h5dump` (H5Converter, src/H5Writer.cpp:75-200) reads the file back with its own reader
    (csrc/h5_reader.hpp) and writes the plaintext files: after quant -> abundance.h5 -> h5dump, abundance.tsv is the file
    quant wrote itself, byte for byte (the HDF5 file holds the full doubles; the formatting code is the same), and there is
    one bs_abundance file per bootstrap; run_info.json carries what the reference's converter can know."""
    import json
    exe = build(str(tmp_path / "stub"))
    ds = os.path.join(util.GOLDEN, "synth_small")
    files = [os.path.join(ds, "reads_%d.fastq.gz" % m) for m in (1, 2)]
    env = dict(os.environ, KB_CLI_CLEANUP="1")
    out = tmp_path / "o"
    r = subprocess.run([exe, "quant", "-i", os.path.join(ds, "transcripts.kidx"), "-o", str(out), "-b", "5"] + files, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-500:]
    dump = tmp_path / "dump"
    r = subprocess.run([exe, "h5dump", "-o", str(dump), str(out / "abundance.h5")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    assert "[h5dump] number of targets: 3" in r.stderr and "[h5dump] number of bootstraps: 5" in r.stderr
    assert sorted(os.listdir(dump)) == ["abundance.tsv"] + ["bs_abundance_%d.tsv" % b for b in range(5)] + ["run_info.json"]
    assert open(dump / "abundance.tsv", "rb").read() == open(out / "abundance.tsv", "rb").read()
    txt = open(dump / "run_info.json").read()
    assert '"n_unique": -1,' in txt and '"k-mer length": dummy k-mer length,' in txt and '"n_bootstraps": 5,' in txt and '"n_processed": 20000,' in txt
    assert json.load(open(out / "run_info.json"))["call"] in txt
    # option checking of the reference (CheckOptionsH5Dump, src/main.cpp:2027-2072)
    for args, msg in [([str(out / "abundance.h5")], "Error: You must specify an output directory."),
                      (["-o", str(tmp_path / "d2")], "Error: Missing H5 files"),
                      (["-o", str(tmp_path / "d3"), "nope.h5"], "Error: H5 file not found nope.h5"),
                      (["-o", str(tmp_path / "d4"), str(out / "abundance.tsv")], "is not an HDF5 file")]:
        r = subprocess.run([exe, "h5dump"] + args, capture_output=True, text=True)
        assert r.returncode == 1 and msg in r.stderr, (args, r.stderr[-300:])


def test_h5dump_large_file(tmp_path):
    """5000 targets, 300 bootstraps (a /bootstrap group of 38 symbol table nodes under a B-tree node of raised order)."""
    drv = str(tmp_path / "h5_driver")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + CSRC, "-o", drv, os.path.join(util.ROOT, "tests", "stub", "h5_driver.cpp"), "-lz", "-pthread"])
    path = str(tmp_path / "t.h5")
    subprocess.check_call([drv, path, "5000", "300"])
    exe = build(str(tmp_path / "stub"))
    dump = tmp_path / "dump"
    r = subprocess.run([exe, "h5dump", "-o", str(dump), path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    assert len(os.listdir(dump)) == 302
    est = np.array([i * 0.5 + 1.0 / (i + 1) for i in range(5000)])
    names, lens, eff, got, tpm = util.read_abundance(str(dump / "bs_abundance_299.tsv"))
    assert names[4999] == "ENST4999%s|gene" % ("x" * (4999 % 40)) and np.array_equal(lens, 200 + 7 * np.arange(5000))
    np.testing.assert_allclose(got, est * 300, rtol=1e-5)
    np.testing.assert_allclose(eff, lens - 150.25, rtol=1e-5)
    x = est * 300 / (lens - 150.25)
    np.testing.assert_allclose(tpm, x / x.sum() * 1e6, rtol=1e-4)


def test_h5dump_never_crashes_on_damaged_files(tmp_path):
    """Truncated files and files with flipped bytes end with an error message (exit 1) or, when the damage hits only
    padding, with a normal conversion -- never with a signal or a hang (csrc/h5_reader.hpp bounds every address)."""
    import random
    drv = str(tmp_path / "h5_driver")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + CSRC, "-o", drv, os.path.join(util.ROOT, "tests", "stub", "h5_driver.cpp"), "-lz", "-pthread"])
    path = str(tmp_path / "t.h5")
    subprocess.check_call([drv, path, "200", "20"])
    good = open(path, "rb").read()
    exe = build(str(tmp_path / "stub"))
    rnd = random.Random(5)
    cases = [good[:n] for n in (0, 7, 95, 96, 500, len(good) // 2, len(good) - 1)]
    for _ in range(150):
        b = bytearray(good)
        for _ in range(rnd.choice((1, 1, 2, 8, 64))):
            b[rnd.randrange(len(b))] = rnd.randrange(256)
        cases.append(bytes(b))
    # a group B-tree node whose children all point back at the node itself: bounded, not exponential
    import re
    import struct
    b = bytearray(good)
    for m in re.finditer(b"TREE\x00\x00", good):
        if struct.unpack_from("<H", b, m.start() + 6)[0] >= 2:
            for i in range(30):
                struct.pack_into("<Q", b, m.start() + 32 + 16 * i, m.start())
            struct.pack_into("<H", b, m.start() + 6, 30)
            break
    cases.append(bytes(b))
    n_err = 0
    for i, data in enumerate(cases):
        p = tmp_path / "bad.h5"
        p.write_bytes(data)
        r = subprocess.run([exe, "h5dump", "-o", str(tmp_path / "d"), str(p)], capture_output=True, text=True, timeout=60)
        assert r.returncode in (0, 1), (i, r.returncode, r.stderr[-300:])
        n_err += r.returncode
    assert n_err >= 10
