"""CPU: the host side of the command line (parser threads, parallel reader for plain files, gzip decoder + prefetch
thread, lock-step hand-over of the per-file batches, writers) run end to end against tests/stub/stub_abi.cpp, a
stand-in for the library that only digests the fragments it receives, in order.  Whatever the thread count, window
size, batch cut or input compression, the device must be handed the same reads in the same order; one configuration
also runs under ThreadSanitizer.  (The real kernels behind the same CLI are covered by tests/test_gpu_cli.py.)"""
import gzip
import os
import shutil
import subprocess

import pytest

from tests import util

CSRC = os.path.join(util.ROOT, "kallisto_b200", "csrc")
INC = os.path.join(util.ROOT, "include")
pytestmark = pytest.mark.skipif(not shutil.which("g++"), reason="no g++")


def build(dst, extra=()):
    os.makedirs(dst, exist_ok=True)
    lib = os.path.join(dst, "libkallisto_b200.so")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-I" + INC, *extra, "-o", lib,
                           os.path.join(util.ROOT, "tests", "stub", "stub_abi.cpp")])
    exe = os.path.join(dst, "cli")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-I" + INC, "-I" + CSRC, *extra, "-o", exe,
                           os.path.join(CSRC, "cli_main.cpp"), "-L" + dst, "-lkallisto_b200", "-Wl,-rpath," + dst, "-lz", "-lpthread"])
    return exe


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    d = tmp_path_factory.mktemp("stubcli")
    exe = build(str(d / "plain"))
    ds = os.path.join(util.GOLDEN, "synth_small")
    plain = []
    for m in (1, 2):
        p = d / ("r%d.fq" % m)
        p.write_bytes(gzip.open(os.path.join(ds, "reads_%d.fastq.gz" % m)).read())
        plain.append(str(p))
    gz = [os.path.join(ds, "reads_%d.fastq.gz" % m) for m in (1, 2)]
    return dict(dir=d, exe=exe, plain=plain, gz=gz, idx=os.path.join(ds, "transcripts.kidx"))


def quant(exe, idx, out, args, env=None):
    e = dict(os.environ, KB_CLI_CLEANUP="1")      # orderly release at the end (what the sanitizer runs look at)
    e.update(env or {})
    r = subprocess.run([exe, "quant", "-i", idx, "-o", str(out), "--plaintext"] + args, capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    return open(os.path.join(str(out), "abundance.tsv")).read(), r.stderr


def test_same_reads_in_the_same_order_whatever_the_ingest_configuration(setup):
    s = setup
    base, _ = quant(s["exe"], s["idx"], s["dir"] / "o0", ["-t", "1"] + s["plain"])
    assert "digest_lo" in base
    configs = [
        (["-t", "8"] + s["plain"], {"KB_FASTX_WINDOW": "30000", "KB_CLI_BATCH_READS": "700,1100"}),
        (["-t", "4"] + s["plain"], {"KB_FASTX_WINDOW": "5000", "KB_CLI_BATCH_READS": "333,1000"}),
        (["-t", "2"] + s["plain"], {"KB_CLI_BATCH_READS": "1,7"}),
        (["-t", "1"] + s["gz"], {}),
        (["-t", "8"] + s["gz"], {"KB_CLI_BATCH_READS": "512,4096"}),
        (["-t", "1"] + s["gz"], {"KB_FASTGZ": "0"}),
        (["-t", "3", s["plain"][0], s["gz"][1]], {"KB_FASTX_WINDOW": "9999"}),       # one plain, one compressed
    ]
    for i, (args, env) in enumerate(configs):
        got, _ = quant(s["exe"], s["idx"], s["dir"] / ("o%d" % (i + 1)), args, env)
        assert got == base, (args, env)
    # single-end: only the first file, a different digest, but again independent of the configuration
    se = ["--single", "-l", "200", "-s", "20"]
    b1, _ = quant(s["exe"], s["idx"], s["dir"] / "s0", se + ["-t", "1", s["plain"][0]])
    b2, _ = quant(s["exe"], s["idx"], s["dir"] / "s1", se + ["-t", "8", s["plain"][0]], {"KB_FASTX_WINDOW": "20000"})
    b3, _ = quant(s["exe"], s["idx"], s["dir"] / "s2", se + ["-t", "8", s["gz"][0]])
    assert b1 == b2 == b3 and b1 != base


def test_file_sets_stay_in_step_when_a_pair_is_ragged(setup):
    """FastqSequenceReader::fetchSequences (src/ProcessReads.cpp:3178-3262) stops a file set at its SHORTEST file
    and opens the next set in step.  A first pair whose second file holds only 100 reads must therefore give the
    reads of (first 100 pairs) + (the whole second pair) -- not mates shifted by the difference."""
    s = setup
    lines = [open(f, "rb").readlines() for f in s["plain"]]
    short2 = s["dir"] / "short_2.fq"
    short2.write_bytes(b"".join(lines[1][:4 * 100]))
    short1 = s["dir"] / "short_1.fq"
    short1.write_bytes(b"".join(lines[0][:4 * 100]))
    want, _ = quant(s["exe"], s["idx"], s["dir"] / "rag0", ["-t", "1", str(short1), str(short2)] + s["plain"])
    for i, (t, env) in enumerate([("1", {}), ("4", {"KB_FASTX_WINDOW": "30000"}), ("8", {"KB_CLI_BATCH_READS": "64,200"})]):
        got, _ = quant(s["exe"], s["idx"], s["dir"] / ("rag%d" % (i + 1)), ["-t", t, s["plain"][0], str(short2)] + s["plain"], env)
        assert got == want, (t, env)
    # ... and the other way round (first file the short one)
    got, _ = quant(s["exe"], s["idx"], s["dir"] / "rag9", ["-t", "4", str(short1), s["plain"][1]] + s["plain"])
    assert got == want


def test_bus_records_in_read_order(setup):
    s = setup
    d = os.path.join(util.GOLDEN, "bus10x")
    files = [os.path.join(d, "sc_reads_1.fastq.gz"), os.path.join(d, "sc_reads_2.fastq.gz")]
    idx = os.path.join(util.GOLDEN, "config1", "transcripts.kidx")
    outs = []
    for i, env in enumerate([{}, {"KB_CLI_BATCH_READS": "300,470"}, {"KB_FASTGZ": "0", "KB_CLI_BATCH_READS": "1000,64"}]):
        e = dict(os.environ)
        e.update(env)
        out = s["dir"] / ("b%d" % i)
        r = subprocess.run([s["exe"], "bus", "-i", idx, "-o", str(out), "-x", "10xv2", "-t", "4"] + files, capture_output=True, text=True, env=e, timeout=600)
        assert r.returncode == 0, r.stderr[-500:]
        outs.append(open(out / "output.bus", "rb").read())
    assert outs[0] == outs[1] == outs[2] and len(outs[0]) > 1000


def test_bus_bulk_samples_follow_the_file_sets(setup, tmp_path):
    """`bus -x bulk [--paired]` (src/main.cpp:1050-1107): every file (pair) is a sample.  The stub stamps the sample it was
    told about (kb_bus_begin_sample) into the records, so the record counts per sample must equal the read counts of the
    files whatever the batch cuts; matrix.cells, matrix.sample.barcodes and index.saved are host work and must equal the
    reference's files (tests/golden/buspaired)."""
    import numpy as np
    from oracle import oracle as O
    s = setup
    inputs = util.buspaired_inputs(str(tmp_path))
    ref = os.path.join(util.GOLDEN, "buspaired", "ref_bulk_paired")
    for i, env in enumerate([{}, {"KB_CLI_BATCH_READS": "700,1100"}, {"KB_CLI_BATCH_READS": "12000,5"}]):
        e = dict(os.environ)
        e.update(env)
        out = tmp_path / ("bulk%d" % i)
        r = subprocess.run([s["exe"], "bus", "-i", s["idx"], "-o", str(out), "-x", "bulk", "--paired", "-t", "4"] +
                           [inputs[k] for k in ("a_1", "a_2", "b_1", "b_2")], capture_output=True, text=True, env=e, timeout=600)
        assert r.returncode == 0, r.stderr[-500:]
        hdr, rec = O.read_bus(str(out / "output.bus"))
        assert (hdr["bclen"], hdr["umilen"]) == (16, 1)
        assert np.array_equal(rec["umi"], np.repeat([0, 1], [12000, 8000]))
        for fn in ("matrix.cells", "matrix.sample.barcodes", "index.saved"):
            assert open(out / fn, "rb").read() == open(os.path.join(ref, fn), "rb").read(), fn
        fl = [line.split() for line in open(out / "flens.txt")]
        assert len(fl) == 2 and all(len(x) == 1000 for x in fl) and fl[0][1] == "0" and fl[1][1] == "1"
    # single-end: three files, three samples; no flens.txt
    out = tmp_path / "bulk_se"
    r = subprocess.run([s["exe"], "bus", "-i", s["idx"], "-o", str(out), "-x", "bulk", "-t", "2"] +
                       [inputs[k] for k in ("a_1", "b_1", "a_2")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    _, rec = O.read_bus(str(out / "output.bus"))
    assert np.array_equal(rec["umi"], np.repeat([0, 1, 2], [12000, 8000, 12000]))
    assert not os.path.exists(out / "flens.txt") and os.path.exists(out / "index.saved")
    assert open(out / "matrix.cells").read() == "batch0\nbatch1\nbatch2\n"
    # an odd number of files with --paired, and --paired with a single-read technology, are refused like the reference does
    r = subprocess.run([s["exe"], "bus", "-i", s["idx"], "-o", str(tmp_path / "bad"), "-x", "bulk", "--paired", inputs["a_1"]],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Error: paired-end mode requires an even number of input files" in r.stderr
    r = subprocess.run([s["exe"], "bus", "-i", s["idx"], "-o", str(tmp_path / "bad2"), "-x", "10xv2", "--paired", inputs["a_1"], inputs["a_2"]],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Error: Paired reads are not compatible with the specified technology" in r.stderr


def test_bus_interleaved_input(setup, tmp_path):
    """`bus --inleaved` (src/main.cpp:583,739-741,1000-1012): one file with the reads of every set one after the other gives
    the device exactly the read sets of the separate files, whatever the batch cuts; an incomplete last set is dropped."""
    s = setup
    d = os.path.join(util.GOLDEN, "bus10x")
    files = [os.path.join(d, "sc_reads_1.fastq.gz"), os.path.join(d, "sc_reads_2.fastq.gz")]
    idx = os.path.join(util.GOLDEN, "config1", "transcripts.kidx")
    a, b = (gzip.open(f, "rb").read().split(b"\n") for f in files)
    n = len(a) // 4
    il = tmp_path / "il.fq"
    with open(il, "wb") as f:
        for i in range(n):
            f.write(b"\n".join(a[4 * i:4 * i + 4]) + b"\n" + b"\n".join(b[4 * i:4 * i + 4]) + b"\n")
        f.write(b"\n".join(a[:4]) + b"\n")                     # a first mate without its second: dropped
    def run(args, env=None):
        e = dict(os.environ, KB_CLI_CLEANUP="1")
        e.update(env or {})
        out = tmp_path / ("o%d" % len(os.listdir(tmp_path)))
        r = subprocess.run([s["exe"], "bus", "-i", idx, "-o", str(out), "-x", "10xv2"] + args, capture_output=True, text=True, env=e, timeout=600)
        assert r.returncode == 0, r.stderr[-500:]
        return open(out / "output.bus", "rb").read()
    want = run(["-t", "2"] + files)
    assert run(["-t", "4", "--inleaved", str(il)]) == want
    assert run(["-t", "1", "--inleaved", str(il)], {"KB_CLI_BATCH_READS": "333"}) == want
    assert run(["-t", "8", "--inleaved", str(il)], {"KB_CLI_BATCH_READS": "1", "KB_FASTX_WINDOW": "5000"}) == want
    r = subprocess.run([s["exe"], "bus", "-i", idx, "-o", str(tmp_path / "bad"), "-x", "10xv2", "--inleaved"] + files, capture_output=True, text=True)
    assert r.returncode == 1 and "Error: interleaved input cannot consist of more than one input" in r.stderr


def test_quant_write_index_host(setup, tmp_path):
    """`quant --write-index` (src/ProcessReads.cpp:242-249, src/main.cpp:2658-2661): counts.txt and index.saved next to the
    usual files; index.saved equals the reference's (tests/golden/config1/ref_quant_paired), counts.txt has one
    "id <tab> count" line per equivalence class (the stub reports one class)."""
    d = os.path.join(util.GOLDEN, "config1")
    out = tmp_path / "o"
    r = subprocess.run([setup["exe"], "quant", "-i", os.path.join(d, "transcripts.kidx"), "-o", str(out), "--plaintext", "--write-index",
                        os.path.join(d, "reads_1.fastq.gz"), os.path.join(d, "reads_2.fastq.gz")], capture_output=True, text=True,
                       env=dict(os.environ, KB_CLI_CLEANUP="1"), timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    assert sorted(os.listdir(out)) == ["abundance.tsv", "counts.txt", "index.saved", "run_info.json"]
    assert open(out / "index.saved", "rb").read() == open(os.path.join(d, "ref_quant_paired", "index.saved"), "rb").read()
    assert open(out / "counts.txt").read() == "0\t1\n"


def test_bus_batch_file_host(setup, tmp_path):
    """`bus --batch FILE` on the host side: sample names and barcodes (lines with the same id share one) as the reference
    writes them (tests/golden/buspaired/ref_batchfile), the samples switch with the lines, and the reference's messages
    for a missing batch file / read files next to --batch."""
    import numpy as np
    from oracle import oracle as O
    s = setup
    inputs = util.buspaired_inputs(str(tmp_path))
    bf = util.write_batch_file(str(tmp_path / "batch.txt"), inputs)
    ref = os.path.join(util.GOLDEN, "buspaired", "ref_batchfile")
    out = tmp_path / "o"
    r = subprocess.run([s["exe"], "bus", "-i", s["idx"], "-o", str(out), "--batch", bf, "-t", "4"], capture_output=True, text=True,
                       env=dict(os.environ, KB_CLI_BATCH_READS="900,2000"), timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    for fn in ("matrix.cells", "matrix.sample.barcodes", "index.saved"):
        assert open(out / fn, "rb").read() == open(os.path.join(ref, fn), "rb").read(), fn
    _, rec = O.read_bus(str(out / "output.bus"))
    assert np.array_equal(rec["umi"], np.repeat([0, 1, 0], [12000, 8000, 8000]))       # the stub stamps the sample barcode
    assert len(open(out / "flens.txt").readlines()) == 3
    r = subprocess.run([s["exe"], "bus", "-i", s["idx"], "-o", str(tmp_path / "b1"), "--batch", str(tmp_path / "nope.txt")], capture_output=True, text=True)
    assert r.returncode == 1 and "Error: file not found " + str(tmp_path / "nope.txt") in r.stderr
    r = subprocess.run([s["exe"], "bus", "-i", s["idx"], "-o", str(tmp_path / "b2"), "--batch", bf, inputs["a_1"]], capture_output=True, text=True)
    assert r.returncode == 1 and "Error: cannot specify batch mode and supply read files" in r.stderr


def test_index_saved_of_a_dlist_index_host(setup, tmp_path):
    dl = os.path.join(util.GOLDEN, "dlist")
    out = tmp_path / "o"
    r = subprocess.run([setup["exe"], "bus", "-i", os.path.join(dl, "transcripts.kidx"), "-o", str(out), "-x", "bulk", "--paired",
                        os.path.join(dl, "reads_1.fastq.gz"), os.path.join(dl, "reads_2.fastq.gz")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    assert open(out / "index.saved", "rb").read() == open(os.path.join(util.GOLDEN, "buspaired", "dlist_index.saved"), "rb").read()


def test_host_pipeline_under_thread_sanitizer(setup, tmp_path):
    s = setup
    probe = subprocess.run(["g++", "-fsanitize=thread", "-x", "c++", "-", "-o", str(tmp_path / "probe")], input="int main(){}", text=True,
                           capture_output=True)
    if probe.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available")
    exe = build(str(tmp_path / "tsan"), extra=("-fsanitize=thread",))
    for args, env in [(["-t", "8"] + s["plain"], {"KB_FASTX_WINDOW": "30000", "KB_CLI_BATCH_READS": "700,1100"}),
                      (["-t", "8"] + s["gz"], {"KB_CLI_BATCH_READS": "512,4096"})]:
        _, err = quant(exe, s["idx"], tmp_path / "o", args, env)
        assert "ThreadSanitizer" not in err, err[-2000:]
