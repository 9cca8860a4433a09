"""GPU: the kallisto_b200 command line against the files the unmodified reference wrote for the
same commands (tests/golden/*/ref_quant_*)."""
import json
import os
import subprocess

import pytest

from tests import util

pytestmark = pytest.mark.gpu

BIN = os.path.join(util.ROOT, "kallisto_b200", "kallisto_b200")


def run(args, cwd=None):
    return subprocess.run([BIN] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def same_run_info(a, b):
    ja, jb = json.load(open(a)), json.load(open(b))
    for k in ("n_targets", "n_bootstraps", "n_processed", "n_pseudoaligned", "n_unique", "p_pseudoaligned", "p_unique",
              "kallisto_version", "index_version", "k-mer length"):
        assert ja[k] == jb[k], k
    assert list(ja.keys()) == list(jb.keys())


@pytest.mark.parametrize("name", ["config1", "synth_small", "manyecs"])
def test_quant_paired_with_bootstrap(name, tmp_path):
    ds = util.dataset(name)
    out = tmp_path / "o"
    r = run(["quant", "-i", ds["index"], "-o", str(out), "--plaintext", "-b", "3", "--seed", "42",
             os.path.join(ds["dir"], "reads_1.fastq.gz"), os.path.join(ds["dir"], "reads_2.fastq.gz")])
    assert r.returncode == 0, r.stderr
    ref = os.path.join(ds["dir"], "ref_quant_paired")
    for fn in ["abundance.tsv", "bs_abundance_0.tsv", "bs_abundance_1.tsv", "bs_abundance_2.tsv"]:
        assert open(out / fn).read() == open(os.path.join(ref, fn)).read(), fn
    same_run_info(out / "run_info.json", os.path.join(ref, "run_info.json"))
    assert "reads pseudoaligned" in r.stderr and "Expectation-Maximization algorithm ran for" in r.stderr


def test_quant_fr_stranded(tmp_path):
    ds = util.dataset("synth_small")
    out = tmp_path / "o"
    r = run(["quant", "-i", ds["index"], "-o", str(out), "--plaintext", "--fr-stranded",
             os.path.join(ds["dir"], "reads_1.fastq.gz"), os.path.join(ds["dir"], "reads_2.fastq.gz")])
    assert r.returncode == 0, r.stderr
    ref = os.path.join(ds["dir"], "ref_quant_paired_fr")
    assert open(out / "abundance.tsv").read() == open(os.path.join(ref, "abundance.tsv")).read()
    same_run_info(out / "run_info.json", os.path.join(ref, "run_info.json"))


def test_quant_single_overhang(tmp_path):
    ds = util.dataset("synth_small")
    out = tmp_path / "o"
    r = run(["quant", "-i", ds["index"], "-o", str(out), "--plaintext", "--single", "--single-overhang", "-l", "200", "-s",
             "20", os.path.join(ds["dir"], "reads_1.fastq.gz")])
    assert r.returncode == 0, r.stderr
    ref = os.path.join(ds["dir"], "ref_quant_single_overhang")
    assert open(out / "abundance.tsv").read() == open(os.path.join(ref, "abundance.tsv")).read()
    same_run_info(out / "run_info.json", os.path.join(ref, "run_info.json"))


def test_cli_errors(tmp_path):
    r = run(["quant", "-i", "/nonexistent.kidx", "-o", str(tmp_path / "o"), "a.fq", "b.fq"])
    assert r.returncode == 1 and "kallisto index file not found" in r.stderr
    ds = util.dataset("config1")
    r = run(["quant", "-i", ds["index"], "-o", str(tmp_path / "o2"), os.path.join(ds["dir"], "reads_1.fastq.gz")])
    assert r.returncode == 1 and "paired-end mode requires an even number of input files" in r.stderr
    r = run(["quant", "-i", ds["index"], "-o", str(tmp_path / "o3"), "--single", os.path.join(ds["dir"], "reads_1.fastq.gz")])
    assert r.returncode == 1 and "fragment length mean and sd must be supplied" in r.stderr


@pytest.mark.parametrize("tag,args", [("10xv2", ["-x", "10xv2"]), ("10xv2_num", ["-x", "10xv2", "--num"]),
                                      ("10xv2", ["-x", "0,0,16:0,16,26:1,0,0", "--fr-stranded"])])
def test_bus_command(tag, args, tmp_path):
    d = os.path.join(util.GOLDEN, "bus10x")
    out = tmp_path / "o"
    r = run(["bus", "-i", os.path.join(util.GOLDEN, "config1", "transcripts.kidx"), "-o", str(out)] + args +
            [os.path.join(d, "sc_reads_1.fastq.gz"), os.path.join(d, "sc_reads_2.fastq.gz")])
    assert r.returncode == 0, r.stderr
    ref = os.path.join(d, "ref_" + tag)
    for fn in ("output.bus", "matrix.ec", "transcripts.txt"):
        assert open(out / fn, "rb").read() == open(os.path.join(ref, fn), "rb").read(), fn
    same_run_info(out / "run_info.json", os.path.join(ref, "run_info.json"))


def test_bus_nothing_aligned_exits_1(tmp_path):
    d = os.path.join(util.GOLDEN, "bus10x")
    out = tmp_path / "o"
    r = run(["bus", "-i", os.path.join(util.GOLDEN, "config1", "transcripts.kidx"), "-o", str(out), "-x", "10xv3",
             "--unstranded", os.path.join(d, "sc_reads_1.fastq.gz"), os.path.join(d, "sc_reads_2.fastq.gz")])
    assert r.returncode == 1
    ref = os.path.join(d, "ref_10xv3_unstranded")
    assert open(out / "output.bus", "rb").read() == open(os.path.join(ref, "output.bus"), "rb").read()
    same_run_info(out / "run_info.json", os.path.join(ref, "run_info.json"))


def _functest_cases():
    d = os.path.join(util.GOLDEN, "functests")
    return d, json.load(open(os.path.join(d, "cases.json")))


@pytest.mark.parametrize("case", _functest_cases()[1], ids=lambda c: c["name"])
def test_reference_functests_md5(case, tmp_path):
    """The quant cases of the reference's own func_tests/runtests.sh:265-304, run through the
    kallisto_b200 command line: abundance.tsv must have the md5 that script pins."""
    import hashlib
    d, _ = _functest_cases()
    out = tmp_path / "o"
    r = run(["quant", "-o", str(out), "-i", os.path.join(d, case["index"])] + case["args"] +
            [os.path.join(d, f) for f in case["files"]])
    assert r.returncode == 0, r.stderr
    assert hashlib.md5(open(out / "abundance.tsv", "rb").read()).hexdigest() == case["md5"]
    same_run_info(out / "run_info.json", os.path.join(d, case["name"], "run_info.json"))


def test_quant_single_with_position_filter(tmp_path):
    ds = util.dataset("synth_small")
    out = tmp_path / "o"
    r = run(["quant", "-i", ds["index"], "-o", str(out), "--plaintext", "--single", "-l", "200", "-s", "20",
             os.path.join(ds["dir"], "reads_1.fastq.gz")])
    assert r.returncode == 0, r.stderr
    ref = os.path.join(ds["dir"], "ref_quant_single")
    assert open(out / "abundance.tsv").read() == open(os.path.join(ref, "abundance.tsv")).read()
    same_run_info(out / "run_info.json", os.path.join(ref, "run_info.json"))


def test_quant_parallel_ingest_plain_fastq(tmp_path, monkeypatch):
    """-t 8 on uncompressed files takes the mapped, multi-threaded reader (csrc/fastx.hpp ParallelFastx); tiny
    windows force many segment cuts and unequal batch cuts between the two files (LockStep rounds)."""
    import gzip
    ds = util.dataset("synth_small")
    files = []
    for m in (1, 2):
        p = tmp_path / ("r%d.fq" % m)
        p.write_bytes(gzip.open(os.path.join(ds["dir"], "reads_%d.fastq.gz" % m)).read())
        files.append(str(p))
    monkeypatch.setenv("KB_FASTX_WINDOW", "30000")
    monkeypatch.setenv("KB_CLI_BATCH_READS", "700,1100")      # the two files cut their batches differently
    out = tmp_path / "o"
    r = run(["quant", "-i", ds["index"], "-o", str(out), "--plaintext", "-t", "8", "-b", "3", "--seed", "42"] + files)
    assert r.returncode == 0, r.stderr
    ref = os.path.join(ds["dir"], "ref_quant_paired")
    for fn in ["abundance.tsv", "bs_abundance_0.tsv", "bs_abundance_1.tsv", "bs_abundance_2.tsv"]:
        assert open(out / fn).read() == open(os.path.join(ref, fn)).read(), fn
    same_run_info(out / "run_info.json", os.path.join(ref, "run_info.json"))


def test_bus_unequal_batch_cuts(tmp_path, monkeypatch):
    """The barcode file and the cDNA file cut their batches at different read counts: records still come out in
    read order with the same EC ids."""
    monkeypatch.setenv("KB_CLI_BATCH_READS", "300,470")
    d = os.path.join(util.GOLDEN, "bus10x")
    out = tmp_path / "o"
    r = run(["bus", "-i", os.path.join(util.GOLDEN, "config1", "transcripts.kidx"), "-o", str(out), "-x", "10xv2", "-t", "4",
             os.path.join(d, "sc_reads_1.fastq.gz"), os.path.join(d, "sc_reads_2.fastq.gz")])
    assert r.returncode == 0, r.stderr
    ref = os.path.join(d, "ref_10xv2")
    for fn in ("output.bus", "matrix.ec", "transcripts.txt"):
        assert open(out / fn, "rb").read() == open(os.path.join(ref, fn), "rb").read(), fn


@pytest.mark.parametrize("extra", [[], ["--fr-stranded"], ["--single", "-l", "200", "-s", "20"]], ids=["paired", "fr", "single"])
def test_quant_devices_equals_one_device(extra, tmp_path, monkeypatch):
    """`--devices a,b,c`: batches dealt to several runs (here three runs on the one GPU of the test box), merged by content
    through peer copies, numbered by global fragment index -- every output byte as with one run, bootstraps included, and
    equal to the reference's files."""
    ds = util.dataset("synth_small")
    monkeypatch.setenv("KB_CLI_BATCH_READS", "1700")          # many small batches, so that every run gets some
    files = [os.path.join(ds["dir"], "reads_1.fastq.gz")] + ([] if "--single" in extra else [os.path.join(ds["dir"], "reads_2.fastq.gz")])
    outs = []
    for tag, dv in (("one", ["--device", "0"]), ("three", ["--devices", "0,0,0"])):
        out = tmp_path / tag
        r = run(["quant", "-i", ds["index"], "-o", str(out), "--plaintext", "-b", "2", "-t", "4"] + dv + extra + files)
        assert r.returncode == 0, r.stderr
        outs.append([open(out / f).read() for f in ("abundance.tsv", "bs_abundance_0.tsv", "bs_abundance_1.tsv")])
        if tag == "three":
            same_run_info(out / "run_info.json", tmp_path / "one" / "run_info.json")
    assert outs[0] == outs[1]
    ref = {"paired": "ref_quant_paired", "fr": "ref_quant_paired_fr", "single": "ref_quant_single"}["single" if "--single" in extra else ("fr" if extra else "paired")]
    assert outs[0][0] == open(os.path.join(ds["dir"], ref, "abundance.tsv")).read()


def test_quant_on_a_dlist_index(tmp_path):
    """An index built with `kallisto index -d`: fragments that hold a distinguishing flanking k-mer are discarded."""
    ds = util.dataset("dlist")
    out = tmp_path / "o"
    r = run(["quant", "-i", ds["index"], "-o", str(out), "--plaintext", "-b", "3", "--seed", "42",
             os.path.join(ds["dir"], "reads_1.fastq.gz"), os.path.join(ds["dir"], "reads_2.fastq.gz")])
    assert r.returncode == 0, r.stderr
    ref = os.path.join(ds["dir"], "ref_quant_paired")
    for fn in ["abundance.tsv", "bs_abundance_0.tsv", "bs_abundance_1.tsv", "bs_abundance_2.tsv"]:
        assert open(out / fn).read() == open(os.path.join(ref, fn)).read(), fn
    same_run_info(out / "run_info.json", os.path.join(ref, "run_info.json"))
