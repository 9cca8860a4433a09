"""GPU: `kallisto bus` with two sequence reads and / or one sample per file -- `-x bulk [--paired]` (batch mode:
src/main.cpp:1050-1107, src/ProcessReads.cpp:371-404,1603-1607), SMARTSEQ2 (no UMI: "bulk_like", :1393) and a
STORM-seq-like custom technology (second mate from base 14) -- through the library (kb_bus_create with paired / seq2,
kb_bus_begin_sample) and through the command line, against the files the unmodified reference wrote
(tests/golden/buspaired).  The reference writes the records of a batch whose ECs are already known first
(:1798-1812, 603-612), so records are compared as sorted multisets plus in order where the order is defined."""
import json
import os
import subprocess

import numpy as np
import pytest

import kallisto_b200 as K
from oracle import oracle as O
from tests import util
from tests.test_oracle_bus_paired import SPECS, case_files, read_ref, sorted_records

pytestmark = pytest.mark.gpu

BIN = os.path.join(util.ROOT, "kallisto_b200", "kallisto_b200")
D = os.path.join(util.GOLDEN, "buspaired")


@pytest.fixture(scope="module")
def inputs(tmp_path_factory):
    return util.buspaired_inputs(str(tmp_path_factory.mktemp("buspaired_in")))


@pytest.fixture(scope="module")
def ix():
    x = K.KmerIndex(os.path.join(util.GOLDEN, "synth_small", "transcripts.kidx"), device=0)
    yield x
    x.close()


def technology(name):
    bc, umi, seq, seq2, strand, num, per_sample, tag = SPECS[name]
    nfiles = 1 + max([seq[0]] + ([seq2[0]] if seq2 else []) + [b[0] for b in bc] + [u[0] for u in (umi or [])])
    if tag:          # the library takes the UMI location as the user gives it: tag + UMI (SPECS holds it advanced by the tag)
        umi = [(umi[0][0], umi[0][1] - len(tag), umi[0][2])] + list(umi[1:])
    t = (nfiles, bc, umi if umi is not None else [(-1, -1, -1)], (seq[0], seq[1], 0), 0)
    if seq2:
        t = t + ((seq2[0], seq2[1], 0),)
    return t


@pytest.mark.parametrize("name", sorted(SPECS))
def test_library_records_ecs_flens_identical_to_reference(inputs, ix, name):
    d, hdr, ref, info, ref_ecs, ref_flens = read_ref(name)
    bc, umi, seq, seq2, strand, num, per_sample, tag = SPECS[name]
    files, samples = case_files(inputs, name)
    bp = K.BUSProcessor(ix, technology(name), strand={0: "unstranded", 1: "fr", 2: "rf"}[strand], num=num, tag=tag)
    parts, flens = [], []
    for si, (lo, hi) in enumerate(samples or [(0, len(files[0]))]):
        if per_sample:
            if si:
                flens.append(bp.flens)
            bp.begin_sample(si)
        mid = lo + (hi - lo) // 3                  # two batches per sample: ids, read numbers and quotas carry over
        for a, b in ((lo, mid), (mid, hi)):
            parts.append(bp.process_sets([O.to_batch(f[a:b]) for f in files]))
    flens.append(bp.flens)
    rec = np.concatenate(parts)
    assert len(rec) == len(ref) == info["n_pseudoaligned"]
    assert sorted_records(rec).tobytes() == sorted_records(ref).tobytes()
    n0 = int((ref["barcode"] == ref["barcode"][0]).sum()) if per_sample else len(ref)
    if per_sample:
        assert rec[:n0].tobytes() == ref[:n0].tobytes()          # first sample: every EC is new, read order
    st = bp.finalize()
    assert st["n_processed"] == info["n_processed"]
    assert st["n_pseudoaligned"] == info["n_pseudoaligned"]
    assert st["n_unique"] == info["n_unique"]
    eo, et, ec, eh = bp.ec_table()
    assert util.ec_sets(eo, et) == ref_ecs
    if ref_flens is not None:
        assert len(flens) == len(ref_flens)
        for a, b in zip(flens, ref_flens):
            np.testing.assert_array_equal(a, b)
    b_h, u_h = bp.lengths()
    if not per_sample:
        assert hdr["bclen"] == int(np.argmax(b_h)) and hdr["umilen"] == int(np.argmax(u_h))
    bp.close()


@pytest.mark.parametrize("name", sorted(SPECS))
def test_command_line_files_identical_to_reference(inputs, name, tmp_path):
    d, hdr, ref, info, ref_ecs, ref_flens = read_ref(name)
    args, keys = util.BUSPAIRED_CASES[name]
    out = tmp_path / "o"
    r = subprocess.run([BIN, "bus", "-i", os.path.join(util.GOLDEN, "synth_small", "transcripts.kidx"), "-o", str(out), "-t", "4"] +
                       args + [inputs[k] for k in keys], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    got_hdr, got = O.read_bus(str(out / "output.bus"))
    assert got_hdr == hdr
    assert sorted_records(got.copy()).tobytes() == sorted_records(ref).tobytes()
    want_files = sorted(f for f in os.listdir(d) if f not in ("output.bus.gz", "run_info.json"))
    assert sorted(f for f in os.listdir(out) if f not in ("output.bus", "run_info.json", "transcripts.txt")) == want_files
    for fn in want_files:          # matrix.ec, flens.txt, index.saved, matrix.cells, matrix.sample.barcodes
        assert open(out / fn, "rb").read() == open(os.path.join(d, fn), "rb").read(), fn
    ja, jb = json.load(open(out / "run_info.json")), info
    for k in ("n_targets", "n_processed", "n_pseudoaligned", "n_unique", "p_pseudoaligned", "p_unique", "index_version", "k-mer length"):
        assert ja[k] == jb[k], k


def test_batch_file_identical_to_reference(inputs, tmp_path):
    """`bus --batch FILE`: sample names from the file, lines with the same id share a barcode; every file the reference
    writes (records as a sorted multiset)."""
    d, hdr, ref, info, ref_ecs, ref_flens = read_ref("batchfile")
    out = tmp_path / "o"
    bf = util.write_batch_file(str(tmp_path / "batch.txt"), inputs)
    r = subprocess.run([BIN, "bus", "-i", os.path.join(util.GOLDEN, "synth_small", "transcripts.kidx"), "-o", str(out), "-t", "4", "--batch", bf],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    got_hdr, got = O.read_bus(str(out / "output.bus"))
    assert got_hdr == hdr
    assert sorted_records(got.copy()).tobytes() == sorted_records(ref).tobytes()
    for fn in ("matrix.ec", "flens.txt", "index.saved", "matrix.cells", "matrix.sample.barcodes"):
        assert open(out / fn, "rb").read() == open(os.path.join(d, fn), "rb").read(), fn
    ja = json.load(open(out / "run_info.json"))
    for k in ("n_targets", "n_processed", "n_pseudoaligned", "n_unique"):
        assert ja[k] == info[k], k


def test_index_saved_of_a_dlist_index(tmp_path):
    dl = os.path.join(util.GOLDEN, "dlist")
    out = tmp_path / "o"
    r = subprocess.run([BIN, "bus", "-i", os.path.join(dl, "transcripts.kidx"), "-o", str(out), "-x", "bulk", "--paired",
                        os.path.join(dl, "reads_1.fastq.gz"), os.path.join(dl, "reads_2.fastq.gz")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out / "index.saved", "rb").read() == open(os.path.join(D, "dlist_index.saved"), "rb").read()


def test_paired_flag_needs_a_paired_technology(tmp_path):
    r = subprocess.run([BIN, "bus", "-i", os.path.join(util.GOLDEN, "synth_small", "transcripts.kidx"), "-o", str(tmp_path / "o"),
                        "-x", "10xv3", "--paired", os.path.join(util.GOLDEN, "bus10xv3", "sc_reads_1.fastq.gz"),
                        os.path.join(util.GOLDEN, "bus10xv3", "sc_reads_2.fastq.gz")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1
    assert "Error: Paired reads are not compatible with the specified technology" in r.stderr




def test_quant_tcc_runs_on_index_saved(tmp_path):
    """index.saved (targets only, no graph) is what kb-python hands to `quant-tcc` after `bus`: the reference gives the
    same files with it as with the full index (checked when the fixture was made), and so must this build."""
    q = os.path.join(util.GOLDEN, "quanttcc")
    saved = os.path.join(D, "ref_bulk_paired", "index.saved")
    out = tmp_path / "o"
    r = subprocess.run([BIN, "quant-tcc", "-i", saved, "-e", os.path.join(q, "matrix.ec"), "-o", str(out), "-t", "2", "-l", "200",
                        "-s", "20", os.path.join(q, "tcc.mtx")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    ref = os.path.join(q, "ref_ls")
    assert sorted(os.listdir(out)) == sorted(os.listdir(ref))
    for fn in os.listdir(ref):
        assert open(out / fn, "rb").read() == open(os.path.join(ref, fn), "rb").read(), fn
    # ... and nothing else can: there are no k-mers to pseudoalign against
    ixs = K.KmerIndex(saved, device=0)
    with pytest.raises(K.KallistoB200Error):
        K.MinCollector(ixs, paired=True)
    ixs.close()


def test_quant_without_plaintext_writes_abundance_h5(tmp_path):
    """`quant -b 3` without --plaintext: abundance.h5 (csrc/h5_writer.hpp, read back by tests/h5mini.py) holds the run the
    reference's text files describe -- est_counts / eff_lengths of abundance.tsv and the three bootstraps of
    bs_abundance_{0,1,2}.tsv (6 significant digits in the text) -- and the fragment-length histogram of the run."""
    from tests import h5mini
    ds = util.dataset("config1")
    out = tmp_path / "o"
    r = subprocess.run([BIN, "quant", "-i", ds["index"], "-o", str(out), "-b", "3", "--seed", "42", os.path.join(ds["dir"], "reads_1.fastq.gz"),
                        os.path.join(ds["dir"], "reads_2.fastq.gz")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    ref = os.path.join(ds["dir"], "ref_quant_paired")
    assert open(out / "abundance.tsv").read() == open(os.path.join(ref, "abundance.tsv")).read()
    assert not os.path.exists(out / "bs_abundance_0.tsv")
    h = h5mini.read(str(out / "abundance.h5"))
    names, lens, eff, est, tpm = util.read_abundance(os.path.join(ref, "abundance.tsv"))
    assert h["aux"]["ids"] == names and np.array_equal(h["aux"]["lengths"], lens)
    np.testing.assert_allclose(h["est_counts"], est, rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(h["aux"]["eff_lengths"], eff, rtol=1e-5)
    for b in range(3):
        _, _, _, est_b, _ = util.read_abundance(os.path.join(ref, "bs_abundance_%d.tsv" % b))
        np.testing.assert_allclose(h["bootstrap"]["bs%d" % b], est_b, rtol=1e-5, atol=1e-12)
    g = util.golden_ecs(ds, "paired")
    np.testing.assert_array_equal(h["aux"]["fld"], g["flens"].astype(np.int32))
    assert list(h["aux"]["num_processed"]) == [10000] and list(h["aux"]["num_bootstrap"]) == [3]
    # ... and back: h5dump turns the file into the reference's own text files, byte for byte (the doubles are all there)
    dump = tmp_path / "dump"
    r = subprocess.run([BIN, "h5dump", "-o", str(dump), str(out / "abundance.h5")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    for fn in ["abundance.tsv", "bs_abundance_0.tsv", "bs_abundance_1.tsv", "bs_abundance_2.tsv"]:
        assert open(dump / fn).read() == open(os.path.join(ref, fn)).read(), fn


def test_quant_write_index(tmp_path):
    """`quant --write-index`: counts.txt (reads per equivalence class, ids of first occurrence = the reference's with -t 1)
    and index.saved identical to the reference's files (tests/golden/config1/ref_quant_paired)."""
    ds = util.dataset("config1")
    out = tmp_path / "o"
    r = subprocess.run([BIN, "quant", "-i", ds["index"], "-o", str(out), "--plaintext", "--write-index", os.path.join(ds["dir"], "reads_1.fastq.gz"),
                        os.path.join(ds["dir"], "reads_2.fastq.gz")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    ref = os.path.join(ds["dir"], "ref_quant_paired")
    for fn in ("counts.txt", "index.saved", "abundance.tsv"):
        assert open(out / fn, "rb").read() == open(os.path.join(ref, fn), "rb").read(), fn


def test_bus_interleaved_input(tmp_path):
    """`bus --inleaved`: one interleaved FASTQ file instead of the two of 10x v2 -> the reference's output.bus, byte for byte."""
    import gzip
    d = os.path.join(util.GOLDEN, "bus10x")
    a, b = (gzip.open(os.path.join(d, "sc_reads_%d.fastq.gz" % m), "rb").read().split(b"\n") for m in (1, 2))
    il = tmp_path / "il.fq"
    with open(il, "wb") as f:
        for i in range(len(a) // 4):
            f.write(b"\n".join(a[4 * i:4 * i + 4]) + b"\n" + b"\n".join(b[4 * i:4 * i + 4]) + b"\n")
    out = tmp_path / "o"
    r = subprocess.run([BIN, "bus", "-i", os.path.join(util.GOLDEN, "config1", "transcripts.kidx"), "-o", str(out), "-x", "10xv2", "-t", "4",
                        "--inleaved", str(il)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    for fn in ("output.bus", "matrix.ec"):
        assert open(out / fn, "rb").read() == open(os.path.join(d, "ref_10xv2", fn), "rb").read(), fn
