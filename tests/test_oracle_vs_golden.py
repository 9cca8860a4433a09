"""CPU: the oracle restatement (oracle/kb_oracle.cpp) against the fixtures produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import util


@pytest.mark.parametrize("name", ["config1", "synth_small", "manyecs", "abundant", "dlist"])
@pytest.mark.parametrize("mode", list(util.MODES))
def test_per_fragment_ecs_match_reference(name, mode):
    ds = util.dataset(name)
    paired, strand, _ = util.MODES[mode]
    g = util.golden_ecs(ds, mode)
    ix = O.OracleIndex(ds["index"])
    run = O.OracleRun(ix, paired, strand, collect_fld=True)
    bases, off = util.batch(ds, paired)
    frag = run.pseudoalign(bases, off)
    assert len(frag) == int(g["n_processed"])
    np.testing.assert_array_equal(frag, g["frag_ec"])          # bit-exact, ids in first-occurrence order
    eo, et, ec = run.ec_table()
    assert util.ec_sets(eo, et) == util.ec_sets(g["ec_off"], g["ec_tids"])
    assert int(ec.sum()) == int(g["n_pseudoaligned"])
    if paired and "flens" in g.files:
        np.testing.assert_array_equal(run.flens(), g["flens"])


@pytest.mark.parametrize("name", ["config1", "synth_small", "manyecs", "abundant", "dlist"])
def test_quant_text_identical_to_reference(name):
    """abundance.tsv of `kallisto quant --plaintext -t 1`, byte for byte (6 significant digits)."""
    ds = util.dataset(name)
    ix = O.OracleIndex(ds["index"])
    run = O.OracleRun(ix, True, 0, True)
    bases, off = util.batch(ds, True)
    run.pseudoalign(bases, off)
    eo, et, ec = run.ec_table()
    fl = O.mean_fl_trunc(run.flens())
    eff = O.eff_lens(ix.target_lens, fl)
    alpha, rounds = O.em(eo, et, ec, eff, ix.n_targets)
    txt = O.abundance_tsv(ix.target_names, ix.target_lens, eff, alpha, O.tpm(alpha, eff))
    ref = open(os.path.join(ds["dir"], "ref_quant_paired", "abundance.tsv")).read()
    assert txt == ref
    if name == "config1":
        # the number quoted in SURVEY.md 8(c) / BASELINE.md
        assert hashlib.md5(txt.encode()).hexdigest() == "0bd5087aba9db4b681073bb84de3fe5f"
        assert rounds == 52


@pytest.mark.parametrize("name", ["config1", "synth_small", "manyecs", "abundant", "dlist"])
def test_bootstrap_text_identical_to_reference(name):
    ds = util.dataset(name)
    ix = O.OracleIndex(ds["index"])
    run = O.OracleRun(ix, True, 0, True)
    bases, off = util.batch(ds, True)
    run.pseudoalign(bases, off)
    eo, et, ec = run.ec_table()
    fl = O.mean_fl_trunc(run.flens())
    eff = O.eff_lens(ix.target_lens, fl)
    for b in range(3):
        samp = O.bootstrap_sample(ec, 42, b)
        assert samp.sum() == ec.sum()
        alpha, _ = O.em(eo, et, samp, eff, ix.n_targets, counts_w=ec)
        txt = O.abundance_tsv(ix.target_names, ix.target_lens, eff, alpha, O.tpm(alpha, eff))
        ref = open(os.path.join(ds["dir"], "ref_quant_paired", "bs_abundance_%d.tsv" % b)).read()
        assert txt == ref


def test_single_overhang_quant_matches_reference():
    """--single --single-overhang -l 200 -s 20: truncated-Gaussian effective lengths."""
    ds = util.dataset("synth_small")
    ix = O.OracleIndex(ds["index"])
    run = O.OracleRun(ix, False, 0, False)
    bases, off = util.batch(ds, False)
    run.pseudoalign(bases, off)
    eo, et, ec = run.ec_table()
    fl = O.mean_fl_trunc(np.zeros(1000, np.uint32), 200.0, 20.0)
    eff = O.eff_lens(ix.target_lens, fl)
    alpha, _ = O.em(eo, et, ec, eff, ix.n_targets)
    txt = O.abundance_tsv(ix.target_names, ix.target_lens, eff, alpha, O.tpm(alpha, eff))
    ref = open(os.path.join(ds["dir"], "ref_quant_single_overhang", "abundance.tsv")).read()
    assert txt == ref


def test_single_end_with_position_filter_matches_reference():
    """--single -l 200 -s 20 (no --single-overhang): KmerIndex::findPosition filter."""
    import json
    ds = util.dataset("synth_small")
    ix = O.OracleIndex(ds["index"])
    run = O.OracleRun(ix, False, 0, False, fp_fl=200)
    bases, off = util.batch(ds, False)
    frag = run.pseudoalign(bases, off)
    eo, et, ec = run.ec_table()
    info = json.load(open(os.path.join(ds["dir"], "ref_quant_single", "run_info.json")))
    assert int(ec.sum()) == info["n_pseudoaligned"]
    assert int(sum(c for c, a, b in zip(ec, eo[:-1], eo[1:]) if b - a == 1)) == info["n_unique"]
    # the filter must actually do something on this data set
    plain = O.OracleRun(ix, False, 0, False)
    assert not np.array_equal(plain.pseudoalign(bases, off), frag) or plain.ec_table()[2].sum() != ec.sum()
    fl = O.mean_fl_trunc(np.zeros(1000, np.uint32), 200.0, 20.0)
    eff = O.eff_lens(ix.target_lens, fl)
    alpha, _ = O.em(eo, et, ec, eff, ix.n_targets)
    txt = O.abundance_tsv(ix.target_names, ix.target_lens, eff, alpha, O.tpm(alpha, eff))
    assert txt == open(os.path.join(ds["dir"], "ref_quant_single", "abundance.tsv")).read()


def _functest_cases():
    import json
    d = os.path.join(util.GOLDEN, "functests")
    return d, json.load(open(os.path.join(d, "cases.json")))


@pytest.mark.parametrize("case", _functest_cases()[1], ids=lambda c: c["name"])
def test_functests_md5_goldens(case):
    """The quant cases of the reference's own func_tests/runtests.sh:265-304 (toy k=5/7/11 indices,
    N bases, lowercase, poly-A clipping, strandedness, the findPosition filter): the oracle must
    reproduce the md5 the script pins for abundance.tsv."""
    d, _ = _functest_cases()
    ix = O.OracleIndex(os.path.join(d, case["index"]))
    single = "--single" in case["args"]
    strand = 1 if "--fr-stranded" in case["args"] else (2 if "--rf-stranded" in case["args"] else 0)
    files = [os.path.join(d, f) for f in case["files"]]
    if single:
        run = O.OracleRun(ix, False, strand, False, fp_fl=5)
        for f in files:
            run.pseudoalign(*O.to_batch(O.read_fastq(f)))
        fl = O.mean_fl_trunc(np.zeros(1000, np.uint32), 5.0, 2.0)
    else:
        run = O.OracleRun(ix, True, strand, True)
        for i in range(0, len(files), 2):
            run.pseudoalign(*O.to_batch(O.read_fastq(files[i]), O.read_fastq(files[i + 1])))
        fl = O.mean_fl_trunc(run.flens())
    eo, et, ec = run.ec_table()
    eff = O.eff_lens(ix.target_lens, fl)
    alpha, _ = O.em(eo, et, ec, eff, ix.n_targets)
    txt = O.abundance_tsv(ix.target_names, ix.target_lens, eff, alpha, O.tpm(alpha, eff))
    assert hashlib.md5(txt.encode()).hexdigest() == case["md5"]


def test_abundant_fixture_has_abundant_unitigs():
    """The `abundant` data set exists to exercise the abundant-unitig branch of CompactedDBG::find
    (ext/bifrost/src/CompactedDBG.tcc:1072-1119): the index must really contain some."""
    n_long, n_short, n_abund = O.index_unitig_kinds(util.dataset("abundant")["index"])
    assert n_abund > 100 and n_short > 100
