"""GPU: the CUDA path through the C ABI against (a) the fixtures of the unmodified reference and
(b) the CPU oracle on the same inputs.  Integer work (per-fragment ECs, EC tables, counts,
fragment-length histogram, resampled bootstrap counts) must be bit-exact; est_counts / TPM /
eff_length within 1e-4 relative (BASELINE north_star) -- in fact we check the text too."""
import os

import numpy as np
import pytest

import kallisto_b200 as K
from oracle import oracle as O
from tests import util

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4   # north_star tolerance for floating point outputs


@pytest.fixture(scope="module")
def indices():
    out = {}
    for name in ("config1", "synth_small", "manyecs", "abundant", "dlist"):
        out[name] = K.KmerIndex(util.dataset(name)["index"], device=0)
    yield out
    for ix in out.values():
        ix.close()


@pytest.mark.parametrize("name", ["config1", "synth_small", "manyecs", "abundant", "dlist"])
@pytest.mark.parametrize("mode", list(util.MODES))
def test_per_fragment_ecs(indices, name, mode):
    ds = util.dataset(name)
    paired, strand, _ = util.MODES[mode]
    g = util.golden_ecs(ds, mode)
    mc = K.MinCollector(indices[name], paired=paired, strand=strand)
    bases, off = util.batch(ds, paired)
    h = mc.process_buffer(bases, off)
    st = mc.finalize()
    eo, et, ec, eh = mc.ec_table()
    assert st["n_processed"] == int(g["n_processed"])
    assert st["n_pseudoaligned"] == int(g["n_pseudoaligned"])
    assert st["n_unique"] == int(g["n_unique"])
    assert util.ec_sets(eo, et) == util.ec_sets(g["ec_off"], g["ec_tids"])
    np.testing.assert_array_equal(util.handles_to_ids(h, eh), g["frag_ec"])
    if paired and "flens" in g.files:
        np.testing.assert_array_equal(mc.flens, g["flens"])
    # same number of k-mer lookups as the CPU restatement executes for match() (mapPair's extra
    # linear scans are free on the device: its first hit is match()'s first hit)
    mc.close()


@pytest.mark.parametrize("name", ["config1", "synth_small", "manyecs", "abundant", "dlist"])
def test_batching_does_not_change_results(indices, name):
    ds = util.dataset(name)
    g = util.golden_ecs(ds, "paired")
    mc = K.MinCollector(indices[name], paired=True)
    n = len(ds["s1"])
    cuts = sorted({0, 1, 7, min(4000, n), min(4001, n), min(12000, n), n})
    hs = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        bases, off = O.to_batch(ds["s1"][a:b], ds["s2"][a:b])
        hs.append(mc.process_buffer(bases, off))
    eo, et, ec, eh = mc.ec_table()
    np.testing.assert_array_equal(util.handles_to_ids(np.concatenate(hs), eh), g["frag_ec"])
    np.testing.assert_array_equal(mc.flens, g["flens"])
    mc.close()


@pytest.mark.parametrize("name", ["config1", "synth_small", "manyecs", "abundant", "dlist"])
def test_quant_matches_reference(indices, name):
    ds = util.dataset(name)
    ix = indices[name]
    mc = K.MinCollector(ix, paired=True)
    bases, off = util.batch(ds, True)
    mc.process_buffer(bases, off, want_handles=False)
    r = mc.run_em()
    names, lens, eff, est, tpm = util.read_abundance(os.path.join(ds["dir"], "ref_quant_paired", "abundance.tsv"))
    assert names == ix.target_names_
    my_tpm = K.counts_to_tpm(r["est_counts"], r["eff_lens"])
    txt = O.abundance_tsv(ix.target_names_, ix.target_lens_, r["eff_lens"], r["est_counts"], my_tpm)
    ref_txt = open(os.path.join(ds["dir"], "ref_quant_paired", "abundance.tsv")).read()
    # tolerance gate (6 significant digits are printed, so compare against the text's own precision)
    np.testing.assert_allclose(r["eff_lens"], eff, rtol=REL_TOL)
    big = est > 1e-8 * est.sum()
    np.testing.assert_allclose(r["est_counts"][big], est[big], rtol=REL_TOL)
    np.testing.assert_allclose(my_tpm[big], tpm[big], rtol=REL_TOL)
    # the EM is accumulated in the reference's order without FMA: the text is identical
    assert txt == ref_txt
    # and bit-identical to the oracle restatement, iteration count included
    o_ix = O.OracleIndex(ds["index"])
    o_run = O.OracleRun(o_ix, True, 0, True)
    o_run.pseudoalign(bases, off)
    oo, ot, oc = o_run.ec_table()
    o_eff = O.eff_lens(o_ix.target_lens, O.mean_fl_trunc(o_run.flens()))
    o_alpha, o_rounds = O.em(oo, ot, oc, o_eff, o_ix.n_targets)
    assert r["rounds"] == o_rounds
    np.testing.assert_array_equal(r["eff_lens"], o_eff)
    np.testing.assert_array_equal(r["est_counts"], o_alpha)
    mc.close()


@pytest.mark.parametrize("name", ["config1", "synth_small", "manyecs", "abundant", "dlist"])
def test_bootstrap_matches_reference(indices, name):
    ds = util.dataset(name)
    ix = indices[name]
    mc = K.MinCollector(ix, paired=True)
    bases, off = util.batch(ds, True)
    mc.process_buffer(bases, off, want_handles=False)
    main = mc.run_em()
    eo, et, ec, eh = mc.ec_table()
    bs = mc.run_bootstrap(3, seed=42, want_samples=True)
    for b in range(3):
        np.testing.assert_array_equal(bs["samples"][b][: len(ec)], O.bootstrap_sample(ec, 42, b))   # bit-exact resampling
        tpm = K.counts_to_tpm(bs["est_counts"][b], main["eff_lens"])
        txt = O.abundance_tsv(ix.target_names_, ix.target_lens_, main["eff_lens"], bs["est_counts"][b], tpm)
        ref = open(os.path.join(ds["dir"], "ref_quant_paired", "bs_abundance_%d.tsv" % b)).read()
        assert txt == ref
    mc.close()


def test_single_overhang_quant(indices):
    ds = util.dataset("synth_small")
    ix = indices["synth_small"]
    mc = K.MinCollector(ix, paired=False, collect_fld=False)
    bases, off = util.batch(ds, False)
    mc.process_buffer(bases, off, want_handles=False)
    r = mc.run_em(fld_mean=200.0, fld_sd=20.0)
    tpm = K.counts_to_tpm(r["est_counts"], r["eff_lens"])
    txt = O.abundance_tsv(ix.target_names_, ix.target_lens_, r["eff_lens"], r["est_counts"], tpm)
    assert txt == open(os.path.join(ds["dir"], "ref_quant_single_overhang", "abundance.tsv")).read()
    mc.close()


def test_fixed_length_path_equals_offsets_path(indices):
    ds = util.dataset("synth_small")
    mc1 = K.MinCollector(indices["synth_small"], paired=True)
    mc2 = K.MinCollector(indices["synth_small"], paired=True)
    bases, off = util.batch(ds, True)
    assert np.all(np.diff(off.astype(np.int64)) == 100)
    h1 = mc1.process_buffer(bases, off)
    h2 = mc2.process_buffer(bases, None, fixed_len=100)
    np.testing.assert_array_equal(h1, h2)
    mc1.close(); mc2.close()


@pytest.mark.parametrize("mode", ["single", "single_fr", "paired_with_l"])
def test_fragment_position_filter_per_fragment(mode):
    """findPosition filter (ProcessReads.cpp:1095-1136): per-fragment ECs against the oracle's literal
    restatement (itself pinned on the reference's single-end abundance.tsv and functests md5s)."""
    ds = util.dataset("synth_small")
    ix = K.KmerIndex(ds["index"], device=0, load_positions=True)
    paired = mode == "paired_with_l"
    strand = 1 if mode == "single_fr" else 0
    mc = K.MinCollector(ix, paired=paired, strand=strand, collect_fld=False, single_overhang=False, fld_mean=200.0)
    bases, off = util.batch(ds, paired)
    h = mc.process_buffer(bases, off)
    eo, et, ec, eh = mc.ec_table()
    o_ix = O.OracleIndex(ds["index"])
    o_run = O.OracleRun(o_ix, paired, strand, False, fp_fl=200)
    ofrag = o_run.pseudoalign(bases, off)
    oo, ot, oc = o_run.ec_table()
    np.testing.assert_array_equal(util.handles_to_ids(h, eh), ofrag)
    assert util.ec_sets(eo, et) == util.ec_sets(oo, ot)
    np.testing.assert_array_equal(ec, oc)
    mc.close(); ix.close()


def test_position_filter_needs_positions():
    ds = util.dataset("synth_small")
    ix = K.KmerIndex(ds["index"], device=0, load_positions=False)
    with pytest.raises(K.KallistoB200Error):
        K.MinCollector(ix, paired=False, single_overhang=False, fld_mean=200.0)
    ix.close()
