"""GPU: edge cases of the read path against the oracle -- ragged and empty reads, reads shorter than k,
Ns and other non-ACGT characters, lowercase, empty batches, odd batch sizes."""
import numpy as np
import pytest

import kallisto_b200 as K
from oracle import oracle as O
from tests import util

pytestmark = pytest.mark.gpu


def mangle(seqs, rng):
    out = []
    for s in seqs:
        r = rng.random()
        s = bytearray(s)
        if r < 0.15:
            s = s[: int(rng.integers(0, 40))]                 # empty / shorter than k / barely k
        elif r < 0.30:
            s = s[int(rng.integers(0, 60)):]
        elif r < 0.45:
            for _ in range(int(rng.integers(1, 6))):
                if s:
                    s[int(rng.integers(0, len(s)))] = ord("N")
        elif r < 0.55:
            s = bytearray(bytes(s).lower())
        elif r < 0.60:
            if s:
                s[int(rng.integers(0, len(s)))] = int(rng.choice(list(b".-*RYUn@")))
        elif r < 0.65:
            s = bytearray(b"N" * len(s))
        out.append(bytes(s))
    return out


@pytest.mark.parametrize("paired,strand", [(True, 0), (False, 0), (True, 1), (False, 2)])
def test_ragged_reads_match_oracle(paired, strand):
    ds = util.dataset("synth_small")
    rng = np.random.default_rng(7 + int(paired) + strand)
    s1 = mangle(ds["s1"][:6000], rng)
    s2 = mangle(ds["s2"][:6000], rng)
    bases, off = O.to_batch(s1, s2 if paired else None)
    ix = K.KmerIndex(ds["index"], device=0)
    mc = K.MinCollector(ix, paired=paired, strand=strand)
    h = mc.process_buffer(bases, off)
    eo, et, ec, eh = mc.ec_table()
    o_run = O.OracleRun(O.OracleIndex(ds["index"]), paired, strand, True)
    ofrag = o_run.pseudoalign(bases, off)
    oo, ot, oc = o_run.ec_table()
    np.testing.assert_array_equal(util.handles_to_ids(h, eh), ofrag)
    assert util.ec_sets(eo, et) == util.ec_sets(oo, ot)
    np.testing.assert_array_equal(ec, oc)
    if paired:
        np.testing.assert_array_equal(mc.flens, o_run.flens())
    st = mc.finalize()
    # paired reads (partial=false): the kernel executes exactly the lookups of the reference's match();
    # the oracle's counter also holds mapPair's linear scans, which are free on the device.  Single-end
    # match() runs with partial=true and may stop early once the running intersection empties
    # (KmerIndex.cpp:1032-1046) - the device keeps probing and gets the same (empty) result later.
    if paired:
        assert st["n_probes"] <= o_run.n_find()
    mc.close(); ix.close()


def test_empty_and_tiny_batches():
    ds = util.dataset("config1")
    ix = K.KmerIndex(ds["index"], device=0)
    mc = K.MinCollector(ix, paired=True)
    assert len(mc.process_buffer(np.zeros(0, np.uint8), np.zeros(1, np.uint32))) == 0
    h = mc.process_buffer(*O.to_batch(ds["s1"][:1], ds["s2"][:1]))
    assert len(h) == 1
    h = mc.process_buffer(*O.to_batch([b"", b"ACGT"], [b"", b""]))
    assert list(h) == [-1, -1]
    st = mc.finalize()
    assert st["n_processed"] == 3
    with pytest.raises(K.KallistoB200Error):      # odd number of reads in a paired run
        mc.process_buffer(*O.to_batch(ds["s1"][:3]))
    mc.close(); ix.close()


def test_long_reads_use_bigger_tiles():
    """250-bp reads: more shared memory per lane, same results as the oracle."""
    ds = util.dataset("synth_small")
    rng = np.random.default_rng(3)
    # concatenate read pairs into long pseudo-reads (not biologically meaningful, just long)
    s1 = [a + b[:150] for a, b in zip(ds["s1"][:3000], ds["s1"][3000:6000])]
    s2 = [a + b[:150] for a, b in zip(ds["s2"][:3000], ds["s2"][3000:6000])]
    bases, off = O.to_batch(s1, s2)
    ix = K.KmerIndex(ds["index"], device=0)
    mc = K.MinCollector(ix, paired=True)
    h = mc.process_buffer(bases, off)
    eo, et, ec, eh = mc.ec_table()
    o_run = O.OracleRun(O.OracleIndex(ds["index"]), True, 0, True)
    np.testing.assert_array_equal(util.handles_to_ids(h, eh), o_run.pseudoalign(bases, off))
    mc.close(); ix.close()


@pytest.mark.parametrize("name", ["synth_small", "manyecs"])
def test_lookup_count_equals_the_reference_match(name):
    """Paired reads without fragment-length sampling (no mapPair scans): the kernel must execute -- or, inside a run of
    misses, answer from the presence filter -- exactly the k-mer lookups of KmerIndex::match, not one more or less."""
    ds = util.dataset(name)
    bases, off = util.batch(ds, True)
    ix = K.KmerIndex(ds["index"], device=0)
    mc = K.MinCollector(ix, paired=True, collect_fld=False)
    mc.process_buffer(bases, off, want_handles=False)
    st = mc.finalize()
    o_run = O.OracleRun(O.OracleIndex(ds["index"]), True, 0, False)
    o_run.pseudoalign(bases, off)
    assert st["n_probes"] == o_run.n_find()
    assert st["n_slot_visits"] < st["n_probes"]        # the presence filter answers part of them without touching the table
    mc.close(); ix.close()
