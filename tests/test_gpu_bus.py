"""GPU: the BUS record path (kb_bus_batch) against output.bus / matrix.ec written by the
unmodified reference `kallisto bus -t 1` (tests/golden/bus10x)."""
import json
import os

import numpy as np
import pytest

import kallisto_b200 as K
from oracle import oracle as O
from tests import util

pytestmark = pytest.mark.gpu

D = os.path.join(util.GOLDEN, "bus10x")


@pytest.fixture(scope="module")
def setup():
    ix = K.KmerIndex(os.path.join(util.GOLDEN, "config1", "transcripts.kidx"), device=0)
    s1 = O.read_fastq(os.path.join(D, "sc_reads_1.fastq.gz"))
    s2 = O.read_fastq(os.path.join(D, "sc_reads_2.fastq.gz"))
    yield ix, O.to_batch(s1), O.to_batch(s2), (s1, s2)
    ix.close()


@pytest.mark.parametrize("tag,tech,kw", [("10xv2", "10xv2", {}), ("10xv2_num", "10xv2", {"num": True}),
                                          ("10xv3_unstranded", "10xv3", {"strand": "unstranded"})])
def test_bus_records_identical_to_reference(setup, tag, tech, kw):
    ix, f1, f2, _ = setup
    hdr, ref = O.read_bus(os.path.join(D, "ref_" + tag, "output.bus"))
    bp = K.BUSProcessor(ix, tech, **kw)
    rec = bp.process_sets([f1, f2])
    assert len(rec) == len(ref)
    # byte-identical records, in read order, EC ids included (-t 1 numbering)
    assert rec.tobytes() == ref.tobytes()
    st = bp.finalize()
    info = json.load(open(os.path.join(D, "ref_" + tag, "run_info.json")))
    assert st["n_processed"] == info["n_processed"]
    assert st["n_pseudoaligned"] == info["n_pseudoaligned"]
    assert st["n_unique"] == info["n_unique"]
    eo, et, ec, eh = bp.ec_table()
    ref_ecs = O.read_matrix_ec(os.path.join(D, "ref_" + tag, "matrix.ec")) if info["n_pseudoaligned"] else []
    assert util.ec_sets(eo, et) == ref_ecs
    bc, um = bp.lengths()
    if tech == "10xv2":
        assert bc[16] == 10000 and um[10] == 10000
    bp.close()


def test_bus_batches_concatenate(setup):
    ix, f1, f2, (s1, s2) = setup
    _, ref = O.read_bus(os.path.join(D, "ref_10xv2", "output.bus"))
    bp = K.BUSProcessor(ix, "10xv2")
    parts = []
    for a, b in ((0, 3), (3, 2500), (2500, 10000)):
        parts.append(bp.process_sets([O.to_batch(s1[a:b]), O.to_batch(s2[a:b])]))
    assert np.concatenate(parts).tobytes() == ref.tobytes()
    bp.close()


def test_bus_short_barcode_reads_are_skipped(setup):
    ix, f1, f2, (s1, s2) = setup
    # truncate some R1 reads below barcode+UMI length: the reference drops those sets (ProcessReads.cpp:1505-1521)
    s1b = list(s1)
    for i in range(0, 200, 7):
        s1b[i] = s1b[i][:20]
    bp = K.BUSProcessor(ix, "10xv2", num=True)
    rec = bp.process_sets([O.to_batch(s1b), f2])
    full = K.BUSProcessor(ix, "10xv2", num=True)
    rec_full = full.process_sets([f1, f2])
    dropped = set(range(0, 200, 7))
    keep = np.array([r not in dropped for r in rec_full["flags"]])
    np.testing.assert_array_equal(rec["flags"], rec_full["flags"][keep])
    np.testing.assert_array_equal(rec["barcode"], rec_full["barcode"][keep])
    bp.close(); full.close()


# ---- 10x v3 layout with mapped records (tests/golden/bus10xv3: 16-nt barcode + 12-nt UMI, 91-nt cDNA on the
#      synth_small index; default --fr-stranded keeps the sense reads, --rf-stranded the antisense ones) ----
D3 = os.path.join(util.GOLDEN, "bus10xv3")


@pytest.fixture(scope="module")
def setup_v3():
    ix = K.KmerIndex(os.path.join(util.GOLDEN, "synth_small", "transcripts.kidx"), device=0)
    s1 = O.read_fastq(os.path.join(D3, "sc_reads_1.fastq.gz"))
    s2 = O.read_fastq(os.path.join(D3, "sc_reads_2.fastq.gz"))
    yield ix, s1, s2
    ix.close()


@pytest.mark.parametrize("tag,kw", [("10xv3", {}), ("10xv3_num", {"num": True}), ("10xv3_unstranded", {"strand": "unstranded"}),
                                     ("10xv3_rf", {"strand": "rf"})])
def test_bus_10xv3_records_identical_to_reference(setup_v3, tag, kw):
    ix, s1, s2 = setup_v3
    _, ref = O.read_bus(os.path.join(D3, "ref_" + tag, "output.bus"))
    info = json.load(open(os.path.join(D3, "ref_" + tag, "run_info.json")))
    assert info["n_pseudoaligned"] > 1000          # the fixture has mapped records in every mode
    bp = K.BUSProcessor(ix, "10xv3", **kw)
    parts = []
    for a, b in ((0, 5000), (5000, len(s1))):      # two batches: EC ids continue across batches
        parts.append(bp.process_sets([O.to_batch(s1[a:b]), O.to_batch(s2[a:b])]))
    rec = np.concatenate(parts)
    assert len(rec) == len(ref)
    assert rec.tobytes() == ref.tobytes()
    st = bp.finalize()
    assert st["n_processed"] == info["n_processed"]
    assert st["n_pseudoaligned"] == info["n_pseudoaligned"]
    assert st["n_unique"] == info["n_unique"]
    eo, et, ec, eh = bp.ec_table()
    assert util.ec_sets(eo, et) == O.read_matrix_ec(os.path.join(D3, "ref_" + tag, "matrix.ec"))
    bc, um = bp.lengths()
    assert bc[16] == um[12] and bc[16] >= info["n_processed"] - 40
    bp.close()
