"""CPU: the BUS record rules of paired / no-UMI / sample-per-file runs (`kallisto bus -x bulk [--paired]`, SMARTSEQ2,
a STORM-seq-like custom technology), restated in oracle/oracle.py:bus_model on top of the oracle's pseudoalignment,
against the outputs of the unmodified reference (tests/golden/buspaired, made by tests/golden/make_golden.py buspaired).
Pins what the GPU tests (tests/test_gpu_zz_bus_paired.py) then demand of the CUDA path."""
import gzip
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import util

D = os.path.join(util.GOLDEN, "buspaired")

# name -> bus_model arguments: (bc, umi, seq, seq2, strand, num, per-sample, tag)
SPECS = {
    "bulk_paired": ([], None, (0, 0), (1, 0), 0, False, True, None),
    "bulk_paired_num_fr": ([], None, (0, 0), (1, 0), 1, True, True, None),
    "bulk_single": ([], None, (0, 0), None, 0, False, True, None),
    "smartseq2_paired": ([(0, 0, 0), (1, 0, 0)], None, (2, 0), (3, 0), 0, False, False, None),
    "smartseq2_single_rf": ([(0, 0, 0), (1, 0, 0)], None, (2, 0), None, 2, False, False, None),
    "stormlike": ([], [(1, 0, 8)], (0, 0), (1, 14), 2, False, False, None),
    # UMI tag sequences: the UMI location is the technology's, advanced by the tag's length (src/main.cpp:1467-1468)
    "smartseq3": ([(0, 0, 0), (1, 0, 0)], [(2, 11, 19)], (2, 22), (3, 0), 1, False, False, util.SMARTSEQ3_TAG),
    "tag_single_fr": ([(0, 0, 8)], [(1, 11, 19)], (1, 22), None, 1, False, False, util.SMARTSEQ3_TAG),
}


def case_files(inputs, name):
    """-> (files: one list of sequences per file of the technology, sample ranges or None)"""
    keys = util.BUSPAIRED_CASES[name][1]
    bc, umi, seq, seq2, strand, num, per_sample, tag = SPECS[name]
    nfiles = 1 + max([seq[0]] + ([seq2[0]] if seq2 else []) + [b[0] for b in bc] + [u[0] for u in (umi or [])])
    reads = [O.read_fastq(inputs[k]) for k in keys]
    if not per_sample:
        return reads, None
    files = [[] for _ in range(nfiles)]
    samples = []
    for s in range(0, len(keys), nfiles):
        lo = len(files[0])
        for f in range(nfiles):
            files[f].extend(reads[s + f])
        samples.append((lo, len(files[0])))
    return files, samples


def sorted_records(r):
    return np.sort(r, order=["barcode", "umi", "ec", "flags", "count"])


def read_ref(name):
    d = os.path.join(D, "ref_" + name)
    raw = gzip.open(os.path.join(d, "output.bus.gz")).read()
    tmp = os.path.join(d, ".output.bus.tmp%d" % os.getpid())
    with open(tmp, "wb") as f:
        f.write(raw)
    try:
        hdr, rec = O.read_bus(tmp)
    finally:
        os.remove(tmp)
    info = json.load(open(os.path.join(d, "run_info.json")))
    ecs = O.read_matrix_ec(os.path.join(d, "matrix.ec"))
    flens = None
    if os.path.exists(os.path.join(d, "flens.txt")):
        flens = [np.array(line.split(), np.uint32) for line in open(os.path.join(d, "flens.txt"))]
    return d, hdr, rec.copy(), info, ecs, flens


@pytest.fixture(scope="module")
def inputs(tmp_path_factory):
    return util.buspaired_inputs(str(tmp_path_factory.mktemp("buspaired_in")))


@pytest.fixture(scope="module")
def oix():
    return O.OracleIndex(os.path.join(util.GOLDEN, "synth_small", "transcripts.kidx"))


@pytest.mark.parametrize("name", sorted(SPECS))
def test_bus_model_reproduces_the_reference(inputs, oix, name):
    d, hdr, ref, info, ref_ecs, ref_flens = read_ref(name)
    bc, umi, seq, seq2, strand, num, per_sample, tag = SPECS[name]
    files, samples = case_files(inputs, name)
    m = O.bus_model(oix, files, bc, umi, seq, seq2, strand=strand, num=num, samples=samples, tag=tag)
    assert m["n_processed"] == info["n_processed"]
    assert len(m["records"]) == info["n_pseudoaligned"] == len(ref)
    assert m["ecs"] == ref_ecs                                   # same sets, same ids (order of first occurrence)
    assert sorted_records(m["records"]).tobytes() == sorted_records(ref).tobytes()
    # within a sample whose ECs are all new (the first one) the reference writes in read order
    n0 = int((m["records"]["barcode"] == m["records"]["barcode"][0]).sum()) if per_sample else len(ref)
    if per_sample:
        assert m["records"][:n0].tobytes() == ref[:n0].tobytes()
    if ref_flens is not None:
        assert len(m["flens"]) == len(ref_flens)
        for a, b in zip(m["flens"], ref_flens):
            np.testing.assert_array_equal(a, b)
        assert sum(int(x.sum()) for x in ref_flens) > 100
    if per_sample:
        assert (hdr["bclen"], hdr["umilen"]) == (16, 1)
    else:                                                        # src/main.cpp:2470-2508: most frequent observed lengths
        assert hdr["bclen"] == int(np.argmax(m["bc_hist"])) and hdr["umilen"] == int(np.argmax(m["umi_hist"]))


def test_bus_model_batch_file(inputs, oix):
    """`bus --batch FILE` (src/main.cpp:1108-1180): one sample per line; lines with the same id share the fake barcode
    (batch_id_mapping, src/ProcessReads.h:211-224) but not the fragment-length histogram."""
    d, hdr, ref, info, ref_ecs, ref_flens = read_ref("batchfile")
    lines = [l for l in util.BATCHFILE_LINES if not l[0].startswith("#")]
    files, samples = [[], []], []
    for _, k1, k2 in lines:
        lo = len(files[0])
        files[0].extend(O.read_fastq(inputs[k1]))
        files[1].extend(O.read_fastq(inputs[k2]))
        samples.append((lo, len(files[0])))
    ids = []
    for name, _, _ in lines:
        if name not in ids:
            ids.append(name)
    m = O.bus_model(oix, files, [], None, (0, 0), (1, 0), samples=samples, sample_barcodes=[ids.index(l[0]) for l in lines])
    assert m["n_processed"] == info["n_processed"] and len(m["records"]) == info["n_pseudoaligned"] == len(ref)
    assert m["ecs"] == ref_ecs
    assert sorted_records(m["records"]).tobytes() == sorted_records(ref).tobytes()
    assert len(ref_flens) == 3
    for a, b in zip(m["flens"], ref_flens):
        np.testing.assert_array_equal(a, b)
    assert open(os.path.join(d, "matrix.cells")).read().split() == [l[0] for l in lines]
    assert (hdr["bclen"], hdr["umilen"]) == (16, 1)
