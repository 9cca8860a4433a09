"""CPU: the C-ABI library loads, exports every symbol include/kallisto_b200.h declares, parses
index files on the host, and fails LOUDLY (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

import kallisto_b200 as K
from tests import util


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(util.ROOT, "include", "kallisto_b200.h")).read()
    declared = set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = K.lib()
    for sym in sorted(declared):
        assert hasattr(L, sym), "missing export: " + sym
    assert declared == set(K.EXPORTED_SYMBOLS)


def test_inspect_index_matches_reference_log_lines():
    # "[index] number of targets: 14", "[index] number of k-mers: 22,118", 21 contigs
    info = K.inspect_index(util.dataset("config1")["index"])
    assert info["k"] == 31 and info["n_targets"] == 14 and info["n_kmers"] == 22118 and info["n_unitigs"] == 21
    assert info["n_ec_blocks"] == 27


def test_inspect_index_saved_targets_only():
    """index.saved of `kallisto bus` (KmerIndex::write(fn, false), src/KmerIndex.cpp:1226-1327; fixture written by the
    reference): version, empty graph / D-list / node sections, then the targets.  The host parser reads it (quant-tcc
    runs on it)."""
    info = K.inspect_index(os.path.join(util.GOLDEN, "buspaired", "ref_bulk_paired", "index.saved"))
    full = K.inspect_index(util.dataset("synth_small")["index"])
    assert info["n_targets"] == full["n_targets"] == 491 and info["n_kmers"] == 0 and info["n_unitigs"] == 0
    assert K.inspect_index(os.path.join(util.GOLDEN, "buspaired", "dlist_index.saved"))["n_targets"] == 491


def test_bad_index_is_rejected():
    with pytest.raises(K.KallistoB200Error):
        K.inspect_index(os.path.join(util.ROOT, "include", "kallisto_b200.h"))
    with pytest.raises(K.KallistoB200Error):
        K.inspect_index("/nonexistent/file.kidx")


def test_counts_to_tpm_host():
    est = np.array([10.0, 0.0, 5.0])
    eff = np.array([100.0, 50.0, 25.0])
    tpm = K.counts_to_tpm(est, eff)
    x = est / eff
    np.testing.assert_allclose(tpm, x / x.sum() * 1e6, rtol=0, atol=0)


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="only meaningful on a box without a GPU")
def test_no_device_fails_loudly():
    with pytest.raises(K.KallistoB200Error) as ei:
        K.KmerIndex(util.dataset("config1")["index"])
    assert ei.value.code == K.KB_ERR_NO_DEVICE


def test_c99_example_compiles_links_and_fails_loudly_without_a_device(tmp_path):
    """examples/minimal_quant.c is plain C99 against include/kallisto_b200.h and links against the shared library
    alone (no Python, no torch, no explicit -lz/-lcudart): the boundary is a real C ABI.  Without a CUDA device the
    first call that needs one returns an error -- there is no CPU path to fall back to."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    exe = str(tmp_path / "minimal_quant")
    libdir = os.path.join(util.ROOT, "kallisto_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(util.ROOT, "include"),
                        os.path.join(util.ROOT, "examples", "minimal_quant.c"), "-L" + libdir, "-lkallisto_b200",
                        "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, os.path.join(util.GOLDEN, "config1", "transcripts.kidx")], capture_output=True, text=True)
    if r.returncode == 0:                      # a GPU is present: the two (identical) fragments are processed
        assert "2 fragments" in r.stdout
    else:
        assert r.returncode == 1 and "no CUDA device available" in r.stderr
