

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
