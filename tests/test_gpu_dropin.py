"""GPU tests of the drop-in seams (B1 compute_results, B2 comp_pval_mat / build_motif_*, B3 compute_qvalues)
against the tables the unmodified reference produced (tests/golden/cases) and its own golden file."""
import os

import numpy as np
import pandas as pd
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def _build(tag, tmp_path):
    from grafimo_b200 import motif_ops as mo
    g = gu.load_motif(tag)
    fx = gu.fixtures()
    fmt = g["fmt"]
    key = g["source"] if g["source"].endswith("_" + fmt) else g["source"] + "_meme"
    path = _write(tmp_path, f"motif_{tag}.{fmt}", fx[key])
    bg = "unfrm_dst" if g["bgfile"] == "unif" else _write(tmp_path, "bg_nt", fx["bg_nt"])
    if fmt == "meme":
        m = mo.build_motif_meme(path, bg, g["pseudo"], g["no_reverse"], 1, False, True)[0]
    else:
        fn = {"jaspar": mo.build_motif_jaspar, "transfac": mo.build_motif_transfac, "pfm": mo.build_motif_pfm}[fmt]
        m = fn(path, bg, g["pseudo"], g["no_reverse"], False, True)
    return m, g


@pytest.mark.parametrize("tag", gu.motif_tags())
def test_build_motif_end_to_end(tag, tmp_path):
    m, g = _build(tag, tmp_path)
    assert np.array_equal(m.score_matrix, g["score_matrix"])
    assert np.array_equal(m.pval_matrix, g["pval_mat"])  # K3, bit-exact
    assert m.is_scaled and m.scale == g["scale"] and m.offset == g["offset"]


def test_comp_pval_mat_seam(tmp_path):
    from grafimo_b200.motif_processing import comp_pval_mat
    from grafimo_b200.grafimo_errors import MotifProcessingError
    from grafimo_b200 import motif_ops as mo
    m, g = _build("ctcf_meme__bgnt", tmp_path)
    assert np.array_equal(comp_pval_mat(m, True), g["pval_mat"])
    fx = gu.fixtures()
    raw = mo._read_meme(_write(tmp_path, "u.meme", fx["ctcf_meme"]), "unfrm_dst", 0.1, False, False, True)[0]
    raw.set_motif_score_matrix(np.ones((4, 19)))
    with pytest.raises(MotifProcessingError):
        comp_pval_mat(raw, True)  # not scaled yet (motif_processing.pyx:574-576)


class _Args:
    def __init__(self, o):
        self.cores, self.threshold, self.noqvalue, self.qvalueT = 1, float(o["threshold"]), o["noqvalue"], o["qvalueT"]
        self.noreverse, self.recomb, self.verbose = o["noreverse"], o["recomb"], False


@pytest.mark.parametrize("tag", gu.scoring_tags())
def test_compute_results_equals_reference_table(tag, tmp_path, capsys):
    from grafimo_b200.score_sequences import compute_results
    c = gu.load_scoring(tag)
    m, g = _build(c["motif_tag"], tmp_path)
    d = tmp_path / "seqs" / f"width_{m.width}"
    d.mkdir(parents=True)
    for k, lines in enumerate(c["files"]):
        (d / f"region_{k}.tsv").write_text("\n".join(lines) + "\n")
    testmode = tag == "fixture_testmode"
    cwd = os.getcwd()
    df = compute_results(m, str(tmp_path / "seqs") + "/", True, None if testmode else _Args(c["options"]), testmode=testmode)
    assert os.getcwd() == cwd
    assert list(df.columns) == c["columns"]
    got = {col: df[col].to_numpy() for col in df.columns}
    gu.assert_tables_equal(got, c["table"], c["columns"])
    assert df["p-value"].is_monotonic_increasing
    assert df["start"].dtype == np.int64 and df["haplotype_frequency"].dtype == np.int64 and df["score"].dtype == np.float64
    out = capsys.readouterr().out
    n = sum(1 for f in c["files"] for ln in f if not (c["options"]["noreverse"] and ln.split()[2][-1] == "-"))
    assert f"Scanned sequences:\t{n}" in out and f"Scanned nucleotides:\t{n * m.width}" in out
    assert f"Scoring hits for motif +{m.motif_id}." in out



@pytest.mark.parametrize("threshold", [0.001, 0.005, 0.006, 0.02, 0.2])
@pytest.mark.parametrize("qval_t", [False, True])
def test_compute_results_across_the_dense_switch(threshold, qval_t, tmp_path, capsys):
    """compute_results on the reference's 704-row fixture for p-value thresholds on both sides of the point where the scan
    switches from hit records to dense scores + partition passes (score_sequences._DENSE_FROM = 0.006), with and without
    --qvalueT: every table == the oracle's (the oracle is pinned on the reference's own golden at threshold 1)."""
    from grafimo_b200.score_sequences import compute_results
    from oracle import oracle as orc
    fx = gu.fixtures()
    m, g = _build("ctcf_meme__unif", tmp_path)
    d = tmp_path / "input" / "width_19"
    d.mkdir(parents=True)
    (d / "scoring_test_input.tsv").write_text(fx["scoring_input_tsv"])
    opts = dict(threshold=threshold, noqvalue=False, qvalueT=qval_t, noreverse=False, recomb=True)
    if qval_t:
        opts["threshold"] = max(threshold, 0.5)  # q-values of this fixture start at 0.47
    df = compute_results(m, str(tmp_path / "input") + "/", True, _Args(opts))
    lines = fx["scoring_input_tsv"].strip("\n").split("\n")
    exp = orc.compute_results(g, lines, **opts)
    cols = ["sequence_name", "start", "stop", "strand", "score", "p-value", "q-value", "matched_sequence", "haplotype_frequency", "reference"]
    got = {c: df[c].to_numpy() for c in df.columns}
    assert len(df) == len(exp["start"]) and (len(df) > 0 or threshold < 0.002)
    gu.assert_tables_equal(got, exp, cols)


def test_compute_results_reference_own_golden(tmp_path):
    """The reference's test_scoring: CTCF on width_19/scoring_test_input.tsv == expected_results/scoring_results.tsv."""
    from grafimo_b200.score_sequences import compute_results
    fx = gu.fixtures()
    m, _ = _build("ctcf_meme__unif", tmp_path)
    d = tmp_path / "input" / "width_19"
    d.mkdir(parents=True)
    (d / "scoring_test_input.tsv").write_text(fx["scoring_input_tsv"])
    results = compute_results(m, str(tmp_path / "input") + "/", True, None, testmode=True)
    results.to_csv(tmp_path / "scoring_test.tsv", sep="\t")
    key = ["p-value", "start", "stop"]
    got = pd.read_csv(tmp_path / "scoring_test.tsv", sep="\t", index_col=0).sort_values(key).reset_index(drop=True)
    (tmp_path / "expected.tsv").write_text(fx["scoring_results_tsv"])
    exp = pd.read_csv(tmp_path / "expected.tsv", sep="\t", index_col=0).sort_values(key).reset_index(drop=True)
    assert got.equals(exp)  # the reference's own assertion (tests/grafimo_run_test.py:127-137)


def test_seams_accept_reference_style_motif_objects(tmp_path):
    """compute_results / comp_pval_mat with an object that only has the reference Motif's properties (src/grafimo/motif.py;
    not an instance of this package's class, no extra methods, does not take new attributes) -- what the unmodified
    reference hands over when `motif_processing` / `score_sequences` resolve to this repository."""
    import motif_processing as top  # the top-level name the reference imports (setup.py:53)
    from grafimo_b200.score_sequences import compute_results
    fx = gu.fixtures()
    m, _ = _build("ctcf_meme__bgnt", tmp_path)

    class RefLikeMotif:
        __slots__ = ("_m",)

        def __init__(self, m):
            self._m = m
        score_matrix = property(lambda s: s._m.score_matrix)
        pval_matrix = property(lambda s: s._m.pval_matrix)
        min_val = property(lambda s: s._m.min_val)
        scale = property(lambda s: s._m.scale)
        width = property(lambda s: s._m.width)
        offset = property(lambda s: s._m.offset)
        is_scaled = property(lambda s: s._m.is_scaled)
        motif_id = property(lambda s: s._m.motif_id)
        motif_name = property(lambda s: s._m.motif_name)
        bg = property(lambda s: s._m.bg)
        nucsmap = property(lambda s: s._m.nucsmap)
        alphabet = property(lambda s: s._m.alphabet)

    r = RefLikeMotif(m)
    assert np.array_equal(top.comp_pval_mat(r, True), m.pval_matrix)
    d = tmp_path / "input" / "width_19"
    d.mkdir(parents=True)
    (d / "scoring_test_input.tsv").write_text(fx["scoring_input_tsv"])
    a = compute_results(r, str(tmp_path / "input") + "/", True, None, testmode=True)
    b = compute_results(m, str(tmp_path / "input") + "/", True, None, testmode=True)
    assert a.equals(b) and len(a) == 704


def test_parsed_rows_are_reused_across_motifs_and_flags(tmp_path):
    """compute_results is called once per motif on the same files (grafimo.py:177-179): the parsed device rows of the
    last file set are reused (no second read / parse), a changed file is read again, and the tables stay the goldens."""
    from grafimo_b200 import score_sequences as ss
    from grafimo_b200.engine import Context
    c1, c2 = gu.load_scoring("fixture_default"), gu.load_scoring("fixture_t05_norecomb")  # same motif, same input rows
    m, _ = _build(c1["motif_tag"], tmp_path)
    d = tmp_path / "seqs" / f"width_{m.width}"
    d.mkdir(parents=True)
    (d / "region_0.tsv").write_text("\n".join(c1["files"][0]) + "\n")
    calls = []
    orig = Context.parse_kmer_tsv

    def counting(self, text, width, skip_minus=False):
        calls.append(int(text.shape[0]))
        return orig(self, text, width, skip_minus)
    Context.parse_kmer_tsv = counting
    try:
        ss.clear_parsed_cache()
        for c in (c1, c2, c1):
            df = ss.compute_results(m, str(tmp_path / "seqs"), True, _Args(c["options"]))
            gu.assert_tables_equal({col: df[col].to_numpy() for col in df.columns}, c["table"], c["columns"])
        assert len(calls) == 1
        lines = c1["files"][0]
        (d / "region_0.tsv").write_text("\n".join(lines[: len(lines) // 2]) + "\n")  # the file changes: parsed again
        df = ss.compute_results(m, str(tmp_path / "seqs"), True, _Args(c1["options"]))
        assert len(calls) == 2 and len(df) <= len(c1["table"]["start"])
        ss.clear_parsed_cache()
        ss.compute_results(m, str(tmp_path / "seqs"), True, _Args(c1["options"]))
        assert len(calls) == 3
    finally:
        Context.parse_kmer_tsv = orig
        ss.clear_parsed_cache()


def test_compute_results_errors(tmp_path):
    from grafimo_b200.score_sequences import compute_results
    m, _ = _build("ctcf_meme__unif", tmp_path)
    with pytest.raises(FileNotFoundError):
        compute_results(m, str(tmp_path / "nope"), True, None, testmode=True)
    (tmp_path / "empty" / "width_19").mkdir(parents=True)
    with pytest.raises(ValueError):  # zero rows (score_sequences.py:189-192)
        compute_results(m, str(tmp_path / "empty"), True, None, testmode=True)
    with pytest.raises(TypeError):
        compute_results("not a motif", str(tmp_path), True, None, testmode=True)
    with pytest.raises(SystemExit) as e:  # debug=False: message + exit code 1 (utils.py:63-78)
        compute_results(m, str(tmp_path / "nope"), False, None, testmode=True)
    assert e.value.code == 1
    d = tmp_path / "bad" / "width_19"
    d.mkdir(parents=True)
    (d / "r.tsv").write_text("1:1-30\tACGT\t1:1+\t1:5+\t1\tref\t1+,\n")
    with pytest.raises(ValueError):
        compute_results(m, str(tmp_path / "bad"), True, None, testmode=True)


def test_compute_qvalues_seam():
    from grafimo_b200.score_sequences import compute_qvalues
    from oracle import oracle as orc
    rng = np.random.default_rng(3)
    p = np.round(rng.random(20000), 4)
    p[::11] = 1.0
    q = compute_qvalues(p.tolist(), True)
    assert isinstance(q, list) and np.array_equal(np.array(q), orc.bh(p))


def test_cli_findmotif_writes_reports(tmp_path):
    from grafimo_b200.__main__ import main
    fx = gu.fixtures()
    meme = _write(tmp_path, "MA0139.1.meme", fx["ctcf_meme"])
    d = tmp_path / "kmers" / "width_19"
    d.mkdir(parents=True)
    (d / "a.tsv").write_text(fx["scoring_input_tsv"])
    out = tmp_path / "out"
    rc = main(["findmotif", "-m", meme, "--kmers-dir", str(tmp_path / "kmers"), "-t", "0.05", "--recomb", "-o", str(out),
               "--debug"])
    assert rc == 0
    tsv = pd.read_csv(out / "grafimo_out.tsv", sep="\t", index_col=0)
    c = gu.load_scoring("fixture_t05_norecomb")
    assert len(tsv) >= len(c["table"]["start"]) and (tsv["p-value"] < 0.05).all()
    gff = (out / "grafimo_out.gff").read_text().split("\n")
    assert gff[0] == "##gff-version 3" and len(gff) == len(tsv) + 2
    assert (out / "grafimo_out.html").exists()


def test_compute_results_multi_chunk_text(tmp_path, monkeypatch):
    """The TSV text is parsed in chunks; force tiny chunks and expect the same table."""
    from grafimo_b200 import score_sequences as ss
    c = gu.load_scoring("fixture_plus_N_2files")
    m, g = _build(c["motif_tag"], tmp_path)
    d = tmp_path / "seqs" / "width_19"
    d.mkdir(parents=True)
    for k, lines in enumerate(c["files"]):
        (d / f"region_{k}.tsv").write_text("\n".join(lines) + ("\n" if k else ""))
    monkeypatch.setattr(ss, "_CHUNK_BYTES", 8192)
    chunks = list(ss._text_chunks(sorted(str(p) for p in d.iterdir()), 8192))
    assert len(chunks) > 5
    df = ss.compute_results(m, str(tmp_path / "seqs"), True, _Args(c["options"]))
    got = {col: df[col].to_numpy() for col in df.columns}
    gu.assert_tables_equal(got, c["table"], c["columns"])


def test_compute_results_over_two_gpus(tmp_path):
    """Needs two visible GPUs: torchrun with 2 ranks, files split over the ranks, global q-values via NCCL."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dist_compute_results_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29577", worker, str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
