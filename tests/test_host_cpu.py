"""CPU-only tests of the host side: the C-ABI library loads and exports every declared symbol, the motif
parsers / PWM maths reproduce the reference's goldens bit for bit, the TSV row parser and the report
writers keep the reference's layout, and the product fails loudly without a GPU."""
import io
import os
import re

import numpy as np
import pandas as pd
import pytest

import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def built():
    from grafimo_b200 import build
    return build.build()


def test_abi_library_exports_every_declared_symbol(built):
    import ctypes
    from grafimo_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "grafimo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gb2_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 22
    lib = ctypes.CDLL(built)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().gb2_abi_version() == 3
    assert _lib.load().gb2_error_string(2).decode().startswith("CUDA")


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(built):
    from grafimo_b200._lib import GrafimoB200Error
    from grafimo_b200.engine import Context
    with pytest.raises(GrafimoB200Error):
        Context()
    import ctypes
    from grafimo_b200 import _lib
    h = ctypes.c_void_p()
    assert _lib.load().gb2_ctx_create(0, None, ctypes.byref(h)) == 2  # GB2_ERR_CUDA


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


@pytest.mark.parametrize("tag", gu.motif_tags())
def test_motif_host_maths_bit_exact(tag, tmp_path):
    """parse -> background -> normalise -> pseudocount -> log-odds -> integer scaling (everything before the DP)."""
    from grafimo_b200 import motif_ops as mo
    g = gu.load_motif(tag)
    fx = gu.fixtures()
    src, fmt = g["source"], g["fmt"]
    ext = {"meme": "meme", "jaspar": "jaspar", "transfac": "transfac", "pfm": "pfm"}[fmt]
    key = src if src.endswith("_" + fmt) else src + "_meme"
    path = _write(tmp_path, f"motif.{ext}", fx[key])
    bg = "unfrm_dst" if g["bgfile"] == "unif" else _write(tmp_path, "bg_nt", fx["bg_nt"])
    reader = {"meme": mo._read_meme, "jaspar": mo._read_jaspar, "transfac": mo._read_transfac, "pfm": mo._read_pfm}[fmt]
    m = reader(path, bg, g["pseudo"], g["no_reverse"], False, True)
    m = m[0] if isinstance(m, list) else m
    assert m.motif_id == g["motif_id"] and m.motif_name == g["motif_name"] and m.width == g["width"]
    assert list(m.bg.keys()) == g["bg_key_order"]
    assert np.array_equal(m.bg_acgt(), g["bg_acgt"])
    assert np.array_equal(m.count_matrix, g["count_matrix"])
    mo._scale_motif(m, True)
    assert np.array_equal(m.score_matrix, g["score_matrix"])
    assert (m.min_val, m.max_val, m.scale) == (g["min_val"], g["max_val"], g["scale"])
    assert isinstance(m.offset, np.double) and m.offset == g["offset"]
    assert np.array_equal(m.score_matrix_acgt(), g["score_matrix"])


def test_reference_integer_matrix_goldens(tmp_path):
    """The reference's own expected matrices (tests/test_data/expected_results/motif_processing_test_*.txt)."""
    from grafimo_b200 import motif_ops as mo
    fx = gu.fixtures()
    exp_meme = np.loadtxt(io.StringIO(fx["expected_matrix_meme"])).astype(int)
    exp_jaspar = np.loadtxt(io.StringIO(fx["expected_matrix_jaspar"])).astype(int)
    m = mo._read_meme(_write(tmp_path, "a.meme", fx["ctcf_meme"]), "unfrm_dst", 0.1, False, False, True)[0]
    assert (mo._scale_motif(m, True).score_matrix == exp_meme).all()
    for fmt, reader in (("jaspar", mo._read_jaspar), ("transfac", mo._read_transfac), ("pfm", mo._read_pfm)):
        m = reader(_write(tmp_path, f"a.{fmt}", fx[f"ctcf_{fmt}"]), "unfrm_dst", 0.1, False, False, True)
        assert (mo._scale_motif(m, True).score_matrix == exp_jaspar).all(), fmt
    assert (exp_meme != exp_jaspar).sum() == 1  # 52 vs 53 in one cell


def test_format_sniffers(tmp_path):
    from grafimo_b200 import utils
    fx = gu.fixtures()
    paths = {fmt: _write(tmp_path, f"MA0139.1.{fmt}", fx[f"ctcf_{fmt}"]) for fmt in ("meme", "jaspar", "transfac", "pfm")}
    assert utils.is_meme(paths["meme"]) and not utils.is_jaspar(paths["meme"])
    assert utils.is_jaspar(paths["jaspar"]) and not utils.is_meme(paths["jaspar"])
    assert utils.is_transfac(paths["transfac"]) and not utils.is_meme(paths["transfac"])
    assert utils.is_pfm(paths["pfm"]) and not utils.is_transfac(paths["pfm"])


def test_kmer_table_parser_matches_reference_field_rules(tmp_path):
    from grafimo_b200.score_sequences import KmerTable
    from oracle import oracle as orc
    c = gu.load_scoring("fixture_plus_N_2files")
    files = []
    for k, lines in enumerate(c["files"]):
        files.append(_write(tmp_path, f"r{k}.tsv", "\n".join(lines) + "\n"))
    for norev in (False, True):
        t = KmerTable.read(files, norev)
        r = orc.parse_rows([ln for f in c["files"] for ln in f], norev)
        assert list(t.seq) == r["seq"] and list(t.seqname) == r["seqname"] and list(t.strand) == r["strand"]
        assert np.array_equal(t.start, r["start"]) and np.array_equal(t.stop, r["stop"]) and np.array_equal(t.freq, r["freq"])
        assert list(t.ref) == r["ref"]
    # the reference's own k-mer fixture (expected_seqs.tsv: no GBWT -> freq 0, single-digit node ids)
    fx = gu.fixtures()
    t = KmerTable.read([_write(tmp_path, "x.tsv", fx["expected_seqs_tsv"])], False)
    assert len(t) == 32 and set(t.strand) == {"+", "-"} and set(t.freq.tolist()) == {0}


@pytest.mark.parametrize("tag", ["fixture_testmode", "fixture_noq", "synth_w8"])
def test_writers_keep_the_reference_layout(tag, tmp_path):
    from grafimo_b200.res_writer import writeGFF3
    c = gu.load_scoring(tag)
    df = pd.DataFrame({col: c["table"][col] for col in c["columns"]}).head(25)
    noq = c["options"]["noqvalue"]
    prefix = str(tmp_path / "out")
    writeGFF3(prefix, df, noq, True)
    assert open(prefix + ".gff").read() == c["gff3_head25"]
    buf = io.StringIO()
    df.to_csv(buf, sep="\t", encoding="utf-8")
    assert buf.getvalue() == c["tsv_head25"]


def test_gff3_known_row():
    """Verbatim reference output for golden row 0."""
    from grafimo_b200.res_writer import gff3_lines
    df = pd.DataFrame({
        "motif_id": ["MA0139.1"], "motif_alt_id": ["CTCF"], "sequence_name": ["22:19723256-19723526"],
        "start": [19723401], "stop": [19723382], "strand": ["-"], "score": [1.2096774193548185],
        "p-value": [0.0013399618583207484], "q-value": [0.49884025391656905], "matched_sequence": ["CTATCGCCGGAGGCCGCAG"],
        "haplotype_frequency": [5096], "reference": ["ref"]})
    assert list(gff3_lines(df, False)) == [
        "22\tgrafimo\tnucleotide_motif\t19723382\t19723401\t1.2\t-\t.\tName=MA0139.1_22:19723256-19723526-:ref;"
        "Alias=CTCF;ID=MA0139.1=-=CTCF=-=22:19723256-19723526;pvalue==1.3399618583207484e-03;"
        "qvalue=4.9884025391656905e-01;sequence==CTATCGCCGGAGGCCGCAG=;\n"]


def test_findmotif_container_and_cli_parser():
    from grafimo_b200.__main__ import get_parser
    from grafimo_b200.workflow import Findmotif
    a = get_parser().parse_args(["findmotif", "-m", "x.meme", "--kmers-dir", "d", "-t", "0.01", "--recomb", "-r"])
    wf = Findmotif(motif=a.motif, kmers_dir=a.kmers_dir, threshold=a.threshold, recomb=a.recomb, no_reverse=a.no_reverse)
    assert (wf.threshold, wf.recomb, wf.noreverse, wf.noqvalue, wf.qvalueT, wf.bgfile, wf.pseudo) == \
        (0.01, True, True, False, False, "unfrm_dst", 0.1)
    with pytest.raises(ValueError):
        Findmotif(qval_t=True, no_qvalue=True)
    with pytest.raises(TypeError):
        Findmotif(threshold=1)


def test_text_chunks_cut_at_line_boundaries(tmp_path):
    from grafimo_b200.score_sequences import _text_chunks
    rng = np.random.default_rng(0)
    files, want = [], b""
    for k in range(4):
        lines = [("x" * int(rng.integers(1, 90))).encode() for _ in range(int(rng.integers(1, 60)))]
        body = b"\n".join(lines) + (b"\n" if k % 2 else b"")  # every other file lacks the final newline
        p = tmp_path / f"f{k}.tsv"
        p.write_bytes(body)
        files.append(str(p))
        want += body if body.endswith(b"\n") else body + b"\n"
    for chunk_bytes, reuse in ((128, False), (128, True), (1000, True), (1000, False), (1 << 20, False), (1 << 20, True)):
        segments = []
        chunks = [bytes(c.numpy()) for c in _text_chunks(files, chunk_bytes, segments, reuse=reuse)]
        assert len(segments) == len(chunks)
        # segments name the file every line of a chunk comes from
        whole = [p.read_bytes() for p in map(type(tmp_path), files)]
        for c, segs in zip(chunks, segments):
            assert segs and segs[0][1] == 0 and [o for _, o in segs] == sorted(o for _, o in segs)
            for k, (fi, off) in enumerate(segs):
                end = segs[k + 1][1] if k + 1 < len(segs) else len(c)
                for ln in c[off:end].split(b"\n"):
                    assert ln == b"" or ln in whole[fi].split(b"\n")
        assert all(c.endswith(b"\n") and len(c) <= chunk_bytes for c in chunks)
        # blank lines may be added between files (the device line index skips them); the lines themselves are intact
        assert [ln for ln in b"".join(chunks).split(b"\n") if ln] == [ln for ln in want.split(b"\n") if ln]
        for c in chunks[:-1]:
            assert c.endswith(b"\n")


def test_top_level_motif_processing_module_and_duck_typed_motifs():
    """The reference imports its Cython extension as the top-level module `motif_processing` (setup.py:53,
    motif_ops.py:29-35): the same name resolves to this package's implementation, and the seams accept any object with
    the reference Motif's properties (not only this package's class)."""
    import importlib
    mp = importlib.import_module("motif_processing")
    for name in ("read_bg_file", "get_uniform_bg", "apply_pseudocount_jaspar_transfac_pfm", "apply_pseudocount_meme",
                 "compute_log_odds", "comp_pval_mat"):
        assert callable(getattr(mp, name))
    from grafimo_b200.motif import bg_acgt, is_motif, score_matrix_acgt

    class RefLikeMotif:  # the property set of src/grafimo/motif.py, rows in the file's own order (here T,G,C,A)
        score_matrix = np.arange(12).reshape(4, 3)
        pval_matrix = np.ones(3001)
        min_val, scale, width, offset, is_scaled = 0, 10, 3, np.float64(-2.0), True
        motif_id, motif_name = "M1", "m"
        bg = {"A": 0.1, "T": 0.4, "C": 0.2, "G": 0.3}
        nucsmap = {"T": 0, "G": 1, "C": 2, "A": 3}

    m = RefLikeMotif()
    assert is_motif(m) and not is_motif(object())
    assert score_matrix_acgt(m).tolist() == [[9, 10, 11], [6, 7, 8], [3, 4, 5], [0, 1, 2]]
    assert bg_acgt(m).tolist() == [0.1, 0.2, 0.3, 0.4]


def test_string_column_helpers_on_host_text():
    """_gather_rows / _fixed_strings / _var_strings (the string columns of reported rows) on a numpy text buffer: the
    same code gathers from the device copy of the text in compute_results."""
    from grafimo_b200.score_sequences import _fixed_strings, _gather_rows, _var_strings
    lines = [b"chr7:100-900\tACGTACG\tx", b"chr7:100-900\tTTTTGGG\ty", b"12:5-6\tCCCCAAA\tz", b"chr7:100-900\tGGGGGGG\tw"]
    text = np.frombuffer(b"\n".join(lines) + b"\n", dtype=np.uint8)
    offs = np.cumsum([0] + [len(ln) + 1 for ln in lines[:-1]]).astype(np.int64)
    name_len = np.array([ln.index(b"\t") for ln in lines], dtype=np.int64)
    seq_off = name_len + 1
    order = np.array([2, 0, 3, 1])  # rows come back in report order, not file order
    assert _fixed_strings(text, (offs + seq_off)[order], 7).tolist() == ["CCCCAAA", "ACGTACG", "GGGGGGG", "TTTTGGG"]
    assert _var_strings(text, offs[order], name_len[order]).tolist() == ["12:5-6", "chr7:100-900", "chr7:100-900", "chr7:100-900"]
    g = _gather_rows(text, np.array([len(text) - 3], dtype=np.int64), 8)  # indices beyond the text are clamped
    assert g.shape == (1, 8) and bytes(g[0, :3]) == b"\tw\n" and set(g[0, 3:].tolist()) == {10}
    assert _fixed_strings(text, np.zeros(0, np.int64), 7).tolist() == [] and _var_strings(text, np.zeros(0, np.int64), np.zeros(0, np.int64)).tolist() == []


@pytest.mark.parametrize("isa", ["", "avx2", "scalar"])
def test_host_packer_equals_numpy_packer(isa, monkeypatch):
    """gb2_pack_sequence_host (csrc/host_pack.cpp: the transfer compression of gb2_scan_host_sequences and a public utility;
    AVX-512 / AVX2 / scalar by what the CPU has) against the plain numpy packer: words, N bits and counters, for lengths around
    the 32- and 64-base steps, lower case, N and other symbols.  No GPU involved."""
    from grafimo_b200 import engine
    if isa:
        monkeypatch.setenv("GB2_HOST_PACK_ISA", isa)
    else:
        monkeypatch.delenv("GB2_HOST_PACK_ISA", raising=False)
    rng = np.random.default_rng(77)
    letters = np.array(list("ACGTacgt"))
    seqs = []
    for n in [0, 1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1000, 4096, 70001]:
        s = rng.choice(letters, size=n)
        if n:
            s = np.where(rng.random(n) < 0.01, "N", s)
            s = np.where(rng.random(n) < 0.003, "n", s)
            s = np.where(rng.random(n) < 0.002, rng.choice(np.array(list("RYKMxz-*"))), s)
        seqs.append("".join(s))
    words, nbits, off, lens, counts = engine.pack_sequences_host(seqs)
    e_words, e_nbits, e_off, e_lens = engine.pack_sequences_2bit(seqs)
    assert np.array_equal(off, e_off) and np.array_equal(lens, e_lens)
    assert np.array_equal(words, e_words) and np.array_equal(nbits, e_nbits)
    flat = "".join(seqs)
    assert int(counts[0]) == sum(ch not in "ACGTacgt" for ch in flat) > 0
    assert int(counts[1]) == sum(ch not in "ACGTacgtNn" for ch in flat) > 0


def test_header_is_plain_c_and_the_example_compiles(tmp_path):
    """include/grafimo_b200.h must be consumable by a C compiler on its own (the boundary is a C ABI: plain pointers and sizes, no
    C++ / CUDA / torch types), and examples/c_abi_scan.c must compile and link against the built library (it is RUN on a GPU box by
    tests/test_gpu_c_abi.py)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc here")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tu = tmp_path / "only_the_header.c"
    tu.write_text('#include "grafimo_b200.h"\nint main(void) { return GB2_ABI_VERSION == 3 ? 0 : 1; }\n')
    for std in ("-std=c99", "-std=c11"):
        subprocess.run(["gcc", std, "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(root, "include"), str(tu)], check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", "-I", os.path.join(root, "include"), str(tu)], check=True)
    lib = os.path.join(root, "grafimo_b200")
    if os.path.exists(os.path.join(lib, "libgrafimo_b200.so")):
        subprocess.run(["gcc", "-O2", "-Wall", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "c_abi_scan.c"), "-o",
                        str(tmp_path / "c_abi_scan"), "-L", lib, "-lgrafimo_b200", f"-Wl,-rpath,{lib}", "-lm"], check=True)


@pytest.mark.parametrize("n", [0, 1, 5000])
def test_arrow_backed_table_equals_generic_table(n):
    """score_sequences._build_table_rows (string columns assembled from byte buffers where pandas backs them by Arrow, report
    order optionally pre-computed) == the generic _build_table on the same rows: values, order, dtypes -- with ties on
    (p, start, stop, strand) that only the sequence resolves, the frequency filter and an empty input."""
    from grafimo_b200 import score_sequences as ss
    rng = np.random.default_rng(5 + n)
    w = 11

    class M:
        motif_id, motif_name = "MA0000.1", "TEST"

    pval = np.round(rng.random(n) * 1e-3, 5)  # many ties
    start = rng.integers(0, 40, n).astype(np.int64)
    stop = start + w
    minus = rng.random(n) < 0.5
    asc = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=(n, w))
    names = [f"chr{c}:0-1000" for c in range(3)]
    region = rng.integers(0, 3, n).astype(np.int64)
    freq = rng.integers(0, 6, n).astype(np.int64)
    isref = rng.random(n) < 0.5
    score, q = rng.random(n), rng.random(n)
    for keep in (np.ones(n, dtype=bool), freq > 0):
        seq = np.ascontiguousarray(asc).view(f"S{w}").ravel().astype(f"U{w}").astype(object) if n else np.array([], dtype=object)
        generic = ss._build_table(M, False, keep, np.array(names, dtype=object)[region] if n else np.array([], dtype=object), start, stop,
                                  np.where(minus, "-", "+").astype(object), score, pval, q, seq, freq,
                                  np.where(isref, "ref", "non.ref").astype(object), 1)
        fast = ss._build_table_rows(M, False, keep, names, region, start, stop, minus, score, pval, q, asc, freq, isref)
        assert list(fast.columns) == list(generic.columns) and fast.dtypes.equals(generic.dtypes)
        assert fast.equals(generic)
        if n > 1:  # rows handed over in (p, start, stop, strand) order, as the device leaves them
            o = np.lexsort((minus.astype(np.int8), stop, start, pval))
            pre = ss._build_table_rows(M, False, keep[o], names, region[o], start[o], stop[o], minus[o], score[o], pval[o], q[o], asc[o],
                                       freq[o], isref[o], presorted=True)
            assert pre.equals(generic)
        noq = ss._build_table_rows(M, True, keep, names, region, start, stop, minus, score, pval, None, asc, freq, isref)
        assert "q-value" not in noq.columns and len(noq) == len(generic)
