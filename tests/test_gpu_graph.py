"""GPU tests of K7 (csrc/graph.cu): k-mer extraction from the variation graph against oracle/graph_oracle.py and the
reference's vg fixture, and the text-free path graph -> K7 -> K2/K5/K6 against the TSV path on the same rows."""
import numpy as np
import pytest

import golden_util as gu
import graph_util as gr
from oracle import graph_oracle as go

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from grafimo_b200.engine import Context
    return Context(0)


def _rows_as_tuples(rows, width):
    from grafimo_b200.extract_regions import decode_kmers
    h = rows.host()
    asc = decode_kmers(h["packed"], width)
    n = rows.n
    bad = (h["nmask"].view(np.uint32)[np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1 if n else np.zeros(0)
    out = []
    for i in range(n):
        seq = asc[i].tobytes().decode()
        nodes = tuple(int(x) + 1 for x in h["walk"][i, :h["walk_len"][i]])
        out.append((int(h["region"][i]), int(h["start"][i]), int(h["stop"][i]), seq, int(h["freq"][i]), bool(h["isref"][i]),
                    nodes, bool(bad[i])))
    return out


def test_extract_equals_reference_vg_fixture(ctx):
    """test.fa + test.vcf.gz, region x:0-20, K=19 -> the 32 lines of expected_seqs.tsv, node paths included."""
    from grafimo_b200.vgraph import VariationGraph
    fx = gu.fixtures()
    ref = "".join(fx["test_fa"].split("\n")[1:])
    variants, _ = go.parse_vcf_text(fx["test_vcf"], "x")
    dg = VariationGraph.build("x", ref, variants, None).to_device(ctx)
    rows = dg.extract([(0, 20)], 19, want_walks=True)
    lines = rows.to_vg_tsv()[0]
    exp = [ln for ln in fx["expected_seqs_tsv"].split("\n") if ln]
    assert sorted(lines) == sorted(exp)


@pytest.mark.parametrize("builder", ["numpy", "native"])
def test_extract_equals_real_vg_kmer_fixture(ctx, builder):
    """K7 pinned on REAL vg output: the local graph of the reference's 704-row fixture (reconstructed from its own rows,
    tests/fixture_graph.py) -> exactly those 704 lines: sequence, start, stop, strand, haplotype frequency (0 included),
    ref flag, node path -- the 36 rows through the 2-bp deletion included."""
    import fixture_graph as fg
    from grafimo_b200.extract_regions import DeviceGraph
    from grafimo_b200.vgraph import VariationGraph
    c = fg.reconstruct(gu.fixtures()["scoring_input_tsv"])
    if builder == "numpy":
        dg = VariationGraph.build(c["chrom"], c["ref"], c["variants"], c["gt"]).to_device(ctx)
    else:
        dg = DeviceGraph.build(ctx, c["chrom"], c["ref"], c["variants"], gt=c["gt"])
        dg.graph = VariationGraph.build(c["chrom"], c["ref"], c["variants"], c["gt"])
    rows = dg.extract([c["local_region"]], c["w"], want_walks=True)
    lines = fg.shift_lines(rows.to_vg_tsv()[0], c)
    exp = [r["line"] for r in c["rows"]]
    assert rows.n == 352 and sorted(lines) == sorted(exp)


def test_real_vg_fixture_graph_to_reference_scoring_table(ctx, tmp_path):
    """The whole replacement path on the reference's own data: reconstructed graph -> K7 -> K2/K5/K6 -> report table ==
    tests/test_data/expected_results/scoring_results.tsv (what the unmodified reference made of vg's rows; its test_scoring
    golden), after shifting the local coordinates back."""
    import io
    import pandas as pd
    import fixture_graph as fg
    from grafimo_b200 import score_sequences as ss
    from grafimo_b200.extract_regions import DeviceGraph
    ss._ctx = ctx
    fx = gu.fixtures()
    c = fg.reconstruct(fx["scoring_input_tsv"])
    motif = _motif(tmp_path)
    dg = DeviceGraph.build(ctx, c["chrom"], c["ref"], c["variants"], gt=c["gt"])
    rows = dg.extract([c["local_region"]], c["w"])
    df = ss.compute_results_rows(motif, rows, True, _Args(threshold=1.0, recomb=True))
    exp = pd.read_csv(io.StringIO(fx["scoring_results_tsv"]), sep="\t", index_col=0, float_precision="round_trip")
    assert len(df) == len(exp) == 704
    got = {k: df[k].to_numpy() for k in df.columns}
    got["start"] = got["start"] + c["offset"]
    got["stop"] = got["stop"] + c["offset"]
    cols = ["start", "stop", "strand", "score", "p-value", "q-value", "matched_sequence", "haplotype_frequency", "reference"]
    gu.assert_tables_equal(got, {k: exp[k].to_numpy() for k in exp.columns}, cols)


@pytest.mark.parametrize("builder", ["numpy", "native"])
@pytest.mark.parametrize("seed", range(6))
def test_extract_equals_oracle_random_graphs(ctx, seed, builder):
    """Both graph builders -- vgraph.VariationGraph.build (numpy, host arrays -> gb2_graph_create) and the library's
    gb2_graph_build -- against the oracle: same rows, coordinates, node ids, ref flags, frequencies."""
    from grafimo_b200.extract_regions import DeviceGraph
    from grafimo_b200.vgraph import VariationGraph
    ref, vs, gt = gr.random_case(200 + seed, length=600, n_var=60, n_hap=70 if seed % 2 else 12,
                                 n_frac=0.01 if seed % 3 == 0 else 0.0)
    m = (8, 32, 64)[seed % 3]  # 64: nodes longer than one packed word take the byte path of the kernel
    if builder == "numpy":
        dg = VariationGraph.build("c", ref, vs, gt, max_node_len=m).to_device(ctx)
    else:
        dg = DeviceGraph.build(ctx, "c", ref, vs, gt=gt, max_node_len=m)
        dg.graph = VariationGraph.build("c", ref, vs, gt, max_node_len=m)  # host arrays, only to spell N rows as text
        assert (dg.info.n_nodes, dg.info.n_edges, dg.info.n_sets) == (dg.graph.n_nodes, dg.graph.n_edges, dg.graph.n_cons)
    regions = [(0, 600), (37, 301), (100, 250), (590, 600), (250, 250)]
    for w in (5, 19, 32, 33, 40) + ((64,) if seed == 1 else ()):  # above 32: two packed words per row, 64-deep stacks
        rows = dg.extract(regions, w, want_walks=True)
        got = _rows_as_tuples(rows, w)
        exp = []
        for r, region in enumerate(regions):
            for (start, stop, seq, freq, isref, nodes) in gr.oracle_rows(ref, vs, gt, region, w, max_node=m):
                exp.append((r, start, stop, seq.replace("N", "A"), freq, isref, nodes, "N" in seq))
        assert sorted(got) == sorted(exp)
        assert [g[0] for g in got] == sorted(g[0] for g in got)  # rows are grouped by region, in order
        assert rows.n_masked() == sum(e[7] for e in exp)
        # the text form spells the N rows from their walks
        if w == 19:
            text = sorted(ln for lines in rows.to_vg_tsv().values() for ln in lines)
            og = go.build_graph(ref, vs, max_node=m)
            want = sorted(ln for region in regions if region[0] < region[1] for ln in
                          go.vg_tsv_lines(go.extract_rows(og, gt.tolist(), region, w), "c", region))
            assert text == want


def test_batch_builder_equals_single_builds(ctx):
    """gb2_graph_build_batch (host passes on worker threads) == gb2_graph_build per chromosome: same graphs, same rows."""
    from grafimo_b200._lib import GrafimoB200Error
    from grafimo_b200.extract_regions import DeviceGraph
    cases = [gr.random_case(900 + k, length=400 + 300 * k, n_var=30 + 25 * k, n_hap=(0, 12, 70, 40, 33)[k]) for k in range(5)]
    items = [(f"c{k}", ref, vs, gt if k else None, None) for k, (ref, vs, gt) in enumerate(cases)]
    for threads in (0, 1, 3):
        many = DeviceGraph.build_many(ctx, items, n_threads=threads)
        assert [g.chrom for g in many] == [f"c{k}" for k in range(5)]
        for k, (ref, vs, gt) in enumerate(cases):
            one = DeviceGraph.build(ctx, f"c{k}", ref, vs, gt=gt if k else None)
            assert (many[k].info.n_nodes, many[k].info.n_edges, many[k].info.n_sets, many[k].info.n_hap) == \
                   (one.info.n_nodes, one.info.n_edges, one.info.n_sets, one.info.n_hap)
            a, b = many[k].extract([(0, len(ref))], 21).host(), one.extract([(0, len(ref))], 21).host()
            assert all(np.array_equal(a[c], b[c]) for c in a)
    assert DeviceGraph.build_many(ctx, []) == []
    bad = list(items)
    bad[2] = ("x", cases[2][0], {"pos": np.array([5, 3], np.int64), "ref_len": np.array([1, 1], np.int32),
                                 "alt_off": np.array([0, 1, 2], np.int64), "alt": np.frombuffer(b"AC", np.uint8)}, None, None)
    with pytest.raises((GrafimoB200Error, ValueError)):
        DeviceGraph.build_many(ctx, bad)


def test_extract_without_haplotypes_and_errors(ctx):
    from grafimo_b200._lib import GrafimoB200Error
    from grafimo_b200.vgraph import VariationGraph
    ref, vs, gt = gr.random_case(7, length=300, n_var=20)
    dg = VariationGraph.build("c", ref, vs, None).to_device(ctx)
    rows = dg.extract([(0, 300)], 11)
    assert rows.n > 0 and int(rows.freq[:rows.n].max().item()) == 0
    from grafimo_b200.extract_regions import DeviceGraph
    dn = DeviceGraph.build(ctx, "c", ref, vs)  # native builder, no genotypes
    rn = dn.extract([(0, 300)], 11)
    assert rn.n == rows.n and bool((rn.packed[:rn.n] == rows.packed[:rows.n]).all()) and int(rn.freq[:rn.n].max().item()) == 0
    with pytest.raises(ValueError):
        DeviceGraph.build(ctx, "c", ref, [(5, "A" if ref[5] != "A" else "C", "G")])  # REF allele mismatch
    empty = dg.extract([], 11)
    assert empty.n == 0
    with pytest.raises(ValueError):
        dg.extract([(0, 10)], 65)
    with pytest.raises(GrafimoB200Error):
        dg.extract([(10, 0)], 11)


def _motif(tmp_path, key="ctcf_meme"):
    from grafimo_b200 import motif_ops as mo
    p = tmp_path / (key + ".meme")
    p.write_text(gu.fixtures()[key])
    return mo.build_motif_meme(str(p), "unfrm_dst", 0.1, False, 1, False, True)[0]


class _Args:
    def __init__(self, threshold=1.0, noqvalue=False, qvalueT=False, noreverse=False, recomb=True):
        self.cores, self.threshold, self.noqvalue, self.qvalueT = 1, float(threshold), noqvalue, qvalueT
        self.noreverse, self.recomb, self.verbose = noreverse, recomb, False


@pytest.mark.parametrize("opts", [dict(threshold=1.0), dict(threshold=0.05, recomb=False), dict(threshold=0.2, noreverse=True),
                                  dict(threshold=0.5, qvalueT=True), dict(threshold=0.01, noqvalue=True),
                                  dict(threshold=1.0, mkey="synth_w35_meme"), dict(threshold=0.3, recomb=False, mkey="synth_w48_meme")])
def test_graph_to_table_equals_tsv_path_and_oracle(ctx, tmp_path, opts, capsys):
    """graph -> K7 -> K2/K5/K6 (no text) == compute_results on the vg-format TSVs of the same regions == the oracle's
    scoring of those TSV rows (score, p, q bit-exact).  The last two cases use motifs wider than one packed word."""
    from grafimo_b200 import score_sequences as ss
    from grafimo_b200.vgraph import VariationGraph
    from oracle import oracle as orc
    ss._ctx = ctx
    opts = dict(opts)
    motif = _motif(tmp_path, opts.pop("mkey", "ctcf_meme"))
    W = motif.width
    ref, vs, gt = gr.random_case(321, length=3000, n_var=150, n_hap=40, n_frac=0.002)
    dg = VariationGraph.build("7", ref, vs, gt).to_device(ctx)
    regions = [(0, 1200), (1100, 3000)]
    rows = dg.extract(regions, W, want_walks=True)
    args = _Args(**opts)
    df = ss.compute_results_rows(motif, rows, True, args)
    d = tmp_path / "seqs" / f"width_{W}"
    d.mkdir(parents=True)
    lines_all = []
    for r, lines in rows.to_vg_tsv().items():
        (d / f"r{r}.tsv").write_text("\n".join(lines) + "\n")
        lines_all += lines
    df2 = ss.compute_results(motif, str(tmp_path / "seqs"), True, args)
    assert list(df.columns) == list(df2.columns)
    gu.assert_tables_equal({c: df[c].to_numpy() for c in df.columns}, {c: df2[c].to_numpy() for c in df2.columns}, list(df.columns))
    out = capsys.readouterr().out
    n = rows.n * (1 if args.noreverse else 2)
    assert out.count(f"Scanned sequences:\t{n}") == 2
    # the oracle on the TSV rows
    f = [ln.split("\t") for ln in lines_all if not (args.noreverse and ln.split("\t")[2][-1] == "-")]
    a = orc.kmers_to_matrix([x[1] for x in f], W)
    _, lo, p = orc.score_rows(a, motif.score_matrix_acgt(), motif.pval_matrix, motif.min_val, motif.scale, float(motif.offset))
    q = orc.bh(p)
    start = np.array([int(x[2].split(":")[1][:-1]) for x in f]); stop = np.array([int(x[3].split(":")[1][:-1]) for x in f])
    freq = np.array([int(x[4]) for x in f])
    keep = (q < args.threshold) if args.qvalueT else (p < args.threshold)
    if not args.recomb:
        keep &= freq > 0
    exp = {"start": start[keep], "stop": stop[keep], "strand": np.array([x[2][-1] for x in f], dtype=object)[keep],
           "score": lo[keep], "p-value": p[keep], "q-value": q[keep],
           "matched_sequence": np.array([x[1] for x in f], dtype=object)[keep], "haplotype_frequency": freq[keep],
           "reference": np.array(["ref" if (x[5] == "ref" and abs(int(b) - int(a_)) == W) else "non.ref"
                                  for x, a_, b in zip(f, start, stop)], dtype=object)[keep]}
    cols = [c for c in df.columns if c in exp and not (c == "q-value" and args.noqvalue)]
    gu.assert_tables_equal({c: df[c].to_numpy() for c in df.columns}, exp, cols)
    assert len(df) > 0


def test_cli_findmotif_from_fasta_vcf_bed(ctx, tmp_path):
    """`findmotif -l FASTA -v VCF -b BED`: graph built and scanned on the GPU == the file interface (scan_graph writes
    vg-format TSVs, compute_results reads them) on the reference's own test.fa / test.vcf.gz."""
    import gzip
    import pandas as pd
    from grafimo_b200 import score_sequences as ss
    from grafimo_b200.__main__ import main
    from grafimo_b200.extract_regions import scan_graph
    from grafimo_b200.vgraph import VariationGraph
    ss._ctx = ctx
    fx = gu.fixtures()
    (tmp_path / "test.fa").write_text(fx["test_fa"])
    with gzip.open(tmp_path / "test.vcf.gz", "wt") as fh:
        fh.write(fx["test_vcf"])
    (tmp_path / "r.bed").write_text("chrx\t0\t30\nchrx\t25\t50\n")
    (tmp_path / "ctcf.meme").write_text(fx["ctcf_meme"])
    out = tmp_path / "out"
    rc = main(["findmotif", "-m", str(tmp_path / "ctcf.meme"), "-l", str(tmp_path / "test.fa"), "-v", str(tmp_path / "test.vcf.gz"),
               "-b", str(tmp_path / "r.bed"), "-t", "1", "--recomb", "-o", str(out), "--debug"])
    assert rc == 0
    got = pd.read_csv(out / "grafimo_out.tsv", sep="\t", index_col=0, float_precision="round_trip")
    # file interface on the same graph
    dg = VariationGraph.from_files(str(tmp_path / "test.fa"), str(tmp_path / "test.vcf.gz"), "x").to_device(ctx)
    loc = scan_graph({"x": dg}, str(tmp_path / "r.bed"), [19], str(tmp_path / "kmers"), True)
    files = sorted(p.name for p in (tmp_path / "kmers" / "width_19").iterdir())
    assert files == ["x_0-30.tsv", "x_25-50.tsv"]
    motif = _motif(tmp_path)
    exp = ss.compute_results(motif, loc, True, _Args(threshold=1.0, recomb=True))
    assert len(got) == len(exp) > 100
    assert set(got["sequence_name"]) == {"x:0-30", "x:25-50"}
    gu.assert_tables_equal({c: got[c].to_numpy() for c in got.columns}, {c: exp[c].to_numpy() for c in exp.columns},
                           [c for c in exp.columns if c not in ("motif_id", "motif_alt_id")])
    assert (out / "grafimo_out.gff").exists()


def test_cli_many_motifs_share_one_extraction_per_width(ctx, tmp_path):
    """A MEME file with three motifs (two of width 19, one of width 8): every motif gets its own report files, the k-mers
    are extracted once per width, and the CTCF report equals the single-motif run."""
    import gzip
    import pandas as pd
    from grafimo_b200 import score_sequences as ss
    from grafimo_b200.__main__ import main
    from grafimo_b200.extract_regions import DeviceGraph
    ss._ctx = ctx
    fx = gu.fixtures()
    (tmp_path / "test.fa").write_text(fx["test_fa"])
    with gzip.open(tmp_path / "test.vcf.gz", "wt") as fh:
        fh.write(fx["test_vcf"])
    (tmp_path / "r.bed").write_text("chrx\t0\t50\n")
    ctcf = fx["ctcf_meme"]
    block = ctcf[ctcf.index("MOTIF"):]
    second = block.replace("MA0139.1", "MA0139.9").replace("CTCF", "CTCFB")
    w8 = fx["synth_w8_meme"]
    (tmp_path / "one.meme").write_text(ctcf)
    (tmp_path / "three.meme").write_text(ctcf.rstrip("\n") + "\n\n" + second.rstrip("\n") + "\n\n" + w8[w8.index("MOTIF"):])
    calls = []
    orig = DeviceGraph.extract

    def counting(self, regions, width, want_walks=False):
        calls.append(width)
        return orig(self, regions, width, want_walks)
    DeviceGraph.extract = counting
    try:
        common = ["-l", str(tmp_path / "test.fa"), "-v", str(tmp_path / "test.vcf.gz"), "-b", str(tmp_path / "r.bed"), "-t", "1",
                  "--recomb", "--debug"]
        assert main(["findmotif", "-m", str(tmp_path / "three.meme"), "-o", str(tmp_path / "out3")] + common) == 0
        assert sorted(calls) == [8, 19]
        assert main(["findmotif", "-m", str(tmp_path / "one.meme"), "-o", str(tmp_path / "out1")] + common) == 0
    finally:
        DeviceGraph.extract = orig
    files = sorted(p.name for p in (tmp_path / "out3").iterdir() if not p.name.endswith(".html"))
    assert files == ["grafimo_out_MA0139.1.gff", "grafimo_out_MA0139.1.tsv", "grafimo_out_MA0139.9.gff", "grafimo_out_MA0139.9.tsv",
                     "grafimo_out_SYN08.1.gff", "grafimo_out_SYN08.1.tsv"], files
    a = pd.read_csv(tmp_path / "out3" / "grafimo_out_MA0139.1.tsv", sep="\t", index_col=0, float_precision="round_trip")
    b = pd.read_csv(tmp_path / "out1" / "grafimo_out.tsv", sep="\t", index_col=0, float_precision="round_trip")
    assert a.equals(b) and len(a) > 50
    c = pd.read_csv(tmp_path / "out3" / "grafimo_out_MA0139.9.tsv", sep="\t", index_col=0, float_precision="round_trip")
    assert c["motif_id"].unique().tolist() == ["MA0139.9"] and c.drop(columns=["motif_id", "motif_alt_id"]).equals(
        a.drop(columns=["motif_id", "motif_alt_id"]))


def test_graph_path_over_two_gpus(ctx, tmp_path):
    """Chromosomes sharded over two ranks (torchrun, NCCL all-reduce of the histogram) == one process with both."""
    import pickle
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    import dist_graph_worker as wk
    from grafimo_b200 import score_sequences as ss
    from grafimo_b200.extract_regions import DeviceGraph
    ss._ctx = ctx
    for r in (0, 1):
        (tmp_path / f"r{r}").mkdir()
    motif = wk.build_motif(str(tmp_path / "r0"))
    rows = []
    for name, ref, vs, gt in wk.chromosomes():
        dg = DeviceGraph.build(ctx, name, ref, vs, gt=gt)
        rows.append(dg.extract([(0, len(ref) // 2), (len(ref) // 2 - 10, len(ref))], motif.width))
    df = ss.compute_results_rows(motif, rows, True, wk.Args)
    assert len(df) > 50
    pickle.dump({c: df[c].to_numpy() for c in df.columns}, open(tmp_path / "expected.pkl", "wb"))
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dist_graph_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29578", worker, str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
