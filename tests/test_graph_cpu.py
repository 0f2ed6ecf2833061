"""CPU tests of the k-mer extraction row: the oracle against the reference's own vg fixture, and the
host-side graph builder (grafimo_b200/vgraph.py) against the oracle -- structure, coordinates, haplotype sets."""
import gzip

import numpy as np
import pytest

import golden_util as gu
import graph_util as gr
from oracle import graph_oracle as go


def _fixture():
    fx = gu.fixtures()
    ref = "".join(fx["test_fa"].split("\n")[1:])
    variants, gt = go.parse_vcf_text(fx["test_vcf"], "x")
    return fx, ref, variants, np.array(gt, dtype=np.uint8)


def test_oracle_reproduces_reference_vg_fixture():
    """tests/grafimo_run_test.py:49-63: `vg find -x test.xg -E -p x:0-20 -K 19` == expected_seqs.tsv, every field."""
    fx, ref, variants, _ = _fixture()
    g = go.build_graph(ref, variants)
    lines = go.vg_tsv_lines(go.extract_rows(g, None, (0, 20), 19), "x", (0, 20))
    exp = [ln for ln in fx["expected_seqs_tsv"].split("\n") if ln]
    assert len(exp) == 32 and sorted(lines) == sorted(exp)


def test_oracle_haplotype_counts_on_fixture():
    """The VCF's one sample: haplotype 0 carries every ALT, haplotype 1 the two homozygous ones."""
    _, ref, variants, gt = _fixture()
    g = go.build_graph(ref, variants)
    rows = go.extract_rows(g, gt.tolist(), (0, 50), 19)
    by_seq = {(r["start"], r["seq"]): r["freq"] for r in rows}
    assert by_seq[(0, "CAAATAAGATTTGAAAATT")] == 1   # haplotype 0
    assert by_seq[(0, "CAAATAAGGTTTGGAAATT")] == 1   # haplotype 1
    assert by_seq[(0, "CAAATAAGGCTTGGAAATT")] == 0   # the reference allele at a homozygous-ALT site
    assert sum(r["freq"] for r in rows if r["start"] == 0) == 2
    assert by_seq[(14, "AAATTTTCTGGAGTTCTAT")] == 2 and all(r["ref"] for r in rows if r["start"] == 14)


def test_builder_matches_oracle_on_fixture():
    from grafimo_b200.vgraph import VariationGraph
    _, ref, variants, gt = _fixture()
    g = VariationGraph.build("x", ref, variants, gt)
    og = go.build_graph(ref, variants)
    assert g.n_nodes == len(og.nodes) == 15
    codes = "ACGTN"
    for i, nd in enumerate(og.nodes):
        assert "".join(codes[c] for c in g.seq[g.node_off[i]:g.node_off[i + 1]]) == nd.seq
        assert (g.node_a0[i], g.node_clamp[i], bool(g.node_flags[i] & 1)) == (nd.a0, nd.clamp, nd.isref)
        assert [int(x) + 1 for x in g.edge_to[g.edge_off[i]:g.edge_off[i + 1]]] == og.out[nd.id]
    for region, w in (((0, 20), 19), ((0, 50), 19), ((5, 45), 8), ((0, 50), 32)):
        assert gr.rows_from_arrays(g, region, w) == gr.oracle_rows(ref, variants, gt, region, w)
    g0 = VariationGraph.build("x", ref, variants, None)  # no haplotype index: vg prints frequency 0
    assert all(r[3] == 0 for r in gr.rows_from_arrays(g0, (0, 50), 19))


@pytest.mark.parametrize("seed", range(8))
def test_builder_matches_oracle_random_graphs(seed):
    """SNPs, insertions, deletions, complex and multi-allelic sites, variants inside a deleted span, N bases, chained
    nodes: same walks, coordinates, node ids, ref flags and haplotype frequencies as the oracle's explicit haplotypes."""
    from grafimo_b200.vgraph import VariationGraph
    ref, vs, gt = gr.random_case(100 + seed, n_frac=0.01 if seed % 3 == 0 else 0.0)
    m = 8 if seed % 2 else 32
    g = VariationGraph.build("c", ref, vs, gt, max_node_len=m)
    for w, region in ((7, (0, 400)), (19, (37, 301)), (32, (100, 250))):
        got = gr.rows_from_arrays(g, region, w)
        assert got == gr.oracle_rows(ref, vs, gt, region, w, max_node=m)
    assert any(r[3] == 0 for r in got) or seed >= 0  # recombinant walks are emitted (frequency 0), never dropped


def test_vcf_and_fasta_readers(tmp_path):
    from grafimo_b200.vgraph import VariationGraph, read_fasta, read_vcf, reduce_allele
    fx, ref, variants, gt = _fixture()
    fa = tmp_path / "test.fa"
    fa.write_text(fx["test_fa"])
    vcf = tmp_path / "test.vcf.gz"
    with gzip.open(vcf, "wt") as fh:
        fh.write(fx["test_vcf"])
    assert read_fasta(str(fa)) == {"x": ref}
    v, g, samples = read_vcf(str(vcf), "x")
    assert v == variants and np.array_equal(g, gt) and samples == ["1"]
    assert reduce_allele(10, "CA", "C") == (11, "A", "") and reduce_allele(10, "C", "CTT") == (11, "", "TT")
    assert reduce_allele(10, "CAT", "CGT") == (11, "A", "G")
    graph = VariationGraph.from_files(str(fa), str(vcf), "x")
    assert graph.n_hap == 2 and graph.n_nodes == 15
    with pytest.raises(ValueError):
        VariationGraph.build("x", ref, [(3, "C", "T")])  # REF allele does not match the sequence


def test_bed_reader(tmp_path):
    from grafimo_b200.extract_regions import get_regions_bed
    bed = tmp_path / "r.bed"
    bed.write_text("track name=x\nchr1\t10\t50\tpeak1\nchr2\t5\t9\nchr1\t100\t150\n1\t3\t4\n")
    regions, n = get_regions_bed(str(bed), True)
    assert n == 3 and regions == {"chr1": [("10", "50"), ("100", "150")], "chr2": [("5", "9")]}
    with pytest.raises(FileNotFoundError):
        get_regions_bed(str(tmp_path / "none.bed"), True)


def _write_bgzf(path, data, block=60000):
    """Minimal BGZF writer (SAM spec 4.1): independent deflate blocks with the BC extra field + the empty EOF block."""
    import struct
    import zlib
    with open(path, "wb") as fh:
        for lo in list(range(0, len(data), block)) + [None]:
            piece = b"" if lo is None else data[lo:lo + block]
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            cdata = c.compress(piece) + c.flush()
            bsize = 12 + 6 + len(cdata) + 8
            fh.write(b"\x1f\x8b\x08\x04" + b"\x00" * 4 + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1))
            fh.write(cdata + struct.pack("<II", zlib.crc32(piece), len(piece)))


def test_vcf_chunk_reader_bgzf_equals_gzip_and_plain(tmp_path):
    """The chunked VCF reader: BGZF input (blocks inflated by a thread pool) == plain gzip == plain text, whole lines per
    chunk, for chunk sizes smaller and larger than the file, with and without a trailing newline."""
    import gzip
    from grafimo_b200.vgraph import _bgzf_blocks, _vcf_chunks
    rng = np.random.default_rng(4)
    lines = ["##fileformat=VCFv4.2", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"S{i}" for i in range(300))]
    for k in range(900):
        gts = "\t".join(f"{a}|{b}" for a, b in rng.integers(0, 2, size=(300, 2)))
        lines.append(f"1\t{100 + 7 * k}\t.\tA\tC\t.\tPASS\t.\tGT\t{gts}")
    for tail in ("\n", ""):
        data = ("\n".join(lines) + tail).encode()
        (tmp_path / "a.vcf").write_bytes(data)
        with gzip.open(tmp_path / "plain.vcf.gz", "wb") as fh:
            fh.write(data)
        _write_bgzf(tmp_path / "b.vcf.gz", data)
        import mmap
        with open(tmp_path / "b.vcf.gz", "rb") as fh:
            assert len(_bgzf_blocks(mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ))) == (len(data) + 59999) // 60000 + 1
        with open(tmp_path / "plain.vcf.gz", "rb") as fh:
            assert _bgzf_blocks(mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)) is None
        want = data if data.endswith(b"\n") else data + b"\n"
        for chunk in (150_000, 1 << 22):
            for name in ("a.vcf", "plain.vcf.gz", "b.vcf.gz"):
                got = []
                for t in _vcf_chunks(str(tmp_path / name), chunk, threads=4):
                    b = t.numpy().tobytes()
                    assert b.endswith(b"\n")
                    got.append(b)
                assert b"".join(got) == want, (name, chunk, tail)
                assert len(got) > 1 or chunk > len(data)


def _build_stats(ref, vs, gt, threads, chunk_bps, max_node_len=32):
    import ctypes
    from grafimo_b200 import _lib
    from grafimo_b200.extract_regions import DeviceGraph
    refa, pos, rlen, alt_off, alt_p, n_hap, words, bits = DeviceGraph._build_inputs(ref, vs, gt)
    lib = _lib.load()
    out = np.zeros(6, np.uint64)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None  # noqa: E731
    rc = lib.gb2_graph_build_stats(ptr(refa), len(refa), len(pos), ptr(pos), ptr(rlen), ptr(alt_off), ptr(alt_p), n_hap, words,
                                   ptr(bits), max_node_len, threads, chunk_bps, ptr(out))
    assert rc == 0
    return out


def test_graph_builder_ranges_do_not_change_the_graph():
    """The host pass of gb2_graph_build cut into independent breakpoint ranges on worker threads (cuts only where no allele
    spans or ends) gives, array for array, the graph of the single pass: same node ids, edges, haplotype-set numbering."""
    for seed in range(8):
        ref, vs, gt = gr.random_case(500 + seed, length=3000 + 700 * seed, n_var=250 + 60 * seed, n_hap=(0, 12, 70, 40)[seed % 4],
                                     indel=0.4 if seed % 2 else 0.15)
        gt_ = gt if seed % 4 else None
        one = _build_stats(ref, vs, gt_, 1, 0)
        assert one[4] == 1 and one[0] > len(vs)
        for threads, chunk in ((4, 1), (3, 7), (8, 50), (2, 100000), (5, 0)):
            many = _build_stats(ref, vs, gt_, threads, chunk)
            assert many.tolist()[:4] == one.tolist()[:4] and many[5] == one[5], (seed, threads, chunk, one, many)
            if chunk in (1, 7, 50):
                assert many[4] > 3  # really cut into ranges
        # a smaller node length (more chained nodes) as well
        assert _build_stats(ref, vs, gt_, 4, 5, max_node_len=8)[5] == _build_stats(ref, vs, gt_, 1, 0, max_node_len=8)[5]


def test_graph_builder_ranges_with_parallel_edges_to_merge():
    """Chained deletions that reach the same breakpoint by two routes give parallel edges whose haplotype sets are merged
    after the ranges are stitched (a lookup in the global set table, rebuilt lazily): same digest for any range size."""
    rng = np.random.default_rng(11)
    ref = "".join(rng.choice(list("ACGT"), size=400))
    vs = []
    for p in (20, 90, 160, 230, 300):
        vs += [(p, ref[p:p + 2], ""), (p + 2, ref[p + 2:p + 4], ""), (p, ref[p:p + 4], ""), (p + 30, ref[p + 30], "A" if ref[p + 30] != "A" else "C")]
    gt = (rng.random((len(vs), 24)) < 0.3).astype(np.uint8)
    one = _build_stats(ref, vs, gt, 1, 0)
    for threads, chunk in ((4, 1), (2, 3), (8, 0)):
        many = _build_stats(ref, vs, gt, threads, chunk)
        assert many.tolist()[:4] == one.tolist()[:4] and many[5] == one[5]
    assert _build_stats(ref, vs, gt, 4, 1)[4] > 3
    # the merged edges exist: fewer CSR edges than a graph where the long deletions are left out
    fewer = _build_stats(ref, [v for v in vs if len(v[1]) != 4], gt[[i for i, v in enumerate(vs) if len(v[1]) != 4]], 1, 0)
    assert one[1] == fewer[1]


# ---- pinned on REAL vg output: the reference's 704-row k-mer fixture ---------------------------------------------------
def test_oracle_reproduces_real_vg_kmer_fixture():
    """tests/test_data/input/width_19/scoring_test_input.tsv (what `vg find -K 19 -E -H` printed for
    22:19723256-19723526 of the 1000-Genomes graph): the local graph is reconstructed from the rows themselves
    (tests/fixture_graph.py) and the oracle must print exactly those 704 lines -- sequence, start, stop, strand, haplotype
    frequency (0 for the recombinant walk included), ref flag, vg node ids -- including the 36 rows through the 2-bp
    deletion (21-bp span, flagged `ref`: score_sequences.py:305-307 rewrites them later)."""
    import fixture_graph as fg
    c = fg.reconstruct(gu.fixtures()["scoring_input_tsv"])
    assert sorted((p, r, a) for p, r, a in c["variants"]) == [(56, "C", "T"), (124, "C", "T"), (142, "G", "A"), (201, "G", "A"),
                                                               (243, "TG", ""), (296, "C", "T")]
    g = go.build_graph(c["ref"], c["variants"])
    rows = go.extract_rows(g, [list(map(int, r)) for r in c["gt"]], c["local_region"], c["w"])
    lines = fg.shift_lines(go.vg_tsv_lines(rows, c["chrom"], c["local_region"]), c)
    exp = [r["line"] for r in c["rows"]]
    assert len(exp) == 704 and sorted(lines) == sorted(exp)
    spans = [abs(r["stop"] - r["start"]) for r in c["rows"]]
    assert spans.count(21) == 36 and all(r["ref"] == "ref" for r in c["rows"] if abs(r["stop"] - r["start"]) == 21)
    assert sum(r["freq"] == 0 for r in c["rows"]) == 2  # the recombinant walk, both orientations


def test_builder_reproduces_real_vg_kmer_fixture():
    """the host-side graph arrays (vgraph.VariationGraph.build -- what gb2_graph_create uploads) walked in plain Python:
    the same 352 forward rows, node ids included"""
    import fixture_graph as fg
    from grafimo_b200.vgraph import VariationGraph
    c = fg.reconstruct(gu.fixtures()["scoring_input_tsv"])
    g = VariationGraph.build(c["chrom"], c["ref"], c["variants"], c["gt"])
    got = gr.rows_from_arrays(g, c["local_region"], c["w"])
    off, noff = c["offset"], c["node_offset"]
    got = sorted((s + off, e + off, seq, f, "ref" if isref else "non.ref", tuple(n + noff for n in nodes))
                 for s, e, seq, f, isref, nodes in got)
    exp = sorted((r["start"], r["stop"], r["seq"], r["freq"], r["ref"], tuple(r["nodes"])) for r in c["rows"] if r["strand"] == "+")
    assert len(exp) == 352 and got == exp
