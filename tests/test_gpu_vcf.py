"""GPU tests of K9 (csrc/vcf.cu): the VCF tokenised on the device against the plain-Python reader (vgraph.read_vcf)
and against the variant set the text was written from."""
import gzip

import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from grafimo_b200.engine import Context
    return Context(0)


def vcf_text(chrom, ref, variants, gt, fmt="GT", extra_chrom=True):
    """VCF text of a reduced variant set: indels get their anchor base back, calls are phased diploid."""
    n_hap = gt.shape[1]
    names = [f"S{i}" for i in range(n_hap // 2)]
    out = ["##fileformat=VCFv4.2", "##source=test", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names)]
    suffix = "" if fmt == "GT" else ":7:0.5"
    for (s, r, a), g in zip(variants, gt):
        if r and a:
            pos, R, A = s + 1, r, a
        else:
            pos, R, A = s, ref[s - 1] + r, ref[s - 1] + a
        calls = "\t".join(f"{g[2 * i]}|{g[2 * i + 1]}{suffix}" for i in range(n_hap // 2))
        out.append(f"{chrom}\t{pos}\t.\t{R}\t{A}\t99\tPASS\tAC=1;AN=2\t{fmt}\t{calls}")
        if extra_chrom and s % 7 == 0:
            out.append(f"other\t{pos}\t.\t{R}\t{A}\t99\tPASS\t.\t{fmt}\t{calls}")
    return "\n".join(out) + "\n"


def _as_lists(variants, bits, n_hap):
    pos, rl, off, alt = variants["pos"], variants["ref_len"], variants["alt_off"], variants["alt"]
    v = [(int(pos[i]), int(rl[i]), bytes(alt[off[i]:off[i + 1]]).decode()) for i in range(len(pos))]
    g = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :n_hap]
    return v, g


@pytest.mark.parametrize("fmt,gz", [("GT", False), ("GT:DP:AF", True)])
def test_device_reader_equals_python_reader_and_source(ctx, tmp_path, fmt, gz):
    from grafimo_b200 import synth
    from grafimo_b200.vgraph import read_vcf, read_vcf_device
    ref, variants, gt = synth.variant_set(60_000, 208, 5, indel_frac=0.3)
    text = vcf_text("7", ref, variants, gt, fmt)
    path = tmp_path / ("v.vcf.gz" if gz else "v.vcf")
    if gz:
        with gzip.open(path, "wt") as fh:
            fh.write(text)
    else:
        path.write_text(text)
    pv, pgt, psamples = read_vcf(str(path), "7")
    dv, (bits, n_hap), dsamples = read_vcf_device(ctx, str(path), "7", chunk_bytes=1 << 20)  # several chunks
    assert dsamples == psamples and n_hap == 208
    got_v, got_g = _as_lists(dv, bits, n_hap)
    assert got_v == [(s, len(r), a) for s, r, a in pv] == [(s, len(r), a) for s, r, a in variants]
    assert np.array_equal(got_g, pgt) and np.array_equal(got_g, gt)
    assert dv["calls_out_of_range"] == 0 and dv["lines_with_many_alts_read_on_host"] == 0
    # the REF alleles travel with the variants and are checked against the sequence the graph is built on
    assert bytes(dv["ref"]).decode() == "".join(r for _, r, _ in pv)
    # no chromosome filter: the lines of the other chromosome come too
    dv2, _, _ = read_vcf_device(ctx, str(path), None, chunk_bytes=1 << 20)
    assert len(dv2["pos"]) > len(dv["pos"])
    # one pass, split by chromosome (what the command line does for a BED file with several chromosomes)
    parts, psamples2 = read_vcf_device(ctx, str(path), None, chunk_bytes=1 << 20, by_chrom=True)
    assert sorted(parts) == ["7", "other"] and psamples2 == psamples
    v7, (b7, nh7) = parts["7"]
    assert nh7 == n_hap and np.array_equal(b7, bits)
    assert all(np.array_equal(v7[k], dv[k]) for k in ("pos", "ref_len", "alt_off", "alt", "ref"))
    vo, (bo, _) = parts["other"]
    assert len(vo["pos"]) == len(dv2["pos"]) - len(dv["pos"]) and bo.shape[0] == len(vo["pos"])


def test_device_reader_special_cases(ctx, tmp_path):
    """Multi-allelic lines, missing and unphased calls, haploid-looking calls, symbolic alleles, sites-only lines,
    extra FORMAT keys, a header-only file."""
    from grafimo_b200.vgraph import read_vcf, read_vcf_device
    lines = ["##fileformat=VCFv4.2", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\tB\tC",
             "x\t5\t.\tG\tA,T\t.\t.\t.\tGT\t1|2\t0|1\t2|2",
             "x\t9\trs1\tC\t<DEL>,T\t.\t.\t.\tGT:GQ\t2|0:9\t.|.:0\t1/2:3",
             "x\t12\t.\tCAT\tC,CATAT,*\t.\t.\t.\tGT\t1|2\t3|0\t0|0",
             "x\t20\t.\tA\tg\t.\t.\t.\tGT\t1\t.\t0|1",
             "x\t25\t.\tT\t.\t.\t.\t.\tGT\t0|0\t0|0\t0|0",
             "y\t3\t.\tA\tC\t.\t.\t.\tGT\t1|1\t1|1\t1|1",
             "x\t30\t.\tAC\tGT\t.\t.\t.\tGT\t0|1\t1|0\t1|1",
             # fixed-shape lines (the kernel's fast path): missing alleles, unphased separators, two ALT alleles ...
             "x\t40\t.\tA\tC\t.\t.\t.\tGT\t.|1\t1/0\t./.",
             "x\t42\t.\tG\tC,T\t.\t.\t.\tGT\t2|1\t0/2\t1|.",
             # ... and a line of the same length with one call of another shape: the general path takes the whole line
             "x\t45\t.\tA\tC\t.\t.\t.\tGT\t0|1\t1:9\t1|1"]
    p = tmp_path / "s.vcf"
    p.write_text("\n".join(lines) + "\n")
    pv, pgt, ps = read_vcf(str(p), "x")
    dv, (bits, n_hap), ds = read_vcf_device(ctx, str(p), "x")
    got_v, got_g = _as_lists(dv, bits, n_hap)
    assert ds == ps == ["A", "B", "C"] and n_hap == 6
    assert got_v == [(s, len(r), a) for s, r, a in pv]
    assert np.array_equal(got_g, pgt)
    assert (4, 1, "A") in got_v and (4, 1, "T") in got_v and (12, 2, "") in got_v and (14, 0, "AT") in got_v and (19, 1, "G") in got_v
    # sites-only VCF and header-only VCF
    q = tmp_path / "sites.vcf"
    q.write_text("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\nx\t5\t.\tG\tA\t.\t.\t.\n")
    sv, (sb, sh), ss_ = read_vcf_device(ctx, str(q), "x")
    assert sh == 0 and ss_ == [] and list(sv["pos"]) == [4]
    e = tmp_path / "empty.vcf"
    e.write_text("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\n")
    ev, (eb, eh), es = read_vcf_device(ctx, str(e), "x")
    assert len(ev["pos"]) == 0 and es == ["A"]
    bad = tmp_path / "bad.vcf"
    bad.write_text("#CHROM\tPOS\nx\tnotanumber\t.\tA\tC\t.\t.\t.\n")
    with pytest.raises(ValueError):
        read_vcf_device(ctx, str(bad), "x")


def test_lines_with_many_alt_alleles_and_bad_calls(ctx, tmp_path):
    """A line with more ALT alleles than the genotype kernel builds rows for (16) is read on the host: its alleles get
    their real haplotype sets (never empty ones), exactly what the plain-Python reader gives; a call naming an allele
    the line does not have is counted and warned about."""
    import warnings
    from grafimo_b200.vgraph import read_vcf, read_vcf_device
    alts = ["A" * k + "C" for k in range(1, 19)]  # 18 ALT alleles at one site (insertions of different lengths)
    calls = ["17|18", "0|3", "18|1", "5|5"]
    lines = ["##fileformat=VCFv4.2", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\tB\tC\tD",
             "x\t5\t.\tG\tT\t.\t.\t.\tGT\t1|0\t0|1\t0|0\t1|1",
             "x\t9\t.\tC\t" + ",".join(alts) + "\t.\t.\t.\tGT\t" + "\t".join(calls),
             "x\t20\t.\tA\tG\t.\t.\t.\tGT\t0|1\t1|0\t1|1\t0|0"]
    p = tmp_path / "many.vcf"
    p.write_text("\n".join(lines) + "\n")
    pv, pgt, _ = read_vcf(str(p), "x")
    dv, (bits, n_hap), _ = read_vcf_device(ctx, str(p), "x")
    got_v, got_g = _as_lists(dv, bits, n_hap)
    assert n_hap == 8 and dv["lines_with_many_alts_read_on_host"] == 1 and dv["calls_out_of_range"] == 0
    assert got_v == [(s, len(r), a) for s, r, a in pv] and np.array_equal(got_g, pgt)
    assert got_g.sum() == 4 + 7 + 4  # no carrier was lost: the 18-allele line contributes its 7 non-reference calls
    q = tmp_path / "badcall.vcf"
    q.write_text("\n".join(lines[:2] + ["x\t5\t.\tG\tT\t.\t.\t.\tGT\t1|0\t0|2\t0|0\t1|1"]) + "\n")
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        bv, _, _ = read_vcf_device(ctx, str(q), "x")
    assert bv["calls_out_of_range"] == 1 and any("genotype calls" in str(w.message) for w in rec)


def test_ref_allele_mismatch_is_rejected(ctx, tmp_path):
    """FASTA and VCF that do not belong together (wrong assembly / shifted coordinates): the graph is not built"""
    from grafimo_b200.extract_regions import DeviceGraph
    ref = "ACGTACGTACGTACGTACGTACGTACGTAC"
    good = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\nx\t3\t.\tG\tT\t.\t.\t.\tGT\t0|1\nx\t9\t.\tACG\tA\t.\t.\t.\tGT\t1|0\n"
    (tmp_path / "g.vcf").write_text(good)
    dg = DeviceGraph.from_files(ctx, {"x": ref}, str(tmp_path / "g.vcf"), "x")
    assert dg.extract([(0, 30)], 5).n > 26
    (tmp_path / "b.vcf").write_text(good.replace("x\t3\t.\tG\tT", "x\t4\t.\tG\tT"))  # REF says G, the sequence holds T there
    with pytest.raises(ValueError, match="position 4"):
        DeviceGraph.from_files(ctx, {"x": ref}, str(tmp_path / "b.vcf"), "x")
    (tmp_path / "c.vcf").write_text(good.replace("x\t9\t.\tACG\tA", "x\t29\t.\tACG\tA"))  # runs past the end
    with pytest.raises(ValueError):
        DeviceGraph.from_files(ctx, {"x": ref}, str(tmp_path / "c.vcf"), "x")


def test_graph_from_files_on_reference_fixture(ctx, tmp_path):
    """DeviceGraph.from_files (K9 + gb2_graph_build) on the reference's own test.fa / test.vcf.gz reproduces
    expected_seqs.tsv's k-mers and the oracle's haplotype frequencies."""
    from grafimo_b200.extract_regions import DeviceGraph, decode_kmers
    from oracle import graph_oracle as go
    fx = gu.fixtures()
    (tmp_path / "test.fa").write_text(fx["test_fa"])
    with gzip.open(tmp_path / "test.vcf.gz", "wt") as fh:
        fh.write(fx["test_vcf"])
    dg = DeviceGraph.from_files(ctx, str(tmp_path / "test.fa"), str(tmp_path / "test.vcf.gz"), "x")
    rows = dg.extract([(0, 20)], 19)
    h = rows.host()
    got = sorted(decode_kmers(h["packed"], 19)[i].tobytes().decode() for i in range(rows.n))
    exp = sorted({ln.split("\t")[1] for ln in fx["expected_seqs_tsv"].split("\n") if ln and ln.split("\t")[2].endswith("+")})
    assert got == exp
    ref = "".join(fx["test_fa"].split("\n")[1:])
    variants, gt = go.parse_vcf_text(fx["test_vcf"], "x")
    orows = go.extract_rows(go.build_graph(ref, variants), gt, (0, 50), 19)
    r2 = dg.extract([(0, 50)], 19).host()
    assert sorted(zip(r2["start"].tolist(), r2["freq"].tolist())) == sorted((r["start"], r["freq"]) for r in orows)
