"""GPU tests of K8 (csrc/report.cu): TSV / GFF3 bytes written from device-resident hit columns against the host
writers (pandas to_csv / gff3_lines, which are pinned on reference-generated goldens) on the same rows."""
import numpy as np
import pandas as pd
import pytest

import golden_util as gu
import graph_util as gr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from grafimo_b200.engine import Context
    return Context(0)


class _Args:
    def __init__(self, threshold=1.0, noqvalue=False, qvalueT=False, noreverse=False, recomb=True, outdir="out", text_only=True):
        self.cores, self.threshold, self.noqvalue, self.qvalueT = 1, float(threshold), noqvalue, qvalueT
        self.noreverse, self.recomb, self.verbose, self.outdir, self.text_only, self.top_graphs = noreverse, recomb, False, outdir, text_only, 0


def _motif(tmp_path, key="ctcf_meme"):
    from grafimo_b200 import motif_ops as mo
    p = tmp_path / f"{key}.meme"
    p.write_text(gu.fixtures()[key])
    return mo.build_motif_meme(str(p), "unfrm_dst", 0.1, False, 1, False, True)[0]


@pytest.mark.parametrize("opts", [dict(threshold=1.0), dict(threshold=0.05, recomb=False), dict(threshold=0.3, noreverse=True),
                                  dict(threshold=1.0, qvalueT=True), dict(threshold=0.02, noqvalue=True),
                                  dict(threshold=1.0, mkey="synth_w35_meme"), dict(threshold=0.4, recomb=False, mkey="synth_w64_meme")])
def test_device_writer_equals_host_writers(ctx, tmp_path, opts):
    from grafimo_b200 import score_sequences as ss
    from grafimo_b200.extract_regions import DeviceGraph
    from grafimo_b200.res_writer import write_results, write_results_device
    ss._ctx = ctx
    opts = dict(opts)
    motif = _motif(tmp_path, opts.pop("mkey", "ctcf_meme"))  # the last two: motifs wider than one packed word
    W = motif.width
    ref, vs, gt = gr.random_case(77, length=5000, n_var=250, n_hap=40)
    rows = [DeviceGraph.build(ctx, "7", ref, vs, gt=gt).extract([(0, 2100), (2000, 5000)], W)]
    ref2, vs2, gt2 = gr.random_case(78, length=1500, n_var=60, n_hap=40)
    rows.append(DeviceGraph.build(ctx, "X", ref2, vs2, gt=gt2).extract([(10, 1500)], W))
    a_host = _Args(outdir=str(tmp_path / "host"), **opts)
    a_dev = _Args(outdir=str(tmp_path / "dev"), **opts)
    df = ss.compute_results_rows(motif, rows, True, a_host)
    write_results(df, motif, 1, a_host, True)
    report = ss.scan_rows_device(motif, rows, True, a_dev)
    assert report.n == len(df) > 20
    write_results_device(report, motif, 1, a_dev, True)
    for ext in ("tsv", "gff"):
        host = (tmp_path / "host" / f"grafimo_out.{ext}").read_text().split("\n")
        dev = (tmp_path / "dev" / f"grafimo_out.{ext}").read_text().split("\n")
        assert host[0] == dev[0] and host[-1] == dev[-1] == "" and len(host) == len(dev) == len(df) + 2
        if ext == "tsv":  # the index column follows the row order, which differs between the two paths only inside p-value ties
            assert [ln.split("\t", 1)[0] for ln in dev[1:-1]] == [str(i) for i in range(len(df))]
            strip = lambda ln: ln.split("\t", 1)[1]  # noqa: E731
            assert sorted(map(strip, host[1:-1])) == sorted(map(strip, dev[1:-1]))
            pcol = 8
            p = np.array([float(ln.split("\t")[pcol]) for ln in dev[1:-1]])
            assert (np.diff(p) >= 0).all()
        else:
            assert sorted(host[1:-1]) == sorted(dev[1:-1])
    # the DataFrame view of the device report is the same table
    d2 = report.to_df()
    gu.assert_tables_equal({c: d2[c].to_numpy() for c in d2.columns}, {c: df[c].to_numpy() for c in df.columns}, list(df.columns))
    assert list(d2.columns) == list(df.columns)


def test_device_writer_large_and_long_motif(ctx, tmp_path):
    """Every window reported (-t 1) for a w = 30 motif on a denser graph: 1e5+ rows, all bytes equal to the host writers."""
    from grafimo_b200 import score_sequences as ss
    from grafimo_b200.extract_regions import DeviceGraph
    from grafimo_b200.res_writer import gff3_lines
    ss._ctx = ctx
    motif = _motif(tmp_path, "synth_w30_meme")
    ref, vs, gt = gr.random_case(5, length=40000, n_var=1500, n_hap=64, indel=0.3)
    rows = DeviceGraph.build(ctx, "3", ref, vs, gt=gt).extract([(0, 40000)], 30)
    a = _Args(threshold=1.0, recomb=True)
    report = ss.scan_rows_device(motif, rows, True, a)
    assert report.n > 100_000
    df = report.to_df()
    tsv_dev = report.render(0).tobytes().decode().split("\n")
    import io
    buf = io.StringIO()
    df.to_csv(buf, sep="\t", encoding="utf-8")
    tsv_host = buf.getvalue().split("\n")
    assert tsv_host[1:] == tsv_dev  # same order here: both come from the device report
    assert tsv_host[0] + "\n" == report.tsv_header().decode()
    gff_dev = report.render(1).tobytes().decode()
    assert gff_dev == "".join(gff3_lines(df, False, True))


@pytest.mark.parametrize("tag", ["fixture_testmode", "fixture_t05_norecomb", "fixture_plus_N_2files", "fixture_qvalT", "synth_w8",
                                 "synth_w33", "synth_w64"])
def test_tsv_directory_to_files_on_device(ctx, tmp_path, tag):
    """`vg find` TSV directory -> K1b/K2/K5/K6 -> K8 files == compute_results + host writers on the reference-generated
    golden cases (same line sets; the order differs only inside p-value ties)."""
    from grafimo_b200 import motif_ops as mo
    from grafimo_b200 import score_sequences as ss
    from grafimo_b200.res_writer import write_results, write_results_device
    ss._ctx = ctx
    tags = gu.scoring_tags()
    if tag not in tags:
        pytest.skip(f"no golden case {tag}")
    c = gu.load_scoring(tag)
    g = gu.load_motif(c["motif_tag"])
    fx = gu.fixtures()
    key = g["source"] if g["source"].endswith("_" + g["fmt"]) else g["source"] + "_meme"
    (tmp_path / "m.meme").write_text(fx[key])
    (tmp_path / "bg_nt").write_text(fx["bg_nt"])
    bg = "unfrm_dst" if g["bgfile"] == "unif" else str(tmp_path / "bg_nt")
    m = mo.build_motif_meme(str(tmp_path / "m.meme"), bg, g["pseudo"], g["no_reverse"], 1, False, True)[0]
    d = tmp_path / "seqs" / f"width_{m.width}"
    d.mkdir(parents=True)
    for k, lines in enumerate(c["files"]):
        # one file per region name, as vg writes them
        by_name = {}
        for ln in lines:
            by_name.setdefault(ln.split()[0], []).append(ln)
        for j, (nm, ls) in enumerate(by_name.items()):
            (d / f"region_{k}_{j}.tsv").write_text("\n".join(ls) + "\n")
    o = c["options"] if tag != "fixture_testmode" else dict(threshold=1.0, noqvalue=False, qvalueT=False, noreverse=False, recomb=True)
    a_host = _Args(outdir=str(tmp_path / "host"), threshold=o["threshold"], noqvalue=o["noqvalue"], qvalueT=o["qvalueT"],
                   noreverse=o["noreverse"], recomb=o["recomb"])
    a_dev = _Args(outdir=str(tmp_path / "dev"), threshold=o["threshold"], noqvalue=o["noqvalue"], qvalueT=o["qvalueT"],
                  noreverse=o["noreverse"], recomb=o["recomb"])
    df = ss.compute_results(m, str(tmp_path / "seqs"), True, a_host)
    gu.assert_tables_equal({col: df[col].to_numpy() for col in df.columns}, c["table"], c["columns"])  # still the reference's table
    write_results(df, m, 1, a_host, True)
    report = ss.scan_dir_device(m, str(tmp_path / "seqs"), True, a_dev)
    lower = any(ln.split()[1] != ln.split()[1].upper() for f in c["files"] for ln in f)
    if lower:  # lower-case k-mers keep their case in the report: the general path must be asked for
        assert report is None
        return
    assert report is not None and report.n == len(df)
    write_results_device(report, m, 1, a_dev, True)
    for ext in ("tsv", "gff"):
        host = (tmp_path / "host" / f"grafimo_out.{ext}").read_text().split("\n")
        dev = (tmp_path / "dev" / f"grafimo_out.{ext}").read_text().split("\n")
        assert host[0] == dev[0] and len(host) == len(dev)
        strip = (lambda ln: ln.split("\t", 1)[1]) if ext == "tsv" else (lambda ln: ln)
        assert sorted(map(strip, host[1:-1])) == sorted(map(strip, dev[1:-1]))


def test_tsv_directory_fallback_conditions(ctx, tmp_path):
    from grafimo_b200 import score_sequences as ss
    ss._ctx = ctx
    m = _motif(tmp_path)
    d = tmp_path / "k" / "width_19"
    d.mkdir(parents=True)
    lines = [ln for ln in gu.fixtures()["scoring_input_tsv"].split("\n") if ln][:50]
    (d / "a.tsv").write_text("\n".join(lines) + "\n")
    a = _Args(threshold=1.0)
    assert ss.scan_dir_device(m, str(tmp_path / "k"), True, a) is not None
    mixed = lines[:25] + [ln.replace(ln.split()[0], "other:1-2", 1) for ln in lines[25:]]
    (d / "a.tsv").write_text("\n".join(mixed) + "\n")
    assert ss.scan_dir_device(m, str(tmp_path / "k"), True, a) is None  # two region names in one file
    low = [ln.split("\t")[0] + "\t" + ln.split("\t")[1].lower() + "\t" + "\t".join(ln.split("\t")[2:]) for ln in lines]
    (d / "a.tsv").write_text("\n".join(low) + "\n")
    assert ss.scan_dir_device(m, str(tmp_path / "k"), True, a) is None  # lower-case k-mers keep their case in the general path
