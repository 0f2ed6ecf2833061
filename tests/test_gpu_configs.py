"""Reduced-scale GPU parity tests shaped like the BASELINE.json configurations:
C2 parity form (haplotype windows -> tally -> deduplicated rows -> full hit table vs oracle),
C3 (a JASPAR-sized batch of motifs through the batched DP), C4 (shards + summed histogram = global q),
C5 (long motifs, both strands, no threshold)."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    from grafimo_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _orc():
    from oracle import oracle as orc
    return orc


def test_c3_batched_dp_of_a_jaspar_sized_collection(ctx, tmp_path):
    from grafimo_b200 import motif_ops as mo
    from grafimo_b200 import synth
    orc = _orc()
    text, widths = synth.synthetic_meme_collection(n_motifs=300, seed=7)
    path = tmp_path / "collection.meme"
    path.write_text(text)
    bg = tmp_path / "bg_nt"
    bg.write_text(gu.fixtures()["bg_nt"])
    motifs = mo.build_motif_meme(str(path), str(bg), 0.1, False, 1, False, True)
    assert len(motifs) == 300 and [m.width for m in motifs] == widths.tolist()
    for m in motifs[::7]:  # every 7th motif against the CPU DP, bit for bit
        assert np.array_equal(m.pval_matrix, orc.pval_dp(m.score_matrix_acgt(), m.bg_acgt())), m.motif_id
    for m in motifs:
        assert abs(m.pval_matrix.sum() - 1.0) < 1e-9 and m.pval_matrix.shape == (1000 * m.width + 1,)


def _rows_from_tally(u_pos, u_packed, u_freq, u_isref, w, chrom="5"):
    """vg-find-like rows (both strands) from deduplicated windows."""
    comp = str.maketrans("ACGT", "TGCA")
    lines = []
    for p, x, f, r in zip(u_pos.tolist(), u_packed.tolist(), u_freq.tolist(), u_isref.tolist()):
        seq = "".join("ACGT"[(x >> (2 * i)) & 3] for i in range(w))
        ref = "ref" if r else "non.ref"
        lines.append(f"{chrom}:0-99999\t{seq}\t{chrom}:{p}+\t{chrom}:{p + w}+\t{f}\t{ref}\t1+,")
        lines.append(f"{chrom}:0-99999\t{seq.translate(comp)[::-1]}\t{chrom}:{p + w}-\t{chrom}:{p}-\t{f}\t{ref}\t1-,")
    return lines


def test_c3_many_motif_scan_equals_one_scan_per_motif(ctx):
    """ManyScan (shared hit buffer, ONE sort for all motifs, gb2_finalize_hits_many) == a Scan per motif: same rows, same
    columns, same (p, row, strand) order inside every motif; with and without the q-value filter."""
    from grafimo_b200.engine import ManyScan, Scan
    tags = ["ctcf_meme__unif", "synth_w8_meme__bgnt", "ctcf_meme__bgnt", "synth_w6_meme__bgnt", "synth_w30_meme__bgnt", "synth_w25_meme__bgnt",
            "synth_w8_meme__unif"]
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    dms, sets = [], {}
    for tag in tags:
        m = gu.load_motif(tag)
        dms.append(ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"]))
        w = m["width"]
        if w not in sets:
            sets[w] = torch.randint(0, 1 << (2 * w), (200_003,), dtype=torch.int64, device="cuda", generator=g)
    for thr, qf in ((0.01, False), (0.3, True)):
        many = ManyScan(ctx, dms, strands=2, threshold=thr, hit_capacity=1 << 22)
        for k, dm in enumerate(dms):
            many.score(k, sets[dm.width])
        many.qvalues()
        kept = many.finalize_device(q_filter=qf)
        out = {c: (t[:kept].cpu().numpy() if t is not None else None) for c, t in many.out.items()}
        assert (np.diff(out["motif"]) >= 0).all()  # the rows of one motif are contiguous, motifs in order
        total = 0
        for k, dm in enumerate(dms):
            one = Scan(ctx, dm, strands=2, threshold=thr, hit_capacity=1 << 20)
            one.score(sets[dm.width])
            exp = one.finalize(q_filter=qf)
            sel = out["motif"] == k
            total += len(exp["row"])
            assert np.array_equal(out["row"][sel].astype(np.uint64), exp["row"].astype(np.uint64)), (k, thr)
            assert np.array_equal(out["strand"][sel], exp["strand"]) and np.array_equal(out["iscore"][sel], exp["int_score"])
            assert np.array_equal(out["score"][sel], exp["score"]) and np.array_equal(out["p"][sel], exp["p-value"])
            assert np.array_equal(out["q"][sel], exp["q-value"])
        assert total == kept and (qf or kept > 1000)  # random k-mers: few rows survive the q-value filter


def test_c2_parity_form_haplotypes_to_hit_table(ctx):
    """per-haplotype windows -> gb2_tally_haplotypes -> strands=2 scan of the deduplicated k-mers ==
    the oracle run on the equivalent vg-like TSV rows (frequency, ref flag, p, q, score)."""
    from grafimo_b200 import synth
    from grafimo_b200.engine import Scan
    orc = _orc()
    m = gu.load_motif("ctcf_meme__bgnt")
    w, L, H = 19, 6000, 48
    model = synth.variant_model(L, H, 99, device="cuda")
    codes, idx = synth.haplotype_codes(model, 0, H, return_index=True)
    per = L - w + 1
    packed = synth.pack_windows(codes, w).reshape(-1)
    pos = idx[:, :per].reshape(-1).contiguous()  # reference coordinate of the window start (vg reports graph positions)
    ref = synth.reference_windows(model, w)
    u_pos, u_packed, u_freq, u_isref = ctx.tally_haplotypes(pos, packed.clone(), ref, pos_base=0)
    n = u_pos.shape[0]
    assert per <= n < per * 4 and int(u_freq.sum()) == per * H
    assert 0 < int(u_isref.sum()) <= per
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    thr = 0.01
    sc = Scan(ctx, dm, strands=2, threshold=thr, hit_capacity=2 * n)
    kp = u_packed.contiguous()
    if kp.data_ptr() % 16:
        kp = kp.clone()
    sc.score(kp)
    out = sc.finalize()
    lines = _rows_from_tally(u_pos, u_packed, u_freq, u_isref, w)
    exp = orc.compute_results(m, lines, threshold=thr, recomb=True)
    rows = out["row"].astype(np.int64)
    up, uf, ur = u_pos.cpu().numpy(), u_freq.cpu().numpy(), u_isref.cpu().numpy()
    strand = np.where(out["strand"] == 0, "+", "-").astype(object)
    start = np.where(out["strand"] == 0, up[rows], up[rows] + w)
    stop = np.where(out["strand"] == 0, up[rows] + w, up[rows])
    seqs = np.array([ln.split("\t")[1] for ln in lines], dtype=object)[2 * rows + out["strand"]]
    got = {"start": start, "stop": stop, "strand": strand, "score": out["score"], "p-value": out["p-value"],
           "q-value": out["q-value"], "matched_sequence": seqs, "haplotype_frequency": uf[rows].astype(np.int64),
           "reference": np.where(ur[rows] == 1, "ref", "non.ref").astype(object)}
    gu.assert_tables_equal(got, exp, list(got.keys()))


def test_c4_sharded_scan_gives_the_global_qvalues(ctx):
    """Two shards scored separately, histograms summed (what the NCCL all-reduce does) == one scan of everything."""
    from grafimo_b200.engine import Scan
    from grafimo_b200 import dist as gdist
    m = gu.load_motif("ctcf_meme__unif")
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    n = 3_000_001
    packed = torch.randint(0, 1 << 38, (n + 1,), dtype=torch.int64, device="cuda", generator=g)[:n]
    whole = Scan(ctx, dm, strands=2, threshold=1e-3, hit_capacity=1 << 16)
    whole.score(packed)
    ref = whole.finalize()
    tables, scans = [], []
    for r in range(2):
        lo, hi = gdist.shard_bounds(n, r, 2)
        s = Scan(ctx, dm, strands=2, threshold=1e-3, hit_capacity=1 << 16)
        s.score(packed[lo:hi], row_base=lo)
        scans.append(s)
    ctx.sync()
    total = scans[0].histogram() + scans[1].histogram()
    for s in scans:
        s.histogram().copy_(total)
        tables.append(s.finalize())
    merged = gdist.merge_hit_tables(tables)
    for k in ("row", "strand", "int_score", "score", "p-value", "q-value"):
        assert np.array_equal(merged[k], ref[k]), k
    assert tables[0]["total"] == 2 * n


def test_c5_long_motifs_both_strands_no_threshold(ctx):
    """w = 25..64, threshold 1 (every window with p < 1 is reported), both strands, N rows included."""
    from grafimo_b200.engine import Scan
    orc = _orc()
    rng = np.random.default_rng(21)
    for tag in ("synth_w25_meme__bgnt", "synth_w27_meme__bgnt", "synth_w30_meme__bgnt", "synth_w32_meme__bgnt",
                "synth_w35_meme__bgnt", "synth_w64_meme__bgnt"):
        m = gu.load_motif(tag)
        w, n = m["width"], 2500
        seqs = ["".join(rng.choice(list("ACGT"), size=w)) for _ in range(n)]
        seqs[11] = seqs[11][:5] + "N" + seqs[11][6:]
        a = orc.kmers_to_matrix(seqs, w)
        packed, nmask, _ = ctx.encode(torch.from_numpy(a).cuda())
        dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
        sc = Scan(ctx, dm, strands=2, threshold=1.0, hit_capacity=2 * n)
        sc.score(packed, nmask)
        out = sc.finalize()
        comp = str.maketrans("ACGTN", "TGCAN")
        a_r = orc.kmers_to_matrix([s.translate(comp)[::-1] for s in seqs], w)
        args = (m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
        isf, lof, pf = orc.score_rows(a, *args, nthreads=4)
        isr, lor, pr = orc.score_rows(a_r, *args, nthreads=4)
        p_all = np.concatenate([pf, pr]); q_all = orc.bh(p_all); lo_all = np.concatenate([lof, lor])
        idx = out["row"].astype(np.int64) + n * out["strand"].astype(np.int64)
        assert sorted(idx.tolist()) == np.nonzero(p_all < 1.0)[0].tolist()
        assert np.array_equal(out["p-value"], p_all[idx]) and np.array_equal(out["q-value"], q_all[idx])
        assert np.array_equal(out["score"], lo_all[idx])
