"""Size-independent properties at the FULL size of the headline configuration (BASELINE.json configs[1]:
CTCF, 1 Mb x 2,504 haplotype paths = 2.5e9 packed k-mers, 20 GB, both strands): the oracle cannot score that
row by row, so the run is checked through invariants plus an oracle check of every 100th reported hit."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_full_size_headline_workload_properties():
    from grafimo_b200 import synth
    from grafimo_b200.engine import Context, Scan
    from oracle import oracle as orc
    free, total = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~45 GB of free device memory")
    ctx = Context(0)
    m = gu.load_motif("ctcf_meme__unif")
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    w, L, H, thr = 19, 1_000_000, 2504, 1e-4
    with torch.cuda.stream(ctx.stream):
        windows, model = synth.haplotype_windows(L, H, w, 20240, device=ctx.device, hap_batch=32)
    n = windows.shape[0]
    assert n == (L - w + 1) * H
    whole = Scan(ctx, dm, strands=2, threshold=thr, hit_capacity=1 << 22)
    whole.score(windows)
    out = whole.finalize()
    hist = whole.histogram().cpu().numpy()
    # (1) every window lands in exactly one bin; (2) hits == histogram mass above the p-value cut
    assert int(hist.sum()) == 2 * n == out["total"]
    cut = int(np.nonzero(dm.ptable < thr)[0][0])
    assert int(hist[cut:dm.span].sum()) == len(out["row"]) and hist[dm.span] == 0
    # (3) sorted by p, p and q are functions of the score, q >= p, q non-decreasing along the table
    assert np.all(np.diff(out["p-value"]) >= 0) and np.all(np.diff(out["q-value"]) >= 0)
    assert np.array_equal(out["p-value"], dm.ptable[out["int_score"] - dm.lo]) and np.all(out["q-value"] >= out["p-value"])
    # (4) two shards with global row indices + summed histograms == the whole scan (what N GPUs do)
    half = (n // 2) & ~1
    parts = []
    scans = [Scan(ctx, dm, strands=2, threshold=thr, hit_capacity=1 << 22) for _ in range(2)]
    scans[0].score(windows[:half], row_base=0)
    scans[1].score(windows[half:], row_base=half)
    ctx.sync()
    tot = scans[0].histogram() + scans[1].histogram()
    assert torch.equal(tot, whole.histogram())
    for s in scans:
        s.histogram().copy_(tot)
        parts.append(s.finalize())
    from grafimo_b200 import dist as gdist
    merged = gdist.merge_hit_tables(parts)
    for k in ("row", "strand", "int_score", "p-value", "q-value", "score"):
        assert np.array_equal(merged[k], out[k]), k
    # (5) every 100th hit re-scored by the oracle from its k-mer
    sel = np.arange(0, len(out["row"]), 100)
    rows = torch.from_numpy(out["row"][sel].astype(np.int64)).to(ctx.device)
    xs = windows[rows].cpu().numpy()
    seqs = ["".join("ACGT"[(int(v) >> (2 * i)) & 3] for i in range(w)) for v in xs]
    comp = str.maketrans("ACGT", "TGCA")
    seqs = [s if st == 0 else s.translate(comp)[::-1] for s, st in zip(seqs, out["strand"][sel])]
    isc, lo, pv = orc.score_rows(orc.kmers_to_matrix(seqs, w), m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"],
                                 m["offset"], nthreads=8)
    assert np.array_equal(isc, out["int_score"][sel]) and np.array_equal(pv, out["p-value"][sel])
    assert np.array_equal(lo, out["score"][sel])
    # (6) the q-values are exactly BH over the histogram (numpy restatement of statsmodels on the binned p-values)
    ptab_all = np.concatenate([dm.ptable, [1.0]])
    order = np.argsort(ptab_all, kind="stable")
    c = np.cumsum(hist[order])
    with np.errstate(divide="ignore", invalid="ignore"):
        raw = np.where(hist[order] > 0, ptab_all[order] / (c / float(2 * n)), np.inf)
    q = np.minimum.accumulate(raw[::-1])[::-1]
    q[q > 1] = 1
    qtab = np.empty_like(q)
    qtab[order] = q
    assert np.array_equal(out["q-value"], qtab[out["int_score"] - dm.lo])
    ctx.close()


def test_full_size_unselective_scan_properties():
    """BASELINE config 5 at its full size (w = 25, 30 M k-mers, both strands, threshold 1: 60 M report rows) through the dense
    form (K2 dense scores -> gb2_finalize_dense, two partition passes): the oracle cannot score that row by row, so the
    table is checked through what the sort must guarantee -- every window with p < 1 is reported exactly once, the rows are
    ordered by (p-rank, row, strand), p / q / score are functions of the integer score, the per-score row counts equal K2's
    histogram -- plus an oracle check of a sample of rows."""
    from grafimo_b200.engine import Context, Scan
    from oracle import oracle as orc
    free, total = torch.cuda.mem_get_info()
    if free < 20e9:
        pytest.skip("needs ~10 GB of free device memory")
    ctx = Context(0)
    m = gu.load_motif("synth_w25_meme__bgnt")
    w = m["width"]
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    n = 30_000_000
    g = torch.Generator(device="cuda"); g.manual_seed(55)
    packed = torch.randint(0, 1 << 62, (n,), dtype=torch.int64, device="cuda", generator=g) & ((1 << (2 * w)) - 1)
    sc = Scan(ctx, dm, strands=2, threshold=1.0, dense_rows=n)
    sc.score(packed)
    kept = sc.finalize_device()
    hist = sc.histogram()
    o = sc.out
    with torch.cuda.stream(ctx.stream):
        ptab = torch.from_numpy(np.concatenate([dm.ptable, [1.0]])).to(ctx.device)
        reported_bins = ptab < 1.0
        assert kept == int(hist[reported_bins].sum().item())  # every window with p < 1, nothing else
        row, strand, isc, p, q, score = (o[k][:kept] for k in ("row", "strand", "iscore", "p", "q", "score"))
        bins = (isc - int(dm.lo)).to(torch.int64)
        assert torch.equal(torch.bincount(bins, minlength=dm.span + 1)[reported_bins], hist[reported_bins])
        assert torch.equal(p, ptab[bins]) and torch.equal(q, sc.qtab[bins])
        rank = sc.rank.to(torch.int64)[bins]
        key = row * 2 + strand.to(torch.int64)
        d_rank, d_key = rank[1:] - rank[:-1], key[1:] - key[:-1]
        assert bool((d_rank >= 0).all()) and bool(((d_rank > 0) | (d_key > 0)).all())  # (p-rank, row, strand), strictly: no duplicates
        assert bool((p[1:] >= p[:-1]).all()) and bool((q >= p).all())
        sel = torch.arange(0, kept, 600_011, device=ctx.device)
        rows_s, strand_s = row[sel].cpu().numpy(), strand[sel].cpu().numpy()
        isc_s, p_s, score_s = isc[sel].cpu().numpy(), p[sel].cpu().numpy(), score[sel].cpu().numpy()
        xs = packed[torch.from_numpy(rows_s).to(ctx.device)].cpu().numpy()
    seqs = ["".join("ACGT"[(int(v) >> (2 * i)) & 3] for i in range(w)) for v in xs]
    comp = str.maketrans("ACGT", "TGCA")
    seqs = [s.translate(comp)[::-1] if st else s for s, st in zip(seqs, strand_s)]
    e_isc, e_lo, e_p = orc.score_rows(orc.kmers_to_matrix(seqs, w), m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    assert np.array_equal(e_isc, isc_s) and np.array_equal(e_p, p_s) and np.array_equal(e_lo, score_s)
    ctx.close()
