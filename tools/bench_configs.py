"""Measurements for the other BASELINE.json configurations (not bench lines; numbers go to DESIGN.md):
C3 -- JASPAR-sized motif collection: host PWM maths + batched DP (K3) + p-tables (K4) + scan of every motif;
C5 -- long motifs, both strands, no threshold (every window reported);   K1 -- encoder throughput."""
import contextlib, io, os, sys, tempfile, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu
from grafimo_b200 import synth, motif_ops as mo
from grafimo_b200.engine import Context, Scan
from grafimo_b200.score_sequences import device_motif, _context

ctx = _context()
tmp = tempfile.mkdtemp(prefix="gb2_cfg_")

def ev():
    return torch.cuda.Event(enable_timing=True)

ONLY = os.environ.get("GB2_ONLY", "")
RESULTS = {}  # section -> list of records; written as JSON to $GB2_JSON (numbers for profiles/, not prose)
import json, atexit
def _dump():
    if os.environ.get("GB2_JSON") and RESULTS:
        with open(os.environ["GB2_JSON"], "w") as fh:
            json.dump(RESULTS, fh, indent=1)
atexit.register(_dump)
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
g = torch.Generator(device="cuda"); g.manual_seed(3)
if ONLY == "k1":
    print("K1 encoder")
    for w in (8, 19, 30):
        n1 = 1 << 27
        a = synth.windows_to_ascii(torch.randint(0, 1 << (2 * w), (1 << 20,), dtype=torch.int64, device="cuda"), w).repeat(n1 >> 20, 1)
        for rep in range(3):
            e0, e1 = ev(), ev()
            e0.record(ctx.stream)
            packed, nmask, counts = ctx.encode(a)
            e1.record(ctx.stream); ctx.sync()
        ms = e0.elapsed_time(e1)
        print(f"    w={w}: {n1 / ms / 1e6:.1f} G k-mers/s, {(n1 * (w + 8)) / ms / 1e6:.0f} GB/s (read {w} + write 8 B per k-mer)")
        del a, packed
    sys.exit(0)
# ---------------------------------------------------------------- C3
if ONLY in ("", "c3"):
    text, widths = synth.synthetic_meme_collection(800, 20242)
    path = os.path.join(tmp, "jaspar_like.meme"); open(path, "w").write(text)
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        raw = mo._read_meme(path, "unfrm_dst", 0.1, False, False, True)
    t_parse = time.time() - t0
    t0 = time.time()
    for m in raw:
        mo._scale_motif(m, True)
    t_scale = time.time() - t0
    for rep in range(2):
        t0 = time.time()
        pvs = ctx.pval_dp_batched([m.score_matrix_acgt() for m in raw], [m.bg_acgt() for m in raw])
        t_dp = time.time() - t0
    for m, pv in zip(raw, pvs):
        m.set_motif_pval_matrix(pv)
    spans = [int(np.count_nonzero(pv)) for pv in pvs]
    work = sum(4 * m.width * s for m, s in zip(raw, spans))
    print(f"C3: 800 motifs (widths {widths.min()}..{widths.max()}, mean {widths.mean():.1f}): parse {t_parse:.2f}s, log-odds+scaling {t_scale:.2f}s, "
          f"batched DP K3 incl. H2D/D2H {t_dp * 1e3:.1f} ms ({800 / t_dp:.0f} motifs/s, {work / t_dp / 1e9:.2f} G mul-add/s)")
    from oracle import oracle as orc
    t0 = time.time()
    for m in raw[:40]:
        orc.pval_dp(m.score_matrix_acgt(), m.bg_acgt())
    t_cpu = (time.time() - t0) / 40 * 800
    print(f"    oracle C DP, 1 thread, extrapolated to 800 motifs: {t_cpu:.1f}s  (the reference's Cython loop: ~0.27 s per w=19 motif)")
    t0 = time.time()
    dms = [device_motif(m, ctx) for m in raw]
    t_up = time.time() - t0
    print(f"    motif upload + K4 p-tables: {t_up:.2f}s ({t_up / 800 * 1e3:.2f} ms per motif)")
    # scan: one k-mer set per distinct width (as scan_graph extracts one TSV set per width), chr22-sized: 5.08e7 positions x 1.4
    n = int(5.08e7 * 1.4)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    sets = {}
    for w in sorted(set(widths.tolist())):
        sets[w] = torch.randint(0, 1 << (2 * w), (n,), dtype=torch.int64, device="cuda", generator=g)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    tot_hits = 0
    scans = [Scan(ctx, dm, strands=2, threshold=1e-4, hit_capacity=1 << 18) for dm in dms[:8]]
    e0.record(ctx.stream)
    t0 = time.time()
    for i, (m, dm) in enumerate(zip(raw, dms)):
        sc = scans[i % 8] if False else Scan(ctx, dm, strands=2, threshold=1e-4, hit_capacity=1 << 18)
        sc.score(sets[m.width])
        sc.qvalues()
        tot_hits += sc.finalize_device()
    e1.record(ctx.stream); ctx.sync()
    dt = time.time() - t0
    print(f"    scan of all 800 motifs over {n} k-mers each (both strands, q-values): {dt:.2f}s wall, {e0.elapsed_time(e1) / 1e3:.2f}s device "
          f"= {800 * 2 * n / dt / 1e9:.1f} G windows/s, {tot_hits} hits")
    # the same work with the host out of the way: K2 + K5 of every motif are queued first (no host read in between), the
    # hit tables are finalized afterwards -- what a many-motif driver does instead of the reference's one-motif-at-a-time loop
    for rep in range(2):
        e0, e1 = ev(), ev()
        e0.record(ctx.stream)
        t0 = time.time()
        scans = []
        for m, dm in zip(raw, dms):
            sc = Scan(ctx, dm, strands=2, threshold=1e-4, hit_capacity=1 << 17)
            sc.score(sets[m.width])
            sc.qvalues()
            scans.append(sc)
        tot2 = sum(sc.finalize_device() for sc in scans)
        e1.record(ctx.stream); ctx.sync()
        dt2 = time.time() - t0
        del scans
    assert tot2 == tot_hits
    print(f"    queued form (score + BH of all motifs first, then the hit tables): {dt2:.2f}s wall, {e0.elapsed_time(e1) / 1e3:.2f}s device "
          f"= {800 * 2 * n / dt2 / 1e9:.1f} G windows/s")
    # one sort for all motifs: shared hit buffer, gb2_finalize_hits_many (engine.ManyScan)
    from grafimo_b200.engine import ManyScan
    for rep in range(2):
        e0, e1 = ev(), ev()
        e0.record(ctx.stream)
        t0 = time.time()
        many = ManyScan(ctx, dms, strands=2, threshold=1e-4, hit_capacity=1 << 24)
        for k, m in enumerate(raw):
            many.score(k, sets[m.width])
        many.qvalues()
        tot3 = many.finalize_device()
        e1.record(ctx.stream); ctx.sync()
        dt3 = time.time() - t0
        del many
    assert tot3 == tot_hits, (tot3, tot_hits)
    print(f"    ManyScan (shared hit buffer, one sort for all motifs): {dt3:.2f}s wall, {e0.elapsed_time(e1) / 1e3:.2f}s device "
          f"= {800 * 2 * n / dt3 / 1e9:.1f} G windows/s")
    del sets

# ---------------------------------------------------------------- C5
if ONLY in ("", "c5"):
    print("C5: long motifs, both strands, threshold 1 (every window with p<1 is a hit)")
    n5 = 30_000_000
    for tag in ("synth_w25_meme__bgnt", "synth_w30_meme__bgnt"):
        m = gu.load_motif(tag)
        dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
        w = m["width"]
        packed = torch.randint(0, 1 << 62, (n5,), dtype=torch.int64, device="cuda", generator=g) & ((1 << (2 * w)) - 1)
        modes = [("hit records + 41-bit keys", {}, False),
                 ("dense scores + library sort (GB2_DENSE_CUB=1, round 1)", {"GB2_DENSE_CUB": "1"}, False),
                 ("dense scores + gb2_finalize_dense (two partition passes)", {}, False),
                 ("dense scores + gb2_finalize_dense, row / strand / integer score only (what the device report writer reads)", {}, True)]
        for mode, env, index_only in modes:
            for k in ("GB2_DENSE_CUB",):
                os.environ.pop(k, None)
            os.environ.update(env)
            for rep in range(3):
                sc = Scan(ctx, dm, strands=2, threshold=1.0, hit_capacity=2 * n5, dense_rows=n5 if mode.startswith("dense") else 0)
                e0, e1, e2 = ev(), ev(), ev()
                e0.record(ctx.stream)
                sc.score(packed)
                e1.record(ctx.stream)
                sc.qvalues()
                kept = sc.finalize_device(index_only=index_only)
                e2.record(ctx.stream); ctx.sync()
            print(f"    {tag}: w={w} span={dm.span} R={dm.info.lut_replicas} [{mode}]: K2 {e0.elapsed_time(e1):.2f} ms, K5+K6 {e1.elapsed_time(e2):.2f} ms, "
                  f"{kept} rows reported of {2 * n5}; {2 * n5 / (e0.elapsed_time(e2) * 1e-3) / 1e9:.2f} G windows/s")
            RESULTS.setdefault("c5", []).append(dict(motif=tag, width=w, span=int(dm.span), chunk_bases=int(dm.info.chunk_bases), n_chunks=int(dm.info.n_chunks),
                replicas=int(dm.info.lut_replicas), mode=mode, kmers=n5, rows_reported=int(kept), k2_ms=e0.elapsed_time(e1), k5_k6_ms=e1.elapsed_time(e2),
                step_ms=e0.elapsed_time(e2), windows_per_s=2 * n5 / (e0.elapsed_time(e2) * 1e-3)))
            del sc
        del packed

# ---------------------------------------------------------------- mid-range thresholds: hit records or dense scores?
if ONLY in ("", "midrange"):
    print("mid-range thresholds (CTCF, 2^26 k-mers, both strands): hit-record form against the dense form")
    nm = 1 << 26
    m = gu.load_motif("ctcf_meme__unif")
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    packed = torch.randint(0, 1 << 62, (nm,), dtype=torch.int64, device="cuda", generator=g) & ((1 << 38) - 1)
    for thr in (1e-3, 3e-3, 0.01, 0.03, 0.1, 0.25):
        for dense in (False, True):
            for rep in range(3):
                sc = Scan(ctx, dm, strands=2, threshold=thr, hit_capacity=int(2 * nm * min(1.0, thr * 1.3)) + (1 << 16), dense_rows=nm if dense else 0)
                e0, e1, e2 = ev(), ev(), ev()
                e0.record(ctx.stream)
                sc.score(packed)
                e1.record(ctx.stream)
                sc.qvalues()
                kept = sc.finalize_device()
                e2.record(ctx.stream); ctx.sync()
            print(f"    t={thr:g} [{'dense' if dense else 'hits '}]: K2 {e0.elapsed_time(e1):.2f} ms, K5+K6 {e1.elapsed_time(e2):.2f} ms, step {e0.elapsed_time(e2):.2f} ms, {kept} rows")
            RESULTS.setdefault("midrange", []).append(dict(threshold=thr, form="dense" if dense else "hit records", kmers=nm, rows_reported=int(kept),
                k2_ms=e0.elapsed_time(e1), k5_k6_ms=e1.elapsed_time(e2), step_ms=e0.elapsed_time(e2)))
            del sc
    del packed

# ---------------------------------------------------------------- wide motifs (two packed words per k-mer)
if ONLY in ("", "wide"):
    print("wide motifs (33..64 bp): K2 wide kernel, thresholded, both strands, q-values on")
    nw = 1 << 26
    for tag in ("synth_w35_meme__bgnt", "synth_w48_meme__bgnt", "synth_w64_meme__bgnt"):
        m = gu.load_motif(tag)
        dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
        w = m["width"]
        packed = torch.randint(0, 1 << 62, (nw, 2), dtype=torch.int64, device="cuda", generator=g)
        packed[:, 1] &= (1 << (2 * (w - 32))) - 1
        for rep in range(3):
            sc = Scan(ctx, dm, strands=2, threshold=1e-4, hit_capacity=1 << 20)
            e0, e1, e2 = ev(), ev(), ev()
            e0.record(ctx.stream)
            sc.score(packed)
            e1.record(ctx.stream)
            sc.qvalues()
            kept = sc.finalize_device()
            e2.record(ctx.stream); ctx.sync()
        ms = e0.elapsed_time(e1)
        print(f"    {tag}: w={w} span={dm.span} chunks={dm.info.n_chunks} R={dm.info.lut_replicas} smem={dm.info.smem_bytes}: K2 {ms:.2f} ms = "
              f"{nw / ms / 1e6:.1f} G k-mers/s, {16 * nw / ms / 1e6:.0f} GB/s; K5+K6 {e1.elapsed_time(e2):.2f} ms, {kept} hits")
        RESULTS.setdefault("wide", []).append(dict(motif=tag, width=w, span=int(dm.span), chunk_bases=int(dm.info.chunk_bases), n_chunks=int(dm.info.n_chunks),
            replicas=int(dm.info.lut_replicas), hist_global=int(dm.info.hist_global), smem_bytes=int(dm.info.smem_bytes), kmers=nw, k2_ms=ms,
            kmers_per_s=nw / (ms * 1e-3), algorithmic_GBps=16 * nw / ms / 1e6, frac_of_measured_hbm_peak=16 * nw / ms / 1e6 / PEAK, hits=int(kept),
            note="CUDA events around gb2_score on the context stream, third repetition; 16 B per k-mer (two packed words)"))
        del packed, sc

# ---------------------------------------------------------------- narrow motifs: one k-mer word, all widths' chunk plans
if ONLY in ("", "narrow"):
    print("narrow motifs (<= 32 bp): K2, thresholded, both strands, q-values on, 2^28 random k-mers")
    nn = 1 << 28
    for tag in ("synth_w8_meme__bgnt", "synth_w11_meme__bgnt", "ctcf_meme__unif", "synth_w25_meme__bgnt", "synth_w27_meme__bgnt", "synth_w30_meme__bgnt", "synth_w32_meme__bgnt"):
        m = gu.load_motif(tag)
        dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
        w = m["width"]
        packed = torch.randint(0, 1 << 62, (nn,), dtype=torch.int64, device="cuda", generator=g) & ((1 << (2 * w)) - 1)
        for rep in range(3):
            sc = Scan(ctx, dm, strands=2, threshold=1e-4, hit_capacity=1 << 22)
            e0, e1 = ev(), ev()
            e0.record(ctx.stream)
            sc.score(packed)
            e1.record(ctx.stream); ctx.sync()
        ms = e0.elapsed_time(e1)
        print(f"    {tag}: w={w} span={dm.span} cb={dm.info.chunk_bases} chunks={dm.info.n_chunks} R={dm.info.lut_replicas}: K2 {ms:.3f} ms = {8 * nn / ms / 1e6:.0f} GB/s "
              f"= {8 * nn / ms / 1e6 / PEAK:.3f} of the measured peak")
        RESULTS.setdefault("narrow", []).append(dict(motif=tag, width=w, span=int(dm.span), chunk_bases=int(dm.info.chunk_bases), n_chunks=int(dm.info.n_chunks),
            replicas=int(dm.info.lut_replicas), kmers=nn, k2_ms=ms, algorithmic_GBps=8 * nn / ms / 1e6, frac_of_measured_hbm_peak=8 * nn / ms / 1e6 / PEAK))
        del packed, sc

# ---------------------------------------------------------------- K1
if ONLY == "":
    print("K1 encoder")
    for w in (8, 19, 30):
        n1 = 1 << 27
        a = synth.windows_to_ascii(torch.randint(0, 1 << (2 * w), (1 << 20,), dtype=torch.int64, device="cuda"), w).repeat(n1 >> 20, 1)
        for rep in range(2):
            e0, e1 = ev(), ev()
            e0.record(ctx.stream)
            packed, nmask, counts = ctx.encode(a)
            e1.record(ctx.stream); ctx.sync()
        ms = e0.elapsed_time(e1)
        print(f"    w={w}: {n1 / ms / 1e6:.1f} G k-mers/s, {(n1 * (w + 8)) / ms / 1e6:.0f} GB/s (read {w} + write 8 B per k-mer)")
        del a, packed
