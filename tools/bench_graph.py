#!/usr/bin/env python
"""K7 benchmark: k-mer extraction from a synthetic variation graph (C2 parity form: 1 Mb region, 2,504 haplotypes,
1000G-like SNP/indel density) and the text-free path graph -> K7 -> K2/K5/K6 -> report table.

    python tools/bench_graph.py [--region-len 1000000] [--haplotypes 2504] [--width 19] [--reps 5]

Prints one JSON line.  Timings are wall clock around stream-synchronised calls (gb2_graph_prepare synchronises itself).
"""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--region-len", type=int, default=1_000_000)
    ap.add_argument("--haplotypes", type=int, default=2504)
    ap.add_argument("--width", type=int, default=19)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--regions", type=int, default=1, help="split the region into this many BED-like regions")
    ap.add_argument("--indel-frac", type=float, default=0.1)
    ap.add_argument("--density", type=float, default=1.0 / 40.0)
    ap.add_argument("--dense", action="store_true", help="also time the unthresholded report (-t 1): K8 device writer vs DataFrame + host writers")
    a = ap.parse_args()
    import torch
    from grafimo_b200 import engine, synth
    from grafimo_b200 import score_sequences as ss
    from grafimo_b200.motif_ops import build_motif_meme
    from grafimo_b200.vgraph import VariationGraph

    ctx = engine.Context(0)
    ss._ctx = ctx
    t = time.perf_counter()
    ref, variants, gt = synth.variant_set(a.region_len, a.haplotypes, 20240, density=a.density, indel_frac=a.indel_frac)
    t_gen = time.perf_counter() - t
    t = time.perf_counter()
    g = VariationGraph.build("1", ref, variants, gt)
    t_build = time.perf_counter() - t
    t = time.perf_counter()
    dg = g.to_device(ctx)
    ctx.sync()
    t_upload = time.perf_counter() - t
    from grafimo_b200.extract_regions import DeviceGraph
    t = time.perf_counter()
    dn = DeviceGraph.build(ctx, "1", ref, variants, gt=gt)
    ctx.sync()
    t_native = time.perf_counter() - t
    L = a.region_len
    step = (L + a.regions - 1) // a.regions
    regions = [(lo, min(L, lo + step + a.width - 1)) for lo in range(0, L, step)]
    rows = dg.extract(regions, a.width)  # warm-up
    ctx.sync()
    tp, te = [], []
    lib = ctx.lib
    for _ in range(a.reps):
        t0 = time.perf_counter()
        rows = dg.extract(regions, a.width)
        ctx.sync()
        te.append(time.perf_counter() - t0)
    n = rows.n
    # invariant: the frequencies of the walks that share a first base add up to the haplotypes through that base
    rw = dg.extract(regions[:1], a.width, want_walks=True)
    with torch.cuda.stream(ctx.stream):
        first = rw.walk.view(-1, 32)[:rw.n, 0].to(torch.int64) * 64 + rw.walk_off[:rw.n].to(torch.int64)
        uniq, inv = torch.unique_consecutive(first, return_inverse=True)
        sums = torch.zeros(uniq.shape[0], dtype=torch.int64, device=ctx.device).index_add_(0, inv, rw.freq[:rw.n].to(torch.int64))
        node = (uniq // 64).cpu().numpy()
        sums = sums.cpu().numpy()
    ctx.sync()
    through = np.array([a.haplotypes if c == 0xFFFFFFFF else int(np.unpackbits(g.cons_bits[c].view(np.uint8)).sum())
                        for c in g.node_cons[node]])
    # walks that run off the end of the chromosome do not exist: skip first bases in the last w positions
    inner = g.node_a0[node] < regions[0][1] - 64
    invariant_ok = bool(np.array_equal(sums[inner], through[inner]))

    tmp = tempfile.mkdtemp(prefix="gb2_graph_")
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "fixtures.json")))
    mp = os.path.join(tmp, "ctcf.meme")
    open(mp, "w").write(fx["ctcf_meme"])
    with contextlib.redirect_stdout(io.StringIO()):
        motif = build_motif_meme(mp, "unfrm_dst", 0.1, False, 1, False, True)[0]

    class Args:
        cores, threshold, noqvalue, qvalueT, noreverse, recomb, verbose = 1, 1e-4, False, False, False, False, False
    tt = []
    for _ in range(max(2, a.reps // 2 + 1)):
        t0 = time.perf_counter()
        r2 = dg.extract(regions, a.width)
        with contextlib.redirect_stdout(io.StringIO()):
            df = ss.compute_results_rows(motif, r2, True, Args) if a.width == 19 else None
        tt.append(time.perf_counter() - t0)
    dense = None
    if a.dense and a.width == 19:
        from grafimo_b200.res_writer import write_results, write_results_device

        class Dense:
            cores, threshold, noqvalue, qvalueT, noreverse, recomb, verbose, text_only, top_graphs = 1, 1.0, False, False, False, True, False, True, 0
            outdir = os.path.join(tmp, "dense_dev")
        r2 = dg.extract(regions, a.width)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            rep = ss.scan_rows_device(motif, r2, True, Dense)
        ctx.sync()
        t_scan = time.perf_counter() - t0
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            write_results_device(rep, motif, 1, Dense, True)
        t_write = time.perf_counter() - t0
        sizes = [os.path.getsize(os.path.join(Dense.outdir, f)) for f in ("grafimo_out.tsv", "grafimo_out.gff")]
        t0 = time.perf_counter()
        tsv = rep.render(0)
        t_render_tsv = time.perf_counter() - t0
        t0 = time.perf_counter()
        gff = rep.render(1)
        t_render_gff = time.perf_counter() - t0
        Dense.outdir = os.path.join(tmp, "dense_host")
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            dfd = ss.compute_results_rows(motif, r2, True, Dense)
        t_df = time.perf_counter() - t0
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            write_results(dfd, motif, 1, Dense, True)
        t_hostw = time.perf_counter() - t0
        dense = {"rows_reported": rep.n, "scan_to_device_report_s": t_scan, "device_writer_tsv_gff_files_s": t_write,
                 "render_tsv_s": t_render_tsv, "render_gff_s": t_render_gff, "tsv_bytes": sizes[0], "gff_bytes": sizes[1],
                 "dataframe_path_s": t_df, "host_writers_tsv_gff_s": t_hostw}
        del tsv, gff, dfd
    # CPU beside it: the oracle (oracle/graph_oracle.py, pure Python: every haplotype spelled out) on a bounded sample
    from oracle import graph_oracle as go
    samp = 3000
    sub = [(p, r, al) for (p, r, al) in variants if p + len(r) < samp - 10]
    t0 = time.perf_counter()
    og = go.build_graph(ref[:samp], sub)
    orows = go.extract_rows(og, gt[:len(sub)].tolist(), (0, samp), a.width)
    t_or = time.perf_counter() - t0
    out = {
        "cpu_baseline": {"kind": "port", "what": f"oracle/graph_oracle.py (pure Python, 1 core) on the first {samp} bp x "
                         f"{a.haplotypes} haplotypes: graph + walks + explicit haplotype counting", "rows": len(orows),
                         "seconds": t_or, "rows_per_s": len(orows) / t_or,
                         "context": "the reference runs `vg find` here; the paper's supplementary benchmark gives 1246.6 s (1 thread) "
                                    "to 212.5 s (16 threads) for GRAFIMO incl. vg on 1 Mbp of regions with 2,548 individuals "
                                    "(docs/paper_results/time-mem_benchmark, other hardware)"},
        "workload": f"synthetic {L} bp region, {a.haplotypes} haplotypes, {len(variants)} variants "
                    f"({a.indel_frac:.0%} indels), width {a.width}, {len(regions)} region(s)",
        "graph": {"nodes": g.n_nodes, "edges": g.n_edges, "haplotype_set_rows": g.n_cons,
                  "haplotype_set_mb": g.cons_bits.nbytes / 1e6, "gen_s": t_gen, "build_numpy_s": t_build, "upload_s": t_upload,
                  "build_native_incl_upload_s": t_native},
        "kmer_rows": n, "rows_with_freq0": int((rows.freq[:n] == 0).sum().item()),
        "extract_ms_best": min(te) * 1e3, "extract_ms_median": float(np.median(te)) * 1e3,
        "rows_per_s": n / min(te), "scored_windows_equiv_per_s": 2 * n / min(te),
        "freq_sum_invariant_ok": invariant_ok,
        "graph_to_table_s_best": min(tt), "hits": None if df is None else int(len(df)),
        "launches_total": ctx.launches, "dense_report": dense,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
