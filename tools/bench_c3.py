"""BASELINE config 3 on 1..8 GPUs: a JASPAR-CORE-sized motif collection (800 synthetic motifs, widths 6..30) scanned over a
chr22-sized k-mer set per width (7.1e7 k-mers each, both strands, p < 1e-4, q-values on), the collection SHARDED BY MOTIF
over the ranks (what the reference's per-motif loop / mp.Pool over motifs becomes: src/grafimo/grafimo.py:177-183,
src/grafimo/motif_ops.py:303-335).

    python tools/bench_c3.py [--gpus N] [--motifs 800] [--out profiles/r02_c3_<N>gpu.json]

Every rank: host PWM maths of ITS motifs -> batched DP (K3) -> batched upload + K4 (gb2_motif_create_batched) -> ManyScan
(K2 per motif into one hit buffer, ONE K5 launch, ONE sort).  No collective on the data path: a motif is scanned wholly by
one rank, so its q-values are global as they are.  Parity inside the run: the per-motif hit tables of all ranks are gathered
(gb2_allgather_bytes) and rank 0 compares them, bit for bit, with its own single-GPU scan of the WHOLE collection.
Times are device-synchronised wall times, max over ranks (gb2_allreduce_max_f64)."""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def motif_cost(w):
    """relative K2 cost of a width-w motif: lookups (4-base chunks up to 24 bp, 3-base chunks above) + the two histogram updates"""
    return (-(-w // 4) if w < 25 else -(-w // 3)) + 8.4


def prepare(ctx, raw):
    """host scaling + K3 + batched upload/K4 for a list of parsed motifs -> (device motifs, seconds per phase)"""
    from grafimo_b200 import motif_ops as mo
    from grafimo_b200.score_sequences import device_motifs
    t0 = time.perf_counter()
    todo = [m for m in raw if not m.is_scaled]  # rank 0 prepares the rest of the collection later for the parity run
    for m in todo:
        mo._scale_motif(m, True)
    t1 = time.perf_counter()
    pvs = ctx.pval_dp_batched([m.score_matrix_acgt() for m in todo], [m.bg_acgt() for m in todo])
    for m, pv in zip(todo, pvs):
        m.set_motif_pval_matrix(pv)
    t2 = time.perf_counter()
    dms = device_motifs(raw, ctx)
    ctx.sync()
    t3 = time.perf_counter()
    return dms, dict(scale_s=t1 - t0, dp_s=t2 - t1, upload_k4_s=t3 - t2)


def scan(ctx, raw, dms, sets, cap):
    from grafimo_b200.engine import ManyScan
    many = ManyScan(ctx, dms, strands=2, threshold=1e-4, hit_capacity=cap)
    for k, m in enumerate(raw):
        many.score(k, sets[m.width])
    many.qvalues()
    kept = many.finalize_device()
    return many, kept


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--motifs", type=int, default=800)
    ap.add_argument("--kmers", type=int, default=int(5.08e7 * 1.4))
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    if args.gpus > 1 and "RANK" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 400), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    import torch
    from grafimo_b200 import dist as gdist
    from grafimo_b200 import engine, synth
    from grafimo_b200 import motif_ops as mo
    info = gdist.init_from_env("nccl")
    rank, world = info["rank"], info["world"]
    torch.cuda.set_device(info["local"])
    ctx = engine.Context(info["local"])
    gdist.init_comm(ctx)
    tmp = tempfile.mkdtemp(prefix="gb2_c3_")
    text, widths = synth.synthetic_meme_collection(args.motifs, 20242)
    path = os.path.join(tmp, "jaspar_like.meme")
    open(path, "w").write(text)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        raw_all = mo._read_meme(path, "unfrm_dst", 0.1, False, False, True)
    t_parse = time.perf_counter() - t0
    mine = gdist.assign_chromosomes([motif_cost(int(w)) for w in widths], world)[rank]
    raw = [raw_all[i] for i in mine]
    # one k-mer set per distinct width, the same on every rank (seeded)
    g = torch.Generator(device="cuda")
    sets = {}
    for w in sorted(set(int(x) for x in widths)):
        g.manual_seed(1000 + w)
        sets[w] = torch.randint(0, 1 << (2 * w), (args.kmers,), dtype=torch.int64, device="cuda", generator=g)
    torch.cuda.synchronize()
    dms, prep = prepare(ctx, raw)
    cap = 1 << 24
    best = None
    for rep in range(args.reps):
        if world > 1:
            torch.distributed.barrier()
        ctx.sync()
        t0 = time.perf_counter()
        many, kept = scan(ctx, raw, dms, sets, cap)
        ctx.sync()
        dt = time.perf_counter() - t0
        dt = ctx.allreduce_max([dt])[0]
        best = dt if best is None else min(best, dt)
    windows = 2 * args.kmers * args.motifs
    # ---- gather the per-motif tables: (global motif index, row, strand, int score, p, q), padded to the largest rank
    o = many.out
    gid = torch.tensor(mine, dtype=torch.int64, device=ctx.device)[o["motif"][:kept].to(torch.int64)]
    counts = ctx.allgather(torch.tensor([kept], dtype=torch.int64, device=ctx.device))
    ctx.sync()
    counts = counts.view(-1).cpu().tolist()
    cap_rows = max(max(counts), 1)
    cols = {"motif": gid, "row": o["row"][:kept], "strand": o["strand"][:kept], "iscore": o["iscore"][:kept], "p": o["p"][:kept], "q": o["q"][:kept]}
    gathered = {}
    for k, v in cols.items():
        buf = torch.zeros(cap_rows, dtype=v.dtype, device=ctx.device)
        buf[:kept] = v
        gathered[k] = ctx.allgather(buf)
    ctx.sync()
    line = None
    if rank == 0:
        parts = {k: torch.cat([gathered[k][r, :counts[r]] for r in range(world)]).cpu().numpy() for k in cols}
        order = np.lexsort((parts["strand"], parts["row"], -parts["iscore"].astype(np.int64), parts["p"], parts["motif"]))  # ManyScan order: (motif, p-rank, row, strand)
        parts = {k: v[order] for k, v in parts.items()}
        parity = None
        t_single = None
        if world > 1:  # the whole collection on this one GPU: the table the N-GPU run must reproduce
            dms_all, prep_all = prepare(ctx, raw_all)
            ctx.sync()
            t0 = time.perf_counter()
            whole, kept_all = scan(ctx, raw_all, dms_all, sets, 1 << 24)
            ctx.sync()
            t_single = time.perf_counter() - t0
            w = whole.out
            exp = {"motif": w["motif"][:kept_all].to(torch.int64).cpu().numpy(), "row": w["row"][:kept_all].cpu().numpy(),
                   "strand": w["strand"][:kept_all].cpu().numpy(), "iscore": w["iscore"][:kept_all].cpu().numpy(),
                   "p": w["p"][:kept_all].cpu().numpy(), "q": w["q"][:kept_all].cpu().numpy()}
            parity = {k: bool(np.array_equal(parts[k], exp[k])) for k in exp}
            parity["ok"] = all(parity.values())
            parity["hits"] = int(kept_all)
        line = {"config": "C3: JASPAR-sized collection on a chr22-sized k-mer set per width, sharded by motif", "n_gpus": world,
                "motifs": args.motifs, "kmers_per_width": args.kmers, "distinct_widths": len(sets), "scored_windows": windows,
                "scan_s": best, "windows_per_s": windows / best, "hits": int(sum(counts)), "hits_per_rank": counts,
                "motifs_per_rank": [len(x) for x in gdist.assign_chromosomes([motif_cost(int(w)) for w in widths], world)],
                "host_parse_s": t_parse, "prep_this_rank": prep, "single_gpu_scan_s_same_run": t_single,
                "speedup_vs_single_gpu_same_run": (t_single / best) if t_single else None, "parity_vs_single_gpu": parity,
                "timing": "wall time around ManyScan (score all motifs + one K5 launch + one sort), stream-synchronised, best of "
                          f"{args.reps}, max over ranks (gb2_allreduce_max_f64)"}
        print(json.dumps(line))
        if args.out:
            with open(os.path.join(ROOT, args.out) if not os.path.isabs(args.out) else args.out, "w") as fh:
                json.dump(line, fh, indent=1)
    if world > 1:
        torch.distributed.barrier()
        ctx.close()
        torch.distributed.destroy_process_group()
    if rank == 0 and line and line["parity_vs_single_gpu"] and not line["parity_vs_single_gpu"]["ok"]:
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
