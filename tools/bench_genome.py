#!/usr/bin/env python
"""Graph path at chromosome scale, one chromosome (or several) per GPU -- the C3/C4 shape of BASELINE.json
("synthetic chr22-sized graph", "sharded by chromosome across 2/4/8 B200 with global q-value").

    python tools/bench_genome.py [--chrom-len 50818468] [--chroms-per-gpu 1] [--haplotypes 5008] [--width 19]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_genome.py ...

Every rank: synthetic chromosome (reference, ~1/40 bp variants, 10 % indels, genotype bit sets drawn on the GPU) ->
gb2_graph_build -> K7 extraction of every haplotype-aware k-mer -> K2 (both strands) -> NCCL all-reduce of the score
histogram -> K5/K6 -> merged report table on every rank.  Rank 0 prints one JSON line; times are max over ranks.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chrom-len", type=int, default=50_818_468)
    ap.add_argument("--chroms-per-gpu", type=int, default=1)
    ap.add_argument("--total-chroms", type=int, default=0,
                    help="strong scaling: a FIXED genome of this many chromosomes dealt round-robin over the ranks (overrides --chroms-per-gpu)")
    ap.add_argument("--out", default="", help="also write the JSON line to this file")
    ap.add_argument("--haplotypes", type=int, default=5008)
    ap.add_argument("--width", type=int, default=19)
    ap.add_argument("--threshold", type=float, default=1e-4)
    ap.add_argument("--build-threads", type=int, default=0, help="host threads of the graph builder (0 = cores / ranks)")
    a = ap.parse_args()
    import torch
    from grafimo_b200 import dist as gdist
    from grafimo_b200 import score_sequences as ss
    from grafimo_b200 import synth
    from grafimo_b200.extract_regions import DeviceGraph
    from grafimo_b200.motif_ops import build_motif_meme

    info = gdist.init_from_env("nccl")
    rank, world, local = info["rank"], info["world"], info["local"]
    torch.cuda.set_device(local)
    ctx = ss._context()
    gdist.init_comm(ctx)  # histogram all-reduce + hit-column all-gather run inside the C ABI (csrc/comm.cu)
    tmp = tempfile.mkdtemp(prefix="gb2_genome_")
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "fixtures.json")))
    open(os.path.join(tmp, "ctcf.meme"), "w").write(fx["ctcf_meme"])
    with contextlib.redirect_stdout(io.StringIO()):
        motif = build_motif_meme(os.path.join(tmp, "ctcf.meme"), "unfrm_dst", 0.1, False, 1, False, True)[0]
    assert motif.width == a.width or a.width == 19

    class Args:
        cores, threshold, noqvalue, qvalueT, noreverse, recomb, verbose = 1, a.threshold, False, False, False, False, False

    t = dict(gen=0.0, build=0.0, extract=0.0)
    rows, n_var, n_nodes, set_mb = [], 0, 0, 0.0
    items = []
    mine = list(range(rank, a.total_chroms, world)) if a.total_chroms else [rank * a.chroms_per_gpu + c for c in range(a.chroms_per_gpu)]
    for idx in mine:
        t0 = time.perf_counter()
        ref, variants, gtb = synth.variant_arrays(a.chrom_len, a.haplotypes, 5000 + idx, device=ctx.device)
        t["gen"] += time.perf_counter() - t0
        items.append((str(idx + 1), ref, variants, None, gtb))
        n_var += len(variants["pos"])
    # host passes of all chromosomes of this rank on worker threads of the library, uploads as they finish
    t0 = time.perf_counter()
    threads = a.build_threads or max(1, (os.cpu_count() or 1) // world)
    graphs = DeviceGraph.build_many(ctx, items, n_threads=threads)
    ctx.sync()
    t["build"] = time.perf_counter() - t0
    del items
    for dg in graphs:
        n_nodes += int(dg.info.n_nodes); set_mb += dg.info.n_sets * dg.info.words * 4 / 1e6
        t0 = time.perf_counter()
        rows.append(dg.extract([(0, a.chrom_len)], motif.width))
        ctx.sync()
        t["extract"] += time.perf_counter() - t0
    if world > 1:
        torch.distributed.barrier()
    t_calls = []
    for rep in range(3):  # first call: allocations of the process (scratch, hit buffers); then steady state; then the phase timers
        if rep == 2:
            os.environ["GB2_PHASES"] = "1"
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            df = ss.compute_results_rows(motif, rows, True, Args)
        t_calls.append(time.perf_counter() - t0)
    os.environ.pop("GB2_PHASES", None)
    phases = {k: round(v, 5) for k, v in ss.LAST_PHASES.items()}
    t_first = gdist.allreduce_max(t_calls[0], device=ctx.device)
    t_table = t_calls[1]
    n_rows = sum(r.n for r in rows)
    tot_rows = gdist.allreduce_sum(n_rows, device=ctx.device)
    out = {k: gdist.allreduce_max(v, device=ctx.device) for k, v in t.items()}
    t_table = gdist.allreduce_max(t_table, device=ctx.device)
    if rank == 0:
        n_chroms = a.total_chroms or a.chroms_per_gpu * world
        L, H = a.chrom_len * n_chroms, a.haplotypes
        line = json.dumps({
            "workload": f"{n_chroms} synthetic chromosome(s) of {a.chrom_len} bp over {world} GPU(s) ({'fixed genome: strong scaling' if a.total_chroms else 'per-GPU load fixed: weak scaling'}), {H} haplotypes, "
                        f"~1/40 bp variants (10 % indels), CTCF w=19, both strands, p<{a.threshold:g}, global q-values",
            "n_gpus": world, "genome_bp": L, "variants_rank0": n_var, "nodes_rank0": n_nodes, "haplotype_sets_mb_rank0": set_mb,
            "kmer_rows_total": tot_rows, "windows_scored_total": 2 * tot_rows,
            "haplotype_windows_equivalent": 2 * L * H,
            "synth_gen_s": out["gen"], "graph_build_s": out["build"], "graph_build_threads": threads, "extract_s": out["extract"], "score_to_table_s": t_table,
            "score_to_table_first_call_s": t_first, "score_to_table_phases_rank0_s": phases,
            "extract_rows_per_s": tot_rows / out["extract"], "scan_s": out["extract"] + t_table,
            "haplotype_windows_equivalent_per_s": 2 * L * H / (out["extract"] + t_table), "hits": int(len(df)),
            "table_merge": "fixed-width hit columns all-gathered on the device (gb2_allgather_bytes), strings decoded once"})
        print(line)
        if a.out:
            with open(a.out, "w") as fh:
                fh.write(line + "\n")
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
