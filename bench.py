#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native GRAFIMO motif-scanning path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): scored windows / second, both strands, beside the HBM roofline and the reference's
CPU path.  Workload at N=1 (BASELINE.json configs[1]): CTCF MA0139.1 on a synthetic 1 Mb region with 2,504
haplotype paths -- every 19-bp window of every haplotype as a packed uint64, 2.50e9 k-mers = 20.0 GB resident
in HBM (inputs >> 126 MB L2, so no L2 flush is needed between steps).  For N>1 every rank holds its own region
of the same size (weak scaling, as a whole-genome scan shards by chromosome) and the only cross-GPU traffic is
the all-reduce of the score histogram that makes the q-values global.

A step = one pass of the hot path over the resident batch: K2 (score both strands + histogram + hit
compaction) -> [all-reduce] -> K5 (Benjamini-Hochberg from the histogram) -> K6 (filter/sort/annotate hits).
`value` is device-timed (CUDA events on the launching stream, max over ranks).  `e2e` is the same work through
the host-buffer C-ABI call gb2_scan_host (ASCII k-mers in pinned host memory -> hit table in host memory;
host<->device copies inside the timed region).  `cpu_baseline` / `--impl reference` time the CPU oracle
(oracle/, a C restatement of the reference's algorithm incl. its per-row p-value sums) on the host cores.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REGION_LEN = 1_000_000
N_HAP = 2504
THRESHOLD = 1e-4
SEED = 20240
METRIC = "scored_windows_per_sec_both_strands"
UNIT = "windows/s"


def load_fixture_motif_text():
    with open(os.path.join(ROOT, "tests", "golden", "fixtures.json")) as fh:
        return json.load(fh)["ctcf_meme"]


def golden_motif_arrays():
    """CTCF arrays produced by the reference (no GPU needed) -- used by the reference arm only."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "cases", "motif_ctcf_meme__unif.npz"))
    return dict(score_matrix=z["score_matrix"], pval_mat=z["pval_mat"], min_val=int(z["min_val"]), scale=int(z["scale"]),
                offset=float(z["offset"]), width=int(z["width"]))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().split("\n") if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.strip() == "Active":
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_median": float(np.median(power)), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_sample_rows(n_kmers, seed):
    """ASCII rows of the same synthetic workload (forward k-mers + their reverse-complement rows, which is what
    `vg find -E` hands the reference), generated on the CPU."""
    import torch
    from grafimo_b200 import synth
    per = 20000 - 19 + 1
    n_hap = max(1, (n_kmers + per - 1) // per)
    packed, _ = synth.haplotype_windows(20000, n_hap, 19, seed, device="cpu")
    fwd = synth.windows_to_ascii(packed[:n_kmers], 19).numpy()
    return np.ascontiguousarray(np.concatenate([fwd, synth.revcomp_ascii(fwd)]))


def time_oracle(rows, m, threads):
    from oracle import oracle as orc
    t0 = time.perf_counter()
    orc.score_rows(rows, m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"], nthreads=threads)
    return time.perf_counter() - t0


def write_kmer_tsvs(rows, directory, n_files, width):
    """The reference's input: `vg find -K w -E`-format TSVs (7 columns) under <directory>/width_<w>/, `rows` (uint8 [n, w],
    forward k-mers followed by their reverse-complement rows) split over n_files files."""
    d = os.path.join(directory, f"width_{width}")
    os.makedirs(d, exist_ok=True)
    n = rows.shape[0]
    half = n // 2
    kmers = np.ascontiguousarray(rows).view(f"S{width}").ravel()
    per = (n + n_files - 1) // n_files
    for f in range(n_files):
        lo, hi = f * per, min(n, (f + 1) * per)
        if lo >= hi:
            break
        with open(os.path.join(d, f"part{f:03d}.tsv"), "wb") as fh:
            out = []
            for i in range(lo, hi):
                pos = i % half
                if i < half:
                    out.append(b"1:0-1000000\t%s\t1:%d+\t1:%d+\t1\tref\t1+,\n" % (kmers[i], pos, pos + width))
                else:
                    out.append(b"1:0-1000000\t%s\t1:%d-\t1:%d-\t1\tref\t1-,\n" % (kmers[i], pos + width, pos))
            fh.write(b"".join(out))
    return d


def reference_objects(cores):
    """(motif built by the reference's own build_motif_meme, Findmotif-like args, compute_results, warm-up) from oracle/_ref --
    the UNMODIFIED reference, byte-compiled by oracle/build_ref.py -- or None when it is not available."""
    from oracle import build_ref
    if not build_ref.build():
        return None
    import contextlib
    import io
    from grafimo.motif_ops import build_motif_meme
    from grafimo.score_sequences import compute_results, compute_score_seq
    from grafimo.workflow import Findmotif
    tmp = tempfile.mkdtemp(prefix="gb2_refarm_")
    path = os.path.join(tmp, "MA0139.1.meme")
    with open(path, "w") as fh:
        fh.write(load_fixture_motif_text())
    with contextlib.redirect_stdout(io.StringIO()):
        motif = build_motif_meme(path, "unfrm_dst", 0.1, False, 1, False, True)[0]
    wf = object.__new__(Findmotif)  # the attributes compute_results reads (score_sequences.py:93-99)
    wf._cores, wf._thresh, wf._no_qvalue, wf._qvalueT, wf._no_rev, wf._recomb, wf._verbose = cores, THRESHOLD, False, False, False, False, False
    # numba compiles compute_score_seq at its first call; do that once here, in the parent, so that the forked workers
    # inherit the machine code (JIT time is not part of the measured rate)
    compute_score_seq("A" * motif.width, motif.score_matrix, motif.pval_matrix, motif.min_val, motif.scale, motif.width, motif.offset)
    return motif, wf, compute_results


def time_reference(compute_results, motif, wf, directory):
    import contextlib
    import io
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        df = compute_results(motif, directory, True, wf)
    return time.perf_counter() - t0, len(df)


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores -- its compute_results
    (numba scorer, one process per core, Manager funnel, statsmodels-style BH, pandas table) from oracle/_ref when that is
    present (kind "reference"), else the C port of its algorithm (kind "port") -- on bounded samples of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as orc
    orc.build()
    m = golden_motif_arrays()
    threads = os.cpu_count() or 1
    calib = cpu_sample_rows(20000, SEED)
    t = time_oracle(calib, m, threads)
    port_rate = calib.shape[0] / t
    ref = None
    try:
        ref = reference_objects(threads)
    except Exception as e:  # a broken copy must not cost the line: fall back to the port and say so
        sys.stderr.write(f"bench.py: oracle/_ref unusable ({e!r}); timing the port\n")
    workload = "CTCF MA0139.1, synthetic 1 Mb region x 2504 haplotype paths, both strands, t=1e-4 (bounded sample per step)"
    if ref is not None:
        motif, wf, compute_results = ref
        w = motif.width
        # calibrate on 100 k rows, then size a step for about 6 s (the whole run stays within a few minutes)
        tmp = tempfile.mkdtemp(prefix="gb2_refarm_rows_")
        write_kmer_tsvs(cpu_sample_rows(50000, SEED), os.path.join(tmp, "calib"), threads, w)
        t_cal, _ = time_reference(compute_results, motif, wf, os.path.join(tmp, "calib"))
        rate = 100000 / t_cal
        per_step_s = min(8.0, max(1.0, 150.0 / max(1, args.steps + args.warmup)))
        rows_per_step = int(max(100000, min(rate * per_step_s, 6_000_000)))
        rows = cpu_sample_rows(rows_per_step // 2, SEED + 1)
        write_kmer_tsvs(rows, os.path.join(tmp, "step"), 4 * threads, w)
        for _ in range(max(args.warmup, 0)):
            time_reference(compute_results, motif, wf, os.path.join(tmp, "calib"))
        ts, hits = [], 0
        for _ in range(args.steps):
            dt1, hits = time_reference(compute_results, motif, wf, os.path.join(tmp, "step"))
            ts.append(dt1)
        dt = float(np.mean(ts))
        value = rows.shape[0] / dt
        kind = "reference"
        sample = (f"{rows.shape[0]} TSV rows/step ({rows.shape[0] // 2} forward 19-mers of the synthetic haplotype workload + their "
                  f"reverse-complement rows, {4 * threads} files): the UNMODIFIED reference's compute_results (numba scorer, {threads} "
                  f"processes = --cores {threads}, BH, DataFrame) from oracle/_ref; motif build and numba JIT excluded; {hits} rows reported")
        # the C port of the same algorithm on the same rows, once: how conservative the port is as a stand-in
        port_rate = rows.shape[0] / time_oracle(rows, m, threads)
        shutil.rmtree(tmp, ignore_errors=True)
    else:
        per_step_s = min(8.0, max(0.25, 120.0 / max(1, args.steps)))
        rows_per_step = int(max(20000, min(port_rate * per_step_s, 8_000_000)))
        rows = cpu_sample_rows(rows_per_step // 2, SEED + 1)
        for _ in range(max(args.warmup, 0)):
            time_oracle(rows[: max(2000, rows.shape[0] // 20)], m, threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            time_oracle(rows, m, threads)
        dt = (time.perf_counter() - t0) / args.steps
        value = rows.shape[0] / dt
        kind = "port"
        sample = (f"{rows.shape[0]} rows/step ({rows.shape[0] // 2} forward 19-mers of the synthetic haplotype workload + their "
                  f"reverse-complement rows), oracle C port of compute_score_seq incl. its two per-row pval_mat sums, {threads} threads "
                  "(oracle/_ref is absent: the reference itself could not be run)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i64+f64", "data": "synthetic",
        "config": {"workload": workload, "rows_per_step": int(rows.shape[0])},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "port_value_same_rows": port_rate, "port_over_reference": (port_rate / value) if kind == "reference" else None},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def build_ctcf(tmpdir):
    """CTCF through the product path: MEME parser -> log-odds -> scaling (host) -> K3 DP (GPU)."""
    from grafimo_b200.motif_ops import build_motif_meme
    path = os.path.join(tmpdir, "MA0139.1.meme")
    with open(path, "w") as fh:
        fh.write(load_fixture_motif_text())
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return build_motif_meme(path, "unfrm_dst", 0.1, False, 1, False, True)[0]


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 8 << 30


def kernel_label(info):
    """Name of the K2 instantiation a motif runs (template arguments follow from the motif: csrc/score.cu dispatch)."""
    if info.width > 32:
        return f"gb2_score_wide_kernel<{info.n_chunks},{info.lut_replicas}>"
    return f"gb2_score_kernel<{info.chunk_bases},{info.n_chunks},{info.lut_replicas},4>"


def parity_multi_gpu(ctx, dm, rank, world):
    """Hardware multi-GPU parity, run inside the bench so that the driver's N>1 launches carry correctness: a common
    (same seed on every rank) sub-shard of >= 2^21 rows is split over the ranks by rows; every rank scores its part,
    the histogram is all-reduced INSIDE the C ABI (gb2_allreduce_hist), every rank derives the q-table and finalizes its
    own hits, the fixed-width hit columns are all-gathered (gb2_allgather_bytes) and merged; rank 0 also scans the whole
    sub-shard alone and compares bit for bit: q-table of every rank, then (row, strand, integer score, p, q) of every hit.
    What it replaces / must equal: the parent-side merge + BH over all rows, score_sequences.py:171-198."""
    import torch
    from grafimo_b200 import dist as gdist
    from grafimo_b200 import engine, synth
    w = dm.width
    L, H, thr = (1 << 17) + w - 1, 16, 1e-3
    with torch.cuda.stream(ctx.stream):
        rows, _ = synth.haplotype_windows(L, H, w, SEED + 999, device=ctx.device, hap_batch=16)
    n = rows.shape[0]
    lo, hi = gdist.shard_bounds(n, rank, world)
    part = engine.Scan(ctx, dm, strands=2, threshold=thr, want_q=True, hit_capacity=1 << 20)
    part.score(rows[lo:hi], row_base=lo)
    ctx.allreduce_hist(part.histogram())
    qtab, _ = part.qvalues()
    kept = part.finalize_device()
    o = part.out
    # fixed-width columns, padded to the largest per-rank count, gathered by the library
    counts = ctx.allgather(torch.tensor([kept], dtype=torch.int64, device=ctx.device))
    ctx.sync()
    counts = counts.view(-1).cpu().tolist()
    cap = max(max(counts), 1)
    cols = {}
    with torch.cuda.stream(ctx.stream):
        for k, dt in (("row", torch.int64), ("strand", torch.uint8), ("iscore", torch.int32), ("p", torch.float64), ("q", torch.float64)):
            buf = torch.zeros(cap, dtype=dt, device=ctx.device)
            buf[:kept] = o[k][:kept]
            cols[k] = buf
    gathered = {k: ctx.allgather(v) for k, v in cols.items()}
    all_q = ctx.allgather(qtab.contiguous())
    ctx.sync()
    res = {"ok": True, "ranks": world, "rows": int(n), "windows": int(2 * n), "threshold": thr,
           "what": "rows split over the ranks, histogram all-reduced by gb2_allreduce_hist, per-rank finalize, columns merged "
                   "by gb2_allgather_bytes == one single-GPU scan of the same rows: q-table of every rank and (row, strand, "
                   "int score, p, q) of every hit, bit for bit"}
    if rank == 0:
        whole = engine.Scan(ctx, dm, strands=2, threshold=thr, want_q=True, hit_capacity=1 << 20)
        whole.score(rows)
        exp = whole.finalize()
        exp_q = whole.qtab.cpu().numpy()
        aq = all_q.cpu().numpy()
        qt_ok = all(np.array_equal(aq[r], exp_q) for r in range(world))
        parts = []
        for r in range(world):
            c = counts[r]
            parts.append({"row": gathered["row"][r, :c].cpu().numpy(), "strand": gathered["strand"][r, :c].cpu().numpy(),
                          "int_score": gathered["iscore"][r, :c].cpu().numpy(), "p-value": gathered["p"][r, :c].cpu().numpy(),
                          "q-value": gathered["q"][r, :c].cpu().numpy()})
        merged = gdist.merge_hit_tables(parts)
        cols_ok = {k: bool(np.array_equal(merged[k], exp[k])) for k in ("row", "strand", "int_score", "p-value", "q-value")}
        res.update(ok=bool(qt_ok and all(cols_ok.values())), hits=int(len(exp["row"])), hits_per_rank=counts,
                   qtable_equal_on_every_rank=bool(qt_ok), columns_equal=cols_ok)
    return res


def parity_single_gpu(ctx, dm, windows, seq_batch, scan_hist, scan_hits):
    """N=1 self-check at FULL size: the sequence form (windows formed in registers from the 2-bit haplotypes) must give
    the histogram and the hit table of the k-mer form (one packed word per window) bit for bit."""
    import torch
    from grafimo_b200 import engine
    sc = engine.Scan(ctx, dm, strands=2, threshold=THRESHOLD, want_q=True, hit_capacity=1 << 23)
    sc.score_sequences(seq_batch)
    got = sc.finalize()
    hist_ok = bool(torch.equal(sc.histogram(), scan_hist))
    cols_ok = {k: bool(np.array_equal(got[k], scan_hits[k])) for k in ("row", "strand", "int_score", "p-value", "q-value")}
    return {"ok": bool(hist_ok and all(cols_ok.values())), "ranks": 1, "windows": int(2 * windows.shape[0]), "hits": int(len(got["row"])),
            "what": "sequence-form scan (gb2_score_sequences on the 2-bit haplotypes) == k-mer-form scan (gb2_score on the expanded "
                    "packed windows) at full size: score histogram and (row, strand, int score, p, q) of every hit, bit for bit",
            "histogram_equal": hist_ok, "columns_equal": cols_ok}


def run_ours(args):
    import torch
    from grafimo_b200 import dist as gdist
    from grafimo_b200 import engine, synth
    from grafimo_b200.score_sequences import device_motif

    if args.gpus > 1 and "RANK" not in os.environ:  # plain `python bench.py --gpus N`: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 400), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        if args.region_len != REGION_LEN:
            cmd += ["--region-len", str(args.region_len)]
        if args.haplotypes != N_HAP:
            cmd += ["--haplotypes", str(args.haplotypes)]
        return subprocess.call(cmd)

    info = gdist.init_from_env("nccl")
    rank, world, local = info["rank"], info["world"], info["local"]
    torch.cuda.set_device(local)
    ctx = engine.Context(local)
    gdist.init_comm(ctx)  # the library's own NCCL communicator; torch.distributed only carried the 128-byte id
    tmpdir = tempfile.mkdtemp(prefix="gb2_bench_")
    motif = build_ctcf(tmpdir)
    dm = device_motif(motif, ctx)
    w = motif.width
    L, H = args.region_len, args.haplotypes
    per = L - w + 1
    n = per * H
    host_ascii = torch.empty((H, L), dtype=torch.uint8, pin_memory=True)  # the e2e input: the haplotypes as ASCII text
    with torch.cuda.stream(ctx.stream):
        windows, model = synth.haplotype_windows(L, H, w, SEED + rank, device=ctx.device, hap_batch=32)
        seq_words, _ = synth.haplotype_sequences(L, H, SEED + rank, device=ctx.device, hap_batch=32, ascii_out=host_ascii)
    ctx.sync()
    seq_batch = engine.SeqBatch(ctx, np.full(H, L, dtype=np.int64), seq2=seq_words.view(-1))
    scan = engine.Scan(ctx, dm, strands=2, threshold=THRESHOLD, want_q=True, hit_capacity=1 << 23)
    ev_pairs = []

    def step(timed, form="kmers"):
        scan.reset()
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(ctx.stream)
        if form == "kmers":
            scan.score(windows)
        else:
            scan.score_sequences(seq_batch)
        if timed:
            b.record(ctx.stream)
            ev_pairs.append((a, b))
        if world > 1:
            ctx.allreduce_hist(scan.histogram())  # ncclAllReduce issued by the library on the context's stream
        scan.qvalues()
        return scan.finalize_device()

    def timed_steps(form):
        del ev_pairs[:]
        for _ in range(max(args.warmup, 3)):
            kept = step(False, form)
        ctx.sync()
        if world > 1:
            torch.distributed.barrier()
        launches0 = ctx.launches
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if os.environ.get("GB2_PROFILE_RANGE") == form:  # ncu --profile-from-start off: capture the timed region only
            torch.cuda.profiler.start()
        t0.record(ctx.stream)
        for _ in range(args.steps):
            kept = step(True, form)
        t1.record(ctx.stream)
        ctx.sync()
        torch.cuda.synchronize()
        if os.environ.get("GB2_PROFILE_RANGE") == form:
            torch.cuda.profiler.stop()
        ms_total = t0.elapsed_time(t1)
        launches = ctx.launches - launches0
        if world > 1:
            torch.distributed.barrier()
        ms_total = ctx.allreduce_max([ms_total])[0]
        k_ms = float(np.mean([a.elapsed_time(b) for a, b in ev_pairs]))
        return ms_total / args.steps, k_ms, launches, kept

    sampler = ClockSampler(local)
    ms_step, k2_ms, launches, kept = timed_steps("kmers")
    clocks = sampler.stop()
    total_windows = 2 * n * world
    value = total_windows / (ms_step * 1e-3)
    n_hits = scan.n_hits()
    hist_kmers = scan.histogram().clone()
    hits_kmers = scan.finalize() if world == 1 else None
    algo_bytes = 8.0 * n + 32.0 * n_hits  # 8 B read per k-mer + 32 B per survivor (16 B record + its report row)
    peak, peak_src = peaks()
    achieved = algo_bytes / (k2_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "score_kernel_traffic.json")) as fh:
            tj = json.load(fh)
            if int(tj.get("n_kmers", -1)) == n and tj.get("kernel") == kernel_label(dm.info):
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj.get("source")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_label(dm.info), "kernel_ms": k2_ms,
                "kernel_share_of_step": k2_ms / ms_step, "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                "limiter": "shared-memory data pipe (ncu --set full, profiles/): per 32 k-mers n_chunks conflict-free LDS + 2 ATOMS "
                           "x ~4.2 wavefronts for the exact q-value histogram; the same kernel without the histogram (--no-qvalue) "
                           "runs at 0.99 of this peak (DESIGN.md sections 6-7)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"CTCF MA0139.1 (w=19, uniform bg) on a synthetic {L} bp region x {H} haplotype paths per GPU "
                               f"(1000G-like SNP/indel density), both strands, p<{THRESHOLD:g}, q-values on",
                   "kmers_per_gpu": n, "bytes_resident_per_gpu": 8 * n, "l2_policy": "inputs larger than L2 (no flush needed)",
                   "parallelism": f"rows sharded by region over {world} GPU(s); one all-reduce of the score histogram, issued by "
                                  "the library (gb2_allreduce_hist -> ncclAllReduce on the scan's stream)",
                   "hits_per_gpu": n_hits, "kept_after_finalize": kept,
                   "arithmetic": "integer scores as packed u16x2 adds in u32 (both strands per add); p/q-values f64"},
        "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline,
    }

    # ---- the same step with the windows formed on the device from the 2-bit haplotype sequences (0.25 B per window of
    #      HBM traffic instead of 8 B: no longer HBM-bound, the shared-memory pipe is the whole cost) ----------------------
    ms_seq, kseq_ms, launches_seq, _ = timed_steps("sequences")
    line["sequence_form"] = {
        "what": "same step, K2 over 2-bit haplotype sequences (gb2_score_sequences: windows formed in registers by funnel shifts)",
        "value": total_windows / (ms_seq * 1e-3), "unit": UNIT, "ms_per_step": ms_seq, "kernel_ms": kseq_ms,
        "bytes_resident_per_gpu": int(seq_words.numel() * 8), "hbm_bytes_per_window": 0.125, "gpu_launches": int(launches_seq),
        "speedup_vs_kmer_form": ms_step / ms_seq}

    # ---- parity carried by the bench itself (N>1: across real GPUs; N=1: sequence form == k-mer form at full size) --------
    if world > 1:
        parity = parity_multi_gpu(ctx, dm, rank, world)
    else:
        parity = parity_single_gpu(ctx, dm, windows, seq_batch, hist_kmers, hits_kmers)
    line["parity"] = parity

    # ---- e2e: the haplotypes as ASCII text in pinned host memory through gb2_scan_host_sequences (rank-local; max over
    #      ranks): host->device copy of every base + encode + score + BH + finalize + hit table back, all inside the timer --
    offs = np.arange(H, dtype=np.int64) * L
    lens = np.full(H, L, dtype=np.int64)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def time_host(fn):
        out = fn()  # warm-up (grows the context's staging pool)
        if world > 1:
            torch.distributed.barrier()
        ts = time.perf_counter()
        for _ in range(e2e_steps):
            out = fn()
        dt = (time.perf_counter() - ts) / e2e_steps
        return ctx.allreduce_max([dt])[0], out

    # the hit table comes back into reusable pinned buffers the caller owns (engine.HostTable), as the input lives in pinned
    # memory: both allocated once, outside the timed region; the copies themselves are inside
    host_out = engine.HostTable(1 << 23)

    def hit_bytes(out):
        return int(len(out["row"]) * (8 + 1 + 4 + 8 + 8 + 8) + 5 * 8 + 8)

    dt, out = time_host(lambda: engine.scan_host_sequences(ctx, dm, host_ascii.view(-1), offs, lens, fmt="ascii", strands=2,
                                                           threshold=THRESHOLD, hit_capacity=1 << 23, out=host_out))
    if world == 1 and hits_kmers is not None:
        parity["e2e_table_equal"] = bool(all(np.array_equal(out[k], hits_kmers[k]) for k in ("row", "strand", "int_score", "p-value", "q-value")))
        parity["ok"] = bool(parity["ok"] and parity["e2e_table_equal"])
    moved = ctx.last_transfer()  # counted by the library from the copies it issued in the last step
    line["e2e"] = {"value": 2.0 * n * world / dt, "unit": UNIT, "h2d_bytes_per_step": moved["h2d_bytes"], "d2h_bytes_per_step": moved["d2h_bytes"],
                   "input_bytes_per_step": int(H * L), "chunks_as_text": moved["chunks_as_given"], "chunks_packed_on_host": moved["chunks_host_packed"],
                   "host_cpus": os.cpu_count(),
                   "windows_per_step_per_gpu": int(2 * n), "steps": e2e_steps, "ms_per_step": dt * 1e3,
                   "api": "gb2_scan_host_sequences(format=ASCII): haplotype sequences as text in pinned host memory, 1 byte per "
                          "window in host memory; host threads re-code part of the chunks to 2 bits per base while the copy engine moves the others as text "
                          "(transfer compression, csrc/host_pack.cpp; nothing is scored on the host); windows formed on the device; hit table back into "
                          "pinned buffers (engine.HostTable)", "hits": int(len(out["row"]))}
    # what the host can deliver to this GPU while every rank copies at once: a bare pinned host -> device copy of 1 GiB of the same
    # input (no library code) -- the wall the copy-engine-only rate sits on (55 GB/s alone, ~23 GB/s with eight ranks on this pool's hosts)
    try:
        probe_bytes = int(min(host_ascii.numel(), 1 << 30))
        with torch.cuda.stream(ctx.stream):
            probe_dst = torch.empty(probe_bytes, dtype=torch.uint8, device=ctx.device)
        ctx.sync()
        best = None
        for rep in range(3):
            if world > 1:
                torch.distributed.barrier()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(ctx.stream):
                ea.record(ctx.stream)
                probe_dst.copy_(host_ascii.view(-1)[:probe_bytes], non_blocking=True)
                eb.record(ctx.stream)
            ctx.sync()
            ms = ea.elapsed_time(eb)
            best = ms if best is None or (rep and ms < best) else best
        worst_rank_ms = ctx.allreduce_max([best])[0]
        line["e2e"]["raw_h2d_GBps_per_gpu_all_ranks_copying"] = probe_bytes / (worst_rank_ms * 1e-3) / 1e9
        line["e2e"]["copy_engine_only_fraction_of_raw_h2d"] = None  # filled below
        del probe_dst
    except Exception as e:  # informational
        line["e2e"]["raw_h2d_GBps_per_gpu_all_ranks_copying"] = repr(e)
    variants = {}
    # (a) the same call with the host packers off: every base crosses PCIe as one byte of text -- the PCIe wall itself
    os.environ["GB2_HOST_PACK_THREADS"] = "0"
    dt0, out0 = time_host(lambda: engine.scan_host_sequences(ctx, dm, host_ascii.view(-1), offs, lens, fmt="ascii", strands=2,
                                                             threshold=THRESHOLD, hit_capacity=1 << 23, out=host_out))
    os.environ.pop("GB2_HOST_PACK_THREADS", None)
    moved0 = ctx.last_transfer()
    # the same call without any host-side help, next to the headline so that nobody has to look for it
    line["e2e"]["value_without_host_packers"] = 2.0 * n * world / dt0
    line["e2e"]["ms_per_step_without_host_packers"] = dt0 * 1e3
    raw = line["e2e"].get("raw_h2d_GBps_per_gpu_all_ranks_copying")
    if isinstance(raw, float) and raw > 0:
        line["e2e"]["copy_engine_only_fraction_of_raw_h2d"] = (H * L / dt0 / 1e9) / raw
    variants["sequences_ascii_copy_engine_only"] = {"value": 2.0 * n * world / dt0, "unit": UNIT, "ms_per_step": dt0 * 1e3,
                                                    "h2d_bytes_per_step": moved0["h2d_bytes"], "d2h_bytes_per_step": moved0["d2h_bytes"],
                                                    "api": "gb2_scan_host_sequences(format=ASCII), GB2_HOST_PACK_THREADS=0"}
    # (b) the same sequences already 2-bit packed on the host: 0.25 byte per window over PCIe
    host_words = torch.empty(seq_words.shape, dtype=torch.int64, pin_memory=True)
    with torch.cuda.stream(ctx.stream):
        host_words.copy_(seq_words, non_blocking=True)
    ctx.sync()
    woffs = np.arange(H, dtype=np.int64) * seq_words.shape[1]
    dt2, out2 = time_host(lambda: engine.scan_host_sequences(ctx, dm, host_words.view(-1), woffs, lens, fmt="2bit", strands=2,
                                                             threshold=THRESHOLD, hit_capacity=1 << 23, out=host_out))
    variants["sequences_2bit"] = {"value": 2.0 * n * world / dt2, "unit": UNIT, "ms_per_step": dt2 * 1e3,
                                  "h2d_bytes_per_step": int(host_words.numel() * 8), "d2h_bytes_per_step": hit_bytes(out2),
                                  "api": "gb2_scan_host_sequences(format=2-bit words)"}
    del host_words
    if world == 1 and not args.no_kmer_e2e:
        # (c), (d): one k-mer per window from the host, as the reference's TSV rows hold them -- a bounded number of rows
        rows_b = min(n, args.e2e_rows or (1 << 27))
        host_packed = torch.empty(rows_b, dtype=torch.int64, pin_memory=True)
        host_kmers = torch.empty((rows_b, w), dtype=torch.uint8, pin_memory=True)
        chunk = 1 << 24
        with torch.cuda.stream(ctx.stream):
            host_packed.copy_(windows[:rows_b], non_blocking=True)
            for lo in range(0, rows_b, chunk):
                hi = min(lo + chunk, rows_b)
                host_kmers[lo:hi].copy_(synth.windows_to_ascii(windows[lo:hi], w), non_blocking=True)
        ctx.sync()
        dt3, out3 = time_host(lambda: engine.scan_host_packed(ctx, dm, host_packed, None, strands=2, threshold=THRESHOLD, hit_capacity=1 << 23, out=host_out))
        variants["kmers_packed"] = {"value": 2.0 * rows_b / dt3, "unit": UNIT, "ms_per_step": dt3 * 1e3, "rows_per_step": int(rows_b),
                                    "h2d_bytes_per_step": int(rows_b * 8), "d2h_bytes_per_step": hit_bytes(out3),
                                    "api": "gb2_scan_host_packed (8 bytes per window over PCIe)"}
        dt4, out4 = time_host(lambda: engine.scan_host(ctx, dm, host_kmers, strands=2, threshold=THRESHOLD, hit_capacity=1 << 23, out=host_out))
        variants["kmers_ascii"] = {"value": 2.0 * rows_b / dt4, "unit": UNIT, "ms_per_step": dt4 * 1e3, "rows_per_step": int(rows_b),
                                   "h2d_bytes_per_step": int(rows_b * w), "d2h_bytes_per_step": hit_bytes(out4),
                                   "api": "gb2_scan_host (w = 19 ASCII bytes per window over PCIe; round 1's e2e entry)"}
        del host_packed, host_kmers
    line["e2e_variants"] = variants

    # ---- cpu baseline (rank 0, N=1 only): a bounded sample of the same workload on the host cores ----------------------
    if world == 1 and not args.no_cpu_baseline:
        with torch.cuda.stream(ctx.stream):
            fwd = synth.windows_to_ascii(windows[: 1 << 23], w).cpu().numpy()
        line["cpu_baseline"] = cpu_baseline_block(motif, fwd)
    # ---- graph path (informational, N=1 only): the same region as a variation graph -- reference + phased variants ->
    #      K7 extraction of the haplotype-aware k-mers -> K2/K5/K6 -> report table, no `vg`, no text
    if world == 1 and not args.no_graph_path:
        try:
            line["graph_path"] = graph_path_numbers(ctx, motif, L, H)
        except Exception as e:  # never lose the headline line over the informational block
            line["graph_path"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    ok = bool(line["parity"].get("ok", False)) if rank == 0 else True
    if world > 1:
        torch.distributed.barrier()
        ctx.close()
        torch.distributed.destroy_process_group()
    if not ok:
        sys.stderr.write("bench.py: PARITY FAILED -- " + json.dumps(line["parity"]) + "\n")
        return 1
    return 0


def cpu_baseline_block(motif, fwd):
    """The reference's CPU path on a bounded sample (10-20 s of CPU work) of the forward windows `fwd` + their reverse
    complements: the unmodified reference from oracle/_ref when present (kind "reference"), else the oracle port."""
    from grafimo_b200 import synth
    from oracle import oracle as orc
    orc.build()
    m = dict(score_matrix=motif.score_matrix_acgt(), pval_mat=motif.pval_matrix, min_val=motif.min_val,
             scale=motif.scale, offset=float(motif.offset))
    threads = os.cpu_count() or 1
    calib = np.ascontiguousarray(np.concatenate([fwd[:10000], synth.revcomp_ascii(fwd[:10000])]))
    port_rate = calib.shape[0] / time_oracle(calib, m, threads)
    ref = None
    try:
        ref = reference_objects(threads)
    except Exception as e:
        sys.stderr.write(f"bench.py: oracle/_ref unusable ({e!r}); cpu_baseline from the port\n")
    if ref is not None:
        rmotif, wf, compute_results = ref
        tmp = tempfile.mkdtemp(prefix="gb2_cpubase_")
        c = np.ascontiguousarray(np.concatenate([fwd[:50000], synth.revcomp_ascii(fwd[:50000])]))
        write_kmer_tsvs(c, os.path.join(tmp, "calib"), threads, rmotif.width)
        t_cal, _ = time_reference(compute_results, rmotif, wf, os.path.join(tmp, "calib"))
        k = int(max(50000, min(fwd.shape[0], (c.shape[0] / t_cal) * 10.0 / 2)))
        rows = np.ascontiguousarray(np.concatenate([fwd[:k], synth.revcomp_ascii(fwd[:k])]))
        write_kmer_tsvs(rows, os.path.join(tmp, "step"), 4 * threads, rmotif.width)
        t, hits = time_reference(compute_results, rmotif, wf, os.path.join(tmp, "step"))
        shutil.rmtree(tmp, ignore_errors=True)
        return {"value": rows.shape[0] / t, "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": f"{rows.shape[0]} TSV rows = first {k} forward 19-mers of the workload + their reverse-complement rows in "
                          f"{4 * threads} files; the UNMODIFIED reference's compute_results from oracle/_ref (numba scorer, --cores {threads}, "
                          f"BH, DataFrame; motif build and JIT excluded), {t:.1f} s, {hits} rows reported",
                "port_value": port_rate, "port_sample": "oracle C port of the same per-row algorithm, 20000 rows, same threads"}
    k = int(max(10000, min(fwd.shape[0], port_rate * 12.0 / 2)))
    rows = np.ascontiguousarray(np.concatenate([fwd[:k], synth.revcomp_ascii(fwd[:k])]))
    t = time_oracle(rows, m, threads)
    return {"value": rows.shape[0] / t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{rows.shape[0]} rows = first {k} forward 19-mers of the workload + their reverse-complement rows; "
                      f"oracle C port of the reference's per-row scoring (incl. two pval_mat sums per row), {threads} threads, {t:.1f} s"}


def graph_path_numbers(ctx, motif, L, H):
    import contextlib
    import io
    from grafimo_b200 import score_sequences as ss
    from grafimo_b200 import synth
    from grafimo_b200.vgraph import VariationGraph

    class A:
        cores, threshold, noqvalue, qvalueT, noreverse, recomb, verbose = 1, THRESHOLD, False, False, False, False, False
    t0 = time.perf_counter()
    ref, variants, gt = synth.variant_set(L, H, SEED)
    g = VariationGraph.build("1", ref, variants, gt)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    dg = g.to_device(ctx)
    ctx.sync()
    t_up = time.perf_counter() - t0
    w = motif.width
    te, tt = [], []
    for _ in range(4):
        t0 = time.perf_counter()
        rows = dg.extract([(0, L)], w)
        ctx.sync()
        te.append(time.perf_counter() - t0)
        with contextlib.redirect_stdout(io.StringIO()):
            df = ss.compute_results_rows(motif, rows, True, A)
        tt.append(time.perf_counter() - t0)
    return {"what": "variation graph of the same region shape (reference + phased variants, built on the host) -> K7 k-mer "
                    "extraction -> K2/K5/K6 -> report table; replaces `vg find` + TSV parse + scoring",
            "variants": len(variants), "haplotypes": H, "kmer_rows": rows.n, "windows_scored": 2 * rows.n,
            "extract_ms": min(te[1:]) * 1e3, "extract_rows_per_s": rows.n / min(te[1:]),
            "graph_to_table_ms": min(tt[1:]) * 1e3, "hits": int(len(df)),
            "host_graph_build_s": t_build, "graph_upload_s": t_up, "graph_bytes": int(g.cons_bits.nbytes + g.seq.nbytes + 24 * g.n_nodes + 8 * g.n_edges)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--region-len", type=int, default=REGION_LEN)
    ap.add_argument("--haplotypes", type=int, default=N_HAP)
    ap.add_argument("--e2e-rows", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph-path", action="store_true")
    ap.add_argument("--no-kmer-e2e", action="store_true", help="skip the two k-mer-per-window host entries (e2e_variants)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
