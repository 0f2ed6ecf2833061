"""Kernel-level timing of K2 (score) on random packed k-mers; prints k-mers/s and algorithmic GB/s.
Run on the GPU box:  python tools/microbench_score.py [log2_n]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import golden_util as gu  # noqa: E402
from grafimo_b200.engine import Context, Scan  # noqa: E402


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    n = 1 << lg
    ctx = Context(0)
    peaks = {}
    try:
        peaks = json.load(open("MEASURED_PEAKS.json"))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for tag in ("ctcf_meme__unif", "synth_w8_meme__bgnt", "synth_w30_meme__bgnt"):
        m = gu.load_motif(tag)
        w = m["width"]
        dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
        packed = torch.randint(0, 1 << (2 * w), (n,), dtype=torch.int64, device="cuda", generator=g) if w < 32 else \
            torch.randint(-(1 << 62), 1 << 62, (n,), dtype=torch.int64, device="cuda", generator=g)
        torch.cuda.synchronize()
        print(f"## {tag}: w={w} span={dm.span} chunks={dm.info.n_chunks} R={dm.info.lut_replicas} smem={dm.info.smem_bytes}")
        for strands in (2, 1):
            for want_q in (True, False):
                for thr in (1e-4,):
                    sc = Scan(ctx, dm, strands=strands, threshold=thr, want_q=want_q, hit_capacity=1 << 22)
                    for _ in range(3):
                        sc.reset(); sc.score(packed)
                    ctx.sync()
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    reps = 5
                    e0.record(ctx.stream)
                    for _ in range(reps):
                        sc.score(packed)
                    e1.record(ctx.stream)
                    ctx.sync()
                    ms = e0.elapsed_time(e1) / reps
                    hits = sc.n_hits() / (reps + 1) if False else None
                    gbs = n * 8 / ms / 1e6
                    print(f"strands={strands} hist={int(want_q)} thr={thr:g}: {ms:8.3f} ms  {n / ms / 1e6:8.2f} Gkmer/s  "
                          f"{n * strands / ms / 1e6:8.2f} Gwin/s  {gbs:8.1f} GB/s  frac={gbs / peak:.3f}")
        del packed
    # copy roofline reference measured the same way
    a = torch.empty(1 << 28, dtype=torch.int64, device="cuda"); b = torch.empty_like(a)
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"torch copy 2 GiB->2 GiB: {ms:.3f} ms  {2 * a.numel() * 8 / ms / 1e6:.1f} GB/s (read+write)")
    s = a.sum()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        s = a.sum()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"torch sum (read-only) 2 GiB: {ms:.3f} ms  {a.numel() * 8 / ms / 1e6:.1f} GB/s")


if __name__ == "__main__":
    main()
