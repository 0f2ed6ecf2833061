#!/usr/bin/env python
"""End-to-end rate of gb2_scan_host_sequences on the headline workload (CTCF, 1 Mb x 2,504 haplotypes as ASCII text in pinned host
memory -> hit table in pinned host memory) against the number of host packer threads (csrc/host_pack.cpp: transfer compression).

    python tools/bench_e2e.py [--threads 0,4,8,16] [--steps 3] [--out FILE]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", default="0,2,4,8,12,16,20")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--region-len", type=int, default=1_000_000)
    ap.add_argument("--haplotypes", type=int, default=2504)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import numpy as np
    import torch
    import bench as B
    from grafimo_b200 import engine, synth
    from grafimo_b200.score_sequences import device_motif
    ctx = engine.Context(0)
    motif = B.build_ctcf(tempfile.mkdtemp(prefix="gb2_e2e_"))
    dm = device_motif(motif, ctx)
    w, L, H = motif.width, a.region_len, a.haplotypes
    n = (L - w + 1) * H
    host_ascii = torch.empty((H, L), dtype=torch.uint8, pin_memory=True)
    with torch.cuda.stream(ctx.stream):
        synth.haplotype_sequences(L, H, B.SEED, device=ctx.device, hap_batch=32, ascii_out=host_ascii)
    ctx.sync()
    offs = np.arange(H, dtype=np.int64) * L
    lens = np.full(H, L, dtype=np.int64)
    table = engine.HostTable(1 << 23)
    ref = None
    rows = []
    for t in [int(x) for x in a.threads.split(",")]:
        os.environ["GB2_HOST_PACK_THREADS"] = str(t)
        run = lambda: engine.scan_host_sequences(ctx, dm, host_ascii.view(-1), offs, lens, fmt="ascii", strands=2, threshold=B.THRESHOLD,  # noqa: E731
                                                 hit_capacity=1 << 23, out=table)
        out = run()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            out = run()
        dt = (time.perf_counter() - t0) / a.steps
        key = {k: out[k].copy() for k in ("row", "strand", "int_score", "p-value", "q-value")}
        if ref is None:
            ref = key
        same = all(np.array_equal(key[k], ref[k]) for k in ref)
        rows.append({"packer_threads": t, "ms_per_step": dt * 1e3, "windows_per_s": 2.0 * n / dt, "text_GBps_equivalent": H * L / dt / 1e9,
                     "hits": int(len(out["row"])), "table_equals_device_only_route": bool(same)})
        print(json.dumps(rows[-1]), flush=True)
    if a.out:
        with open(a.out, "w") as fh:
            json.dump({"workload": f"CTCF, {L} bp x {H} haplotypes as ASCII text, both strands, p<{B.THRESHOLD:g}", "host_cpus": os.cpu_count(), "runs": rows}, fh, indent=1)


if __name__ == "__main__":
    main()
