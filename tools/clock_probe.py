"""Runs K2 back to back for a few seconds while sampling nvidia-smi clocks (B200_PROFILING.md recipe)."""
import os, subprocess, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import golden_util as gu
from grafimo_b200.engine import Context, Scan

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
ctx = Context(0)
m = gu.load_motif("ctcf_meme__unif")
dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
n = 1 << 30
packed = torch.randint(0, 1 << 38, (n,), dtype=torch.int64, device="cuda")
os.makedirs("gpurun_out", exist_ok=True)
for want_q in (True, False):
    sc = Scan(ctx, dm, strands=2, threshold=1e-4, want_q=want_q, hit_capacity=1 << 22)
    sc.score(packed); ctx.sync()
    q = "index,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    f = open(f"gpurun_out/clocks_hist{int(want_q)}.csv", "w")
    p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv", "-lms", "100"], stdout=f)
    time.sleep(0.3)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.time(); reps = 0
    e0.record(ctx.stream)
    while time.time() - t0 < secs:
        for _ in range(20):
            sc.reset(); sc.score(packed)
        reps += 20
        ctx.sync()
    e1.record(ctx.stream); ctx.sync()
    ms = e0.elapsed_time(e1) / reps
    p.terminate(); p.wait(); f.close()
    rows = [r.split(", ") for r in open(f.name).read().strip().split("\n")[1:]]
    sm = [int(r[1].split()[0]) for r in rows]
    pw = [float(r[4].split()[0]) for r in rows]
    print(f"hist={int(want_q)}: {ms:.3f} ms/launch  {n*8/ms/1e6:.0f} GB/s   sm MHz median {np.median(sm):.0f} min {min(sm)} max {max(sm)}; power median {np.median(pw):.0f} W max {max(pw):.0f}; reasons {set(r[6] for r in rows)} pcap {set(r[10] for r in rows)}")
