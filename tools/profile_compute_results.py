"""Phase timing of compute_results on a synthetic TSV (host phases with perf_counter, GPU with events)."""
import contextlib, io, os, sys, tempfile, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu
from grafimo_b200 import synth, engine
from grafimo_b200.motif_ops import build_motif_meme
from grafimo_b200 import score_sequences as ss

n_kmers = int(float(sys.argv[1]) * 1e6 / 2) if len(sys.argv) > 1 else 5_000_000
tmp = tempfile.mkdtemp(prefix="gb2_cr_")
meme = os.path.join(tmp, "MA0139.1.meme"); open(meme, "w").write(gu.fixtures()["ctcf_meme"])
with contextlib.redirect_stdout(io.StringIO()):
    motif = build_motif_meme(meme, "unfrm_dst", 0.1, False, 1, False, True)[0]
w = 19
per = 200000 - w + 1
packed, _ = synth.haplotype_windows(200000, (n_kmers + per - 1) // per, w, 5, device="cuda")
fwd = synth.windows_to_ascii(packed[:n_kmers], w).cpu().numpy(); rc = synth.revcomp_ascii(fwd)
d = os.path.join(tmp, "kmers", "width_19"); os.makedirs(d)
pos = np.arange(n_kmers) + 1000000
fn = os.path.join(d, "chr7.tsv")
with open(fn, "wb") as fh:
    B = 1 << 20
    for lo in range(0, n_kmers, B):
        hi = min(lo + B, n_kmers)
        f = np.char.decode(fwd[lo:hi].view("S19").ravel(), "ascii"); r = np.char.decode(rc[lo:hi].view("S19").ravel(), "ascii")
        p = pos[lo:hi].astype(str); q = (pos[lo:hi] + w).astype(str)
        plus = np.char.add(np.char.add(np.char.add("7:1000000-9000000\t", f), np.char.add("\t7:", p)), np.char.add(np.char.add("+\t7:", q), "+\t2504\tref\t101+,102+,\n"))
        minus = np.char.add(np.char.add(np.char.add("7:1000000-9000000\t", r), np.char.add("\t7:", q)), np.char.add(np.char.add("-\t7:", p), "-\t2504\tref\t102-,101-,\n"))
        fh.write("".join(np.stack([plus, minus], 1).ravel().tolist()).encode())
size = os.path.getsize(fn)
ctx = ss._context(); dm = ss.device_motif(motif, ctx)
for rep in range(3):
    t0 = time.perf_counter()
    texts = list(ss._text_chunks([fn], ss._CHUNK_BYTES))
    t1 = time.perf_counter()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(ctx.stream)
    with torch.cuda.stream(ctx.stream):
        d_texts = [t.to(ctx.device, non_blocking=True) for t in texts]
    e[1].record(ctx.stream)
    rows = [ctx.parse_kmer_tsv(t, w) for t in d_texts]
    e[2].record(ctx.stream)
    n = sum(r.n for r in rows)
    scan = engine.Scan(ctx, dm, strands=1, threshold=1e-4, hit_capacity=n)
    base = 0
    for r in rows:
        scan.score(r.packed, None, row_base=base); base += r.n
    kept = scan.finalize_device()
    e[3].record(ctx.stream); ctx.sync()
    t2 = time.perf_counter()
    print(f"rep {rep}: {2 * n_kmers} rows, {size / 1e9:.2f} GB | read files->pinned {t1 - t0:.3f}s ({size / (t1 - t0) / 1e9:.1f} GB/s) | "
          f"H2D {e[0].elapsed_time(e[1]):.1f} ms | index+parse {e[1].elapsed_time(e[2]):.1f} ms ({size / e[1].elapsed_time(e[2]) / 1e6:.0f} GB/s of text) | "
          f"score+BH+finalize {e[2].elapsed_time(e[3]):.1f} ms | total {t2 - t0:.3f}s")
