"""Times the drop-in API `compute_results` (TSV files -> report table) on a synthetic vg-like k-mer TSV and, for
scale, the CPU oracle port on a sample of the same rows.   python tools/bench_compute_results.py [million_rows]"""
import contextlib, io, os, sys, tempfile, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu
from grafimo_b200 import synth
from grafimo_b200.motif_ops import build_motif_meme
from grafimo_b200.score_sequences import compute_results
from grafimo_b200.workflow import Findmotif

mrows = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
threshold = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-4   # 1 = report every row (the paper's `-t 1` runs)
n_kmers = int(mrows * 1e6 / 2)
tmp = tempfile.mkdtemp(prefix="gb2_cr_")
meme = os.path.join(tmp, "MA0139.1.meme"); open(meme, "w").write(gu.fixtures()["ctcf_meme"])
with contextlib.redirect_stdout(io.StringIO()):
    motif = build_motif_meme(meme, "unfrm_dst", 0.1, False, 1, False, True)[0]
w = 19
t0 = time.time()
per = 200000 - w + 1
packed, _ = synth.haplotype_windows(200000, (n_kmers + per - 1) // per, w, 5, device="cuda")
fwd = synth.windows_to_ascii(packed[:n_kmers], w).cpu().numpy()
rc = synth.revcomp_ascii(fwd)
d = os.path.join(tmp, "kmers", "width_19"); os.makedirs(d)
pos = np.arange(n_kmers) + 1000000
with open(os.path.join(d, "chr7.tsv"), "wb") as fh:   # vectorised writer: both strands, vg-like columns
    B = 1 << 20
    for lo in range(0, n_kmers, B):
        hi = min(lo + B, n_kmers)
        f = np.char.decode(fwd[lo:hi].view("S19").ravel(), "ascii"); r = np.char.decode(rc[lo:hi].view("S19").ravel(), "ascii")
        p = pos[lo:hi].astype(str); q = (pos[lo:hi] + w).astype(str)
        plus = np.char.add(np.char.add(np.char.add("7:1000000-9000000\t", f), np.char.add("\t7:", p)), np.char.add(np.char.add("+\t7:", q), "+\t2504\tref\t101+,102+,\n"))
        minus = np.char.add(np.char.add(np.char.add("7:1000000-9000000\t", r), np.char.add("\t7:", q)), np.char.add(np.char.add("-\t7:", p), "-\t2504\tref\t102-,101-,\n"))
        fh.write("".join(np.stack([plus, minus], 1).ravel().tolist()).encode())
size = os.path.getsize(os.path.join(d, "chr7.tsv"))
print(f"TSV: {2 * n_kmers} rows, {size / 1e9:.2f} GB, generated in {time.time() - t0:.1f}s")
wf = Findmotif(motif=[meme], kmers_dir=os.path.join(tmp, "kmers"), threshold=threshold, recomb=True, verbose=True)
from grafimo_b200.score_sequences import clear_parsed_cache
for rep in range(4):
    if rep < 3:
        clear_parsed_cache()  # runs 0-2 read and parse the files; run 3 is the "next motif of the same width" (rows reused)
    t = time.time()
    with contextlib.redirect_stdout(io.StringIO()) as out:
        df = compute_results(motif, os.path.join(tmp, "kmers"), True, wf)
    dt = time.time() - t
    print(f"compute_results run {rep}{' (parsed rows reused)' if rep == 3 else ''}: {dt:.3f}s  {2 * n_kmers / dt / 1e6:.1f} M rows/s  "
          f"({size / dt / 1e9:.2f} GB/s of text)  hits={len(df)}")
print(out.getvalue().strip().replace("\n\n", "\n"))
if threshold >= 1.0:
    from grafimo_b200.res_writer import write_results
    wf2 = Findmotif(motif=[meme], kmers_dir=os.path.join(tmp, "kmers"), threshold=threshold, out=os.path.join(tmp, "out"), text_only=True, verbose=True)
    t = time.time()
    write_results(df, motif, 1, wf2, True)
    print(f"write_results (TSV + GFF3, no HTML): {time.time() - t:.2f}s for {len(df)} rows")
    from grafimo_b200.res_writer import write_results_device
    from grafimo_b200.score_sequences import scan_dir_device
    wf3 = Findmotif(motif=[meme], kmers_dir=os.path.join(tmp, "kmers"), threshold=threshold, recomb=True, out=os.path.join(tmp, "out_dev"), text_only=True, verbose=True)
    for rep in range(2):
        t = time.time()
        with contextlib.redirect_stdout(io.StringIO()):
            report = scan_dir_device(motif, os.path.join(tmp, "kmers"), True, wf3)
        t1 = time.time()
        with contextlib.redirect_stdout(io.StringIO()):
            write_results_device(report, motif, 1, wf3, True)
        t2 = time.time()
        print(f"device report path run {rep}: TSV dir -> device columns {t1 - t:.2f}s, K8 TSV + GFF3 files {t2 - t1:.2f}s for {report.n} rows")
from oracle import oracle as orc
k = 200000
rows = np.ascontiguousarray(np.concatenate([fwd[:k], rc[:k]]))
t = time.time()
orc.score_rows(rows, motif.score_matrix_acgt(), motif.pval_matrix, motif.min_val, motif.scale, float(motif.offset), nthreads=os.cpu_count())
dt = time.time() - t
print(f"oracle port (scoring only, {os.cpu_count()} threads): {rows.shape[0] / dt / 1e6:.2f} M rows/s")
