"""The drop-in boundary from plain C: examples/c_abi_scan.c (gcc, include/grafimo_b200.h, libgrafimo_b200.so -- no Python,
no torch in that process) against the Python binding on the same motif and sequences."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_program_equals_python_binding(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc on this box")
    from grafimo_b200 import engine
    exe = str(tmp_path / "c_abi_scan")
    libdir = os.path.join(ROOT, "grafimo_b200")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_scan.c"), "-o", exe,
                    "-L", libdir, "-lgrafimo_b200", f"-Wl,-rpath,{libdir}", "-lm"], check=True)
    m = gu.load_motif("ctcf_meme__bgnt")
    w = m["width"]
    sm = np.ascontiguousarray(m["score_matrix"], dtype=np.int64)
    with open(tmp_path / "motif.bin", "wb") as fh:
        fh.write(struct.pack("<qqqd", w, int(m["min_val"]), int(m["scale"]), float(m["offset"])))
        fh.write(np.asarray(m["bg_acgt"], dtype=np.float64).tobytes())
        fh.write(sm.tobytes())
    rng = np.random.default_rng(314)
    seqs = []
    for n in (5000, 18, 19, 250000, 64, 70001):
        s = rng.choice(np.array(list("ACGTacgt")), size=n)
        s = np.where(rng.random(n) < 0.001, "N", s)
        seqs.append("".join(s))
    (tmp_path / "seqs.txt").write_text("\n".join(seqs) + "\n")
    r = subprocess.run([exe, str(tmp_path / "motif.bin"), str(tmp_path / "seqs.txt"), "0.002"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = [ln.split("\t") for ln in r.stdout.strip("\n").split("\n")]
    ctx = engine.Context(0)
    pv = ctx.pval_dp_batched([sm], [m["bg_acgt"]])[0]
    assert np.array_equal(pv, m["pval_mat"])
    dm = ctx.motif(sm, pv, m["min_val"], m["scale"], m["offset"])
    text = np.frombuffer(("\n".join(seqs) + "\n").encode(), dtype=np.uint8)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(lens + 1)[:-1]]).astype(np.int64)
    exp = engine.scan_host_sequences(ctx, dm, text, offs, lens, fmt="ascii", strands=2, threshold=0.002)
    assert len(got) == len(exp["row"]) > 50
    assert [int(g[0]) for g in got] == exp["row"].astype(np.int64).tolist()
    assert [g[1] for g in got] == ["-" if s else "+" for s in exp["strand"]]
    assert [int(g[2]) for g in got] == exp["int_score"].tolist()
    for col, key in ((3, "score"), (4, "p-value"), (5, "q-value")):
        assert np.array_equal(np.array([float(g[col]) for g in got]), exp[key]), key  # %.17g round-trips a double
    assert f"{len(seqs)} sequences" in r.stderr
    ctx.close()
