#!/usr/bin/env python
"""Builds `oracle/_ref/` -- the UNMODIFIED reference (pinellolab/GRAFIMO at /root/reference) in importable, compiled form --
TEST / BENCH INFRASTRUCTURE ONLY.

    python oracle/build_ref.py            (dev container: /root/reference is mounted read-only)

Outputs, all under oracle/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box like our own .so):
  * motif_processing.<abi>.so   cythonized from /root/reference/src/grafimo/motif_processing.pyx (the reference's one
                                native module, top-level name `motif_processing` as in its setup.py:53);
  * grafimo_ref.zip             the reference's Python modules byte-compiled from the sources where they lie
                                (grafimo/*.pyc inside one archive, imported through zipimport: no reference source
                                text is copied, nothing is written to /root/reference; one binary file because loose
                                *.pyc files are not shipped to the GPU box);
  * BUILD_INFO.json             versions + a self-check: the reference's compute_results on its own test fixture must
                                reproduce its expected table.
The two third-party modules the scoring path imports that are absent from this image (colorama; statsmodels.stats.multitest,
whose BH arithmetic is pinned by the reference's own golden table) come from tests/golden/_shims, exactly as in
tests/golden/make_golden.py.

Used by: bench.py --impl reference (kind "reference": the reference's own compute_results on the host cores) and by
`__graft_entry__.build()`.  Nothing under grafimo_b200/ imports it.  Without /root/reference (the GPU box) this script does
nothing and the prebuilt directory is used as it is."""
import glob
import json
import os
import py_compile
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
ZIP = os.path.join(OUT, "grafimo_ref.zip")
SHIMS = os.path.join(ROOT, "tests", "golden", "_shims")


def available():
    """True when oracle/_ref holds the built reference (nothing is imported, sys.path is left alone)."""
    return os.path.isfile(ZIP) and bool(glob.glob(os.path.join(OUT, "motif_processing*.so")))


def activate():
    """Puts the built reference (and the shims) on sys.path; -> True when it is importable."""
    if not available():
        return False
    for p in (SHIMS, OUT, ZIP):
        if p not in sys.path:
            sys.path.insert(0, p)
    return True


def build(force=False):
    if not os.path.isdir(REF):
        return activate()
    stamp = os.path.join(OUT, "BUILD_INFO.json")
    if os.path.exists(stamp) and not force:
        return activate()
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    scratch = tempfile.mkdtemp(prefix="gb2_refbuild_")
    pyx = os.path.join(REF, "src", "grafimo", "motif_processing.pyx")
    setup_py = os.path.join(scratch, "setup_mp.py")
    with open(setup_py, "w") as fh:
        fh.write("from setuptools import setup, Extension\nfrom Cython.Build import cythonize\nimport numpy\n"
                 f"ext = Extension('motif_processing', [{pyx!r}], include_dirs=[numpy.get_include()])\n"
                 f"setup(name='mp', ext_modules=cythonize([ext], build_dir={scratch!r}, language_level=3))\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(REF, "src"), SHIMS, os.environ.get("PYTHONPATH", "")]))
    subprocess.check_call([sys.executable, setup_py, "build_ext", "--build-lib", OUT, "--build-temp", scratch, "-q"], cwd=scratch, env=env)
    import zipfile
    with zipfile.ZipFile(ZIP, "w", zipfile.ZIP_DEFLATED) as zf:
        for src in sorted(glob.glob(os.path.join(REF, "src", "grafimo", "*.py"))):
            name = os.path.basename(src)
            cfile = os.path.join(scratch, name + "c")
            py_compile.compile(src, cfile=cfile, dfile=f"grafimo/{name}", doraise=True)
            zf.write(cfile, f"grafimo/{name}c")
    shutil.rmtree(scratch, ignore_errors=True)
    info = {"reference": "pinellolab/GRAFIMO", "python": sys.version.split()[0]}
    try:
        import Cython, numba, numpy, pandas  # noqa: E401
        info.update(cython=Cython.__version__, numba=numba.__version__, numpy=numpy.__version__, pandas=pandas.__version__)
    except Exception:
        pass
    info["self_check"] = self_check()
    with open(stamp, "w") as fh:
        json.dump(info, fh, indent=1)
    return activate()


def self_check():
    """the reference's own test_scoring assertion (tests/grafimo_run_test.py:119-140) on the built copy, in a subprocess"""
    code = r'''
import io, json, os, sys, tempfile, contextlib
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
import pandas as pd
from grafimo.motif_ops import build_motif_meme
from grafimo.score_sequences import compute_results
fx = json.load(open(%r))
tmp = tempfile.mkdtemp()
os.makedirs(os.path.join(tmp, "width_19"))
open(os.path.join(tmp, "m.meme"), "w").write(fx["ctcf_meme"])
open(os.path.join(tmp, "width_19", "in.tsv"), "w").write(fx["scoring_input_tsv"])
with contextlib.redirect_stdout(io.StringIO()):
    motif = build_motif_meme(os.path.join(tmp, "m.meme"), "unfrm_dst", 0.1, False, 1, False, True)[0]
    df = compute_results(motif, tmp, True, testmode=True)
exp = pd.read_csv(io.StringIO(fx["scoring_results_tsv"]), sep="\t", index_col=0, float_precision="round_trip")
key = ["start", "stop", "strand", "matched_sequence"]
a = df.sort_values(key, kind="stable").reset_index(drop=True); b = exp.sort_values(key, kind="stable").reset_index(drop=True)
ok = len(a) == len(b) == 704 and all((a[c] == b[c]).all() for c in ["score", "p-value", "q-value", "start", "stop", "haplotype_frequency"])
print("SELFCHECK", "ok" if ok else "MISMATCH")
''' % (SHIMS, OUT, ZIP, os.path.join(ROOT, "tests", "golden", "fixtures.json"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    if "SELFCHECK ok" not in r.stdout:
        raise SystemExit("oracle/_ref self-check failed:\n" + r.stdout[-1500:] + r.stderr[-3000:])
    return "reference compute_results on its own 704-row fixture == its expected table (score, p, q, coordinates, frequency)"


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref", "ready" if ok else "not available (no /root/reference and no prebuilt copy)")
