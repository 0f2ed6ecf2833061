#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Run here (dev container, /root/reference mounted read-only):

    python tests/golden/make_golden.py

What it does
  1. cythonizes /root/reference/src/grafimo/motif_processing.pyx into a scratch
     directory (nothing is copied into this repository, nothing is written to
     /root/reference) and puts /root/reference/src on sys.path together with the two
     import shims in tests/golden/_shims (colorama; statsmodels.stats.multitest -- the
     only third-party pieces the scoring path needs that are not installed here);
  2. runs the reference's own pytest cases for the hot path (5 tests) as a self-check;
  3. runs the reference entry points
        build_motif_{meme,jaspar,transfac,pfm}   (motif_ops.py:51,237,640,809)
        compute_results                          (score_sequences.py:44)
     on the reference's own fixtures and on seeded synthetic inputs, and stores inputs and
     outputs as .npz bundles (no pickles) in tests/golden/cases/.

The bundles are data (numbers and the fixture rows), not reference source code.  They travel to
the GPU box, where /root/reference does not exist; tests/ reads only the bundles.
"""
import argparse
import glob
import io
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
CASES = os.path.join(HERE, "cases")


# ----------------------------------------------------------------------------------------
# reference bootstrap
# ----------------------------------------------------------------------------------------
def bootstrap(scratch):
    pyx = os.path.join(REF, "src", "grafimo", "motif_processing.pyx")
    build = os.path.join(scratch, "build")
    os.makedirs(build, exist_ok=True)
    setup_py = os.path.join(build, "setup_mp.py")
    with open(setup_py, "w") as fh:
        fh.write(
            "from setuptools import setup, Extension\n"
            "from Cython.Build import cythonize\n"
            "import numpy\n"
            f"ext = Extension('motif_processing', [{pyx!r}], include_dirs=[numpy.get_include()])\n"
            f"setup(name='mp', ext_modules=cythonize([ext], build_dir={build!r}, language_level=3))\n"
        )
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(
        [os.path.join(REF, "src"), os.path.join(HERE, "_shims"), env.get("PYTHONPATH", "")]
    )
    subprocess.check_call(
        [sys.executable, setup_py, "build_ext", "--build-lib", build, "--build-temp", build, "-q"],
        cwd=build,
        env=env,
    )
    for p in (build, os.path.join(HERE, "_shims"), os.path.join(REF, "src")):
        if p not in sys.path:
            sys.path.insert(0, p)
    return env, build


def run_reference_pytests(scratch, env, build):
    tdir = os.path.join(scratch, "tests")
    shutil.copytree(os.path.join(REF, "tests"), tdir)
    env = dict(env)
    env["PYTHONPATH"] = os.pathsep.join([build, env["PYTHONPATH"]])
    out = subprocess.run(
        [sys.executable, "-m", "pytest", "grafimo_run_test.py", "-q", "-k", "motif_processing or scoring",
         "-p", "no:cacheprovider"],
        cwd=tdir, env=env, capture_output=True, text=True,
    )
    print(out.stdout[-600:])
    if out.returncode != 0:
        print(out.stderr[-2000:])
        raise SystemExit("reference hot-path tests failed in this container")


# ----------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------
def read_text(path):
    with open(path) as fh:
        return fh.read()


def ustr(xs):
    return np.array(list(xs), dtype=np.str_)


def motif_bundle(motif, extra=None):
    bg = motif.bg
    d = dict(
        width=np.int64(motif.width),
        motif_id=ustr([motif.motif_id]),
        motif_name=ustr([motif.motif_name]),
        count_matrix=np.asarray(motif.count_matrix, dtype=np.float64),
        score_matrix=np.asarray(motif.score_matrix, dtype=np.int64),
        pval_mat=np.asarray(motif.pval_matrix, dtype=np.float64),
        min_val=np.int64(motif.min_val),
        max_val=np.int64(motif.max_val),
        scale=np.int64(motif.scale),
        offset=np.float64(motif.offset),
        bg_acgt=np.array([bg["A"], bg["C"], bg["G"], bg["T"]], dtype=np.float64),
        bg_key_order=ustr(list(bg.keys())),
    )
    if extra:
        d.update(extra)
    return d


def synth_meme(rng, width, alpha, nsites, name):
    probs = rng.dirichlet([alpha] * 4, size=width)
    # MEME files carry 6 decimals; renormalise the last column like real files do not (leave as is)
    lines = [
        "MEME version 4", "", "ALPHABET= ACGT", "", "strands: + -", "",
        "Background letter frequencies", "A 0.25 C 0.25 G 0.25 T 0.25", "",
        f"MOTIF {name} SYN{width}",
        f"letter-probability matrix: alength= 4 w= {width} nsites= {nsites} E= 0",
    ]
    for row in probs:
        row = np.maximum(row, 0.0)
        lines.append(" " + "  ".join(f"{v:.6f}" for v in row))
    lines.append("URL none")
    lines.append("")
    return "\n".join(lines)


def df_bundle(df):
    out = {"columns": ustr(df.columns)}
    for c in df.columns:
        col = df[c].to_numpy()
        key = "col_" + c.replace("-", "_")
        if col.dtype.kind in "OUS" or str(df[c].dtype).startswith(("str", "object")):
            out[key] = ustr(col.tolist())
        elif col.dtype.kind == "f":
            out[key] = col.astype(np.float64)
        else:
            out[key] = col.astype(np.int64)
    return out


class Args:
    pass


def make_findmotif(Findmotif, cores=1, threshold=1e-4, noqvalue=False, qvalueT=False, noreverse=False,
                   recomb=False, verbose=False):
    wf = Findmotif.__new__(Findmotif)
    wf._cores = cores
    wf._thresh = float(threshold)
    wf._no_qvalue = noqvalue
    wf._qvalueT = qvalueT
    wf._no_rev = noreverse
    wf._recomb = recomb
    wf._verbose = verbose
    return wf


def synth_rows(rng, n, width, chrom="7", region_start=1000, with_n=0, lower=0, indel_frac=0.05):
    """vg-find-like 7-column rows for random k-mers, both strands (the '-' twin is the reverse
    complement with swapped coordinates, like the reference fixture)."""
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N", "a": "t", "c": "g", "g": "c", "t": "a"}
    region = f"{chrom}:{region_start}-{region_start + n + width}"
    lines = []
    for i in range(n):
        seq = "".join(rng.choice(list("ACGT"), size=width))
        if i < with_n:
            pos = int(rng.integers(0, width))
            seq = seq[:pos] + "N" + seq[pos + 1:]
        elif i < with_n + lower:
            seq = seq.lower()
        start = region_start + i
        span = width if rng.random() > indel_frac else width + int(rng.integers(1, 4))
        stop = start + span
        freq = int(rng.choice([0, 1, 2, 17, 2504, 5008]))
        ref = "ref" if rng.random() < 0.7 else "non.ref"
        path = ",".join(f"{int(x)}+" for x in rng.integers(1, 10 ** 6, size=2)) + ","
        lines.append(f"{region}\t{seq}\t{chrom}:{start}+\t{chrom}:{stop}+\t{freq}\t{ref}\t{path}")
        rc = "".join(comp[c] for c in reversed(seq))
        rpath = ",".join(p.replace("+", "-") for p in reversed(path.strip(",").split(","))) + ","
        lines.append(f"{region}\t{rc}\t{chrom}:{stop}-\t{chrom}:{start}-\t{freq}\t{ref}\t{rpath}")
    order = rng.permutation(len(lines))
    return [lines[i] for i in order]


def run_scoring(compute_results, motif, lines_per_file, width, scratch, tag, args_obj=None, testmode=False):
    loc = os.path.join(scratch, "seqs_" + tag)
    wdir = os.path.join(loc, f"width_{width}")
    os.makedirs(wdir)
    for k, lines in enumerate(lines_per_file):
        with open(os.path.join(wdir, f"region_{k}.tsv"), "w") as fh:
            fh.write("\n".join(lines) + "\n")
    buf = io.StringIO()
    old = sys.stdout
    sys.stdout = buf
    try:
        df = compute_results(motif, loc + "/", True, args_obj, testmode=testmode)
    finally:
        sys.stdout = old
    return df, buf.getvalue()


# ----------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-pytests", action="store_true")
    args = ap.parse_args()
    if not os.path.isdir(REF):
        raise SystemExit("/root/reference is not mounted; golden vectors can only be regenerated in the dev container")
    os.makedirs(CASES, exist_ok=True)
    scratch = tempfile.mkdtemp(prefix="grafimo_golden_")
    env, build = bootstrap(scratch)
    if not args.skip_pytests:
        run_reference_pytests(scratch, env, build)

    from grafimo.motif_ops import build_motif_meme, build_motif_jaspar, build_motif_transfac, build_motif_pfm
    from grafimo.score_sequences import compute_results
    from grafimo.workflow import Findmotif
    from grafimo.res_writer import writeGFF3

    tdata = os.path.join(REF, "tests", "test_data")
    inp = os.path.join(tdata, "input")
    exp = os.path.join(tdata, "expected_results")
    bg_nt = os.path.join(REF, "tutorials", "findmotif_tutorial", "data", "bg_nt")
    UNIF = "unfrm_dst"

    # ---- (A) reference fixtures, verbatim data -------------------------------------------------
    fixtures = dict(
        ctcf_meme=read_text(os.path.join(inp, "MA0139.1.meme")),
        ctcf_jaspar=read_text(os.path.join(inp, "MA0139.1.jaspar")),
        ctcf_transfac=read_text(os.path.join(inp, "MA0139.1.transfac")),
        ctcf_pfm=read_text(os.path.join(inp, "MA0139.1.pfm")),
        bg_nt=read_text(bg_nt),
        atf3_meme=read_text(os.path.join(REF, "docs/paper_results/tf_motifs/ATF3/MA0605.2.meme")),
        gata1_meme=read_text(os.path.join(REF, "docs/paper_results/tf_motifs/GATA1/MA0035.4.meme")),
        scoring_input_tsv=read_text(os.path.join(inp, "width_19", "scoring_test_input.tsv")),
        scoring_results_tsv=read_text(os.path.join(exp, "scoring_results.tsv")),
        expected_seqs_tsv=read_text(os.path.join(exp, "expected_seqs.tsv")),
        expected_matrix_meme=read_text(os.path.join(exp, "motif_processing_test_meme.txt")),
        expected_matrix_jaspar=read_text(os.path.join(exp, "motif_processing_test_jaspar.txt")),
        # inputs of the reference's vg tests (tests/grafimo_run_test.py:15-63): the graph expected_seqs.tsv came from
        test_fa=read_text(os.path.join(inp, "test.fa")),
        test_vcf=gzip.open(os.path.join(inp, "test.vcf.gz"), "rt").read(),
    )
    rng = np.random.default_rng(20242)
    synth = {}
    for width, alpha, nsites in ((6, 0.5, 120), (8, 0.5, 300), (11, 0.5, 1000), (25, 0.1, 800), (27, 0.3, 2500),
                                 (30, 0.1, 4000), (32, 0.2, 500)):
        synth[f"synth_w{width}"] = synth_meme(rng, width, alpha, nsites, f"SYN{width:02d}.1")
    fixtures.update({k + "_meme": v for k, v in synth.items()})
    # motifs wider than one 64-bit packed word (JASPAR 2022 CORE holds 34/35-bp CTCF profiles): own generator, so that
    # the vectors above do not change
    rng_wide = np.random.default_rng(20244)
    synth_wide = {}
    for width, alpha, nsites in ((33, 0.3, 900), (35, 0.2, 2000), (48, 0.5, 400), (64, 0.3, 1500)):
        synth_wide[f"synth_w{width}"] = synth_meme(rng_wide, width, alpha, nsites, f"SYN{width:02d}.1")
    fixtures.update({k + "_meme": v for k, v in synth_wide.items()})
    with open(os.path.join(HERE, "fixtures.json"), "w") as fh:
        json.dump(fixtures, fh, indent=0, sort_keys=True)

    def write_tmp(name, text):
        p = os.path.join(scratch, name)
        with open(p, "w") as fh:
            fh.write(text)
        return p

    # ---- (B) motif goldens ---------------------------------------------------------------------
    motifs = {}
    ncpu = 1

    def add_meme(tag, text, bg, norev):
        m = build_motif_meme(write_tmp(tag + ".meme", text), bg, 0.1, norev, ncpu, False, True)[0]
        motifs[tag] = m
        np.savez_compressed(os.path.join(CASES, f"motif_{tag}.npz"), **motif_bundle(m, dict(
            source=ustr([tag.split("__")[0]]), fmt=ustr(["meme"]), bgfile=ustr(["unif" if bg == UNIF else "bg_nt"]),
            no_reverse=np.bool_(norev), pseudo=np.float64(0.1))))

    sys_stdout = sys.stdout
    sys.stdout = io.StringIO()
    try:
        add_meme("ctcf_meme__unif", fixtures["ctcf_meme"], UNIF, False)
        add_meme("ctcf_meme__bgnt", fixtures["ctcf_meme"], bg_nt, False)
        add_meme("ctcf_meme__bgnt_norev", fixtures["ctcf_meme"], bg_nt, True)
        add_meme("ctcf_meme__unif_norev", fixtures["ctcf_meme"], UNIF, True)
        add_meme("atf3_meme__bgnt", fixtures["atf3_meme"], bg_nt, False)
        add_meme("gata1_meme__unif", fixtures["gata1_meme"], UNIF, False)
        for k in synth:
            add_meme(f"{k}_meme__bgnt", fixtures[k + "_meme"], bg_nt, False)
        add_meme("synth_w8_meme__unif", fixtures["synth_w8_meme"], UNIF, False)
        add_meme("synth_w30_meme__bgnt_norev", fixtures["synth_w30_meme"], bg_nt, True)
        for k in synth_wide:
            add_meme(f"{k}_meme__bgnt", fixtures[k + "_meme"], bg_nt, False)
        add_meme("synth_w35_meme__unif_norev", fixtures["synth_w35_meme"], UNIF, True)
        for fmt, fn in (("jaspar", build_motif_jaspar), ("transfac", build_motif_transfac), ("pfm", build_motif_pfm)):
            for bgtag, bg in (("unif", UNIF), ("bgnt", bg_nt)):
                tag = f"ctcf_{fmt}__{bgtag}"
                m = fn(write_tmp(f"MA0139.1.{fmt}", fixtures[f"ctcf_{fmt}"]), bg, 0.1, False, False, True)
                motifs[tag] = m
                np.savez_compressed(os.path.join(CASES, f"motif_{tag}.npz"), **motif_bundle(m, dict(
                    source=ustr([f"ctcf_{fmt}"]), fmt=ustr([fmt]), bgfile=ustr([bgtag]),
                    no_reverse=np.bool_(False), pseudo=np.float64(0.1))))
    finally:
        sys.stdout = sys_stdout
    print("motif goldens:", len(motifs))

    # ---- (C) scoring goldens -------------------------------------------------------------------
    fixture_lines = fixtures["scoring_input_tsv"].strip("\n").split("\n")
    rng = np.random.default_rng(20243)
    extra_n = synth_rows(rng, 40, 19, chrom="22", region_start=19723256, with_n=12, lower=8)
    rows_w8 = synth_rows(rng, 3000, 8, chrom="7", with_n=5)
    rows_w30 = synth_rows(rng, 2000, 30, chrom="X", with_n=3)
    rows_w32 = synth_rows(rng, 500, 32, chrom="3", with_n=2)
    rows_w6 = synth_rows(rng, 1500, 6, chrom="1", with_n=0)

    scoring = []

    def add_scoring(tag, motif_tag, files, width, testmode=False, **opts):
        motif = motifs[motif_tag]
        wf = None if testmode else make_findmotif(Findmotif, **opts)
        df, out = run_scoring(compute_results, motif, files, width, scratch, tag, wf, testmode)
        b = df_bundle(df)
        b["motif_tag"] = ustr([motif_tag])
        b["n_files"] = np.int64(len(files))
        for k, lines in enumerate(files):
            b[f"file_{k}"] = ustr(lines)
        o = dict(cores=1, threshold=1.0 if testmode else opts.get("threshold", 1e-4),
                 noqvalue=opts.get("noqvalue", False), qvalueT=opts.get("qvalueT", False),
                 noreverse=opts.get("noreverse", False), recomb=True if testmode else opts.get("recomb", False))
        b["options_json"] = ustr([json.dumps(o, sort_keys=True)])
        b["stdout"] = ustr([out])
        # the reference GFF3 writer on the first rows (layout golden)
        if len(df) > 0:
            cwd = os.getcwd()
            os.chdir(scratch)
            try:
                writeGFF3("gff_" + tag, df.head(25), o["noqvalue"], True)
                b["gff3_head25"] = ustr([read_text(os.path.join(scratch, "gff_" + tag + ".gff"))])
                buf = io.StringIO()
                df.head(25).to_csv(buf, sep="\t", encoding="utf-8")
                b["tsv_head25"] = ustr([buf.getvalue()])
            finally:
                os.chdir(cwd)
        np.savez_compressed(os.path.join(CASES, f"scoring_{tag}.npz"), **b)
        scoring.append((tag, len(df)))

    add_scoring("fixture_testmode", "ctcf_meme__unif", [fixture_lines], 19, testmode=True)
    add_scoring("fixture_bgnt_t1", "ctcf_meme__bgnt", [fixture_lines], 19, threshold=1.0, recomb=True)
    add_scoring("fixture_t05_norecomb", "ctcf_meme__unif", [fixture_lines], 19, threshold=0.05, recomb=False)
    add_scoring("fixture_default", "ctcf_meme__unif", [fixture_lines], 19, threshold=1e-2)
    add_scoring("fixture_qvalT", "ctcf_meme__bgnt", [fixture_lines], 19, threshold=0.9, qvalueT=True, recomb=True)
    add_scoring("fixture_norev", "ctcf_meme__bgnt_norev", [fixture_lines], 19, threshold=0.5, noreverse=True,
                recomb=True)
    add_scoring("fixture_noq", "ctcf_meme__unif", [fixture_lines], 19, threshold=0.2, noqvalue=True, recomb=True)
    add_scoring("fixture_plus_N_2files", "ctcf_meme__bgnt", [fixture_lines, extra_n], 19, threshold=1.0, recomb=True)
    add_scoring("synth_w8", "synth_w8_meme__bgnt", [rows_w8[:3500], rows_w8[3500:]], 8, threshold=0.05, recomb=True)
    add_scoring("synth_w8_unif_t1", "synth_w8_meme__unif", [rows_w8], 8, threshold=1.0, recomb=False)
    add_scoring("synth_w30", "synth_w30_meme__bgnt", [rows_w30], 30, threshold=1.0, recomb=True)
    add_scoring("synth_w30_norev", "synth_w30_meme__bgnt_norev", [rows_w30], 30, threshold=0.3, noreverse=True,
                recomb=True)
    add_scoring("synth_w32", "synth_w32_meme__bgnt", [rows_w32], 32, threshold=1.0, recomb=True)
    add_scoring("synth_w6", "synth_w6_meme__bgnt", [rows_w6], 6, threshold=1.0, recomb=True)
    rng_wide = np.random.default_rng(20245)
    rows_w33 = synth_rows(rng_wide, 1200, 33, chrom="5", with_n=4)
    rows_w35 = synth_rows(rng_wide, 2500, 35, chrom="11", with_n=6, lower=5)
    rows_w48 = synth_rows(rng_wide, 800, 48, chrom="2", with_n=2)
    rows_w64 = synth_rows(rng_wide, 600, 64, chrom="X", with_n=3)
    add_scoring("synth_w33", "synth_w33_meme__bgnt", [rows_w33], 33, threshold=1.0, recomb=True)
    add_scoring("synth_w35", "synth_w35_meme__bgnt", [rows_w35[:2000], rows_w35[2000:]], 35, threshold=0.05, recomb=False)
    add_scoring("synth_w35_norev", "synth_w35_meme__unif_norev", [rows_w35], 35, threshold=0.5, noreverse=True, recomb=True)
    add_scoring("synth_w48", "synth_w48_meme__bgnt", [rows_w48], 48, threshold=1.0, recomb=True)
    add_scoring("synth_w64", "synth_w64_meme__bgnt", [rows_w64], 64, threshold=1.0, qvalueT=True, recomb=True)
    print("scoring goldens:", scoring)
    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
