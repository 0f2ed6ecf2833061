"""In-tree build of libgrafimo_b200.so (hand-written sm_100a CUDA kernels + the C ABI).

    python -m grafimo_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The library is written next to this file so that it travels to
the GPU box with the repository snapshot (it is git-ignored, not gpurun-ignored).
"""
import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libgrafimo_b200.so")
SOURCES = ["context.cu", "encode.cu", "tsv.cu", "score.cu", "score_wide.cu", "pval.cu", "qvalue.cu", "dense_sort.cu", "scan_host.cu", "seqscan.cu", "graph.cu", "graph_build.cu", "vcf.cu", "report.cu", "comm.cu", "host_pack.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--fmad=true", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: grafimo_b200 has no CPU fallback and cannot be built without CUDA")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "internal.cuh"), os.path.join(CSRC, "score_common.cuh"), os.path.join(CSRC, "scan_tail.cuh"), os.path.join(ROOT, "include", "grafimo_b200.h"), os.path.abspath(__file__)]
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc()] + ARCH + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            for out in ex.map(run, jobs):
                if verbose and out:
                    print(out)
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl", "-lpthread"]
        run(cmd)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
