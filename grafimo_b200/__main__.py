"""`python -m grafimo_b200 findmotif ...` -- the findmotif command line for the motif-scanning path.

Keeps the findmotif flags that drive the path in the reference (src/grafimo/__main__.py:119-413): -m/--motif,
-k/--bgfile, -p/--pseudo, -t/--threshold, -q/--no-qvalue, -r/--no-reverse, -f/--text-only, --recomb, --qvalueT,
-o/--out, -j/--cores, --verbose, --debug.  The variation-graph arguments (-g/-d/-b, `vg` subprocesses) are
replaced by --kmers-dir: the directory `scan_graph` would have produced (`width_<w>/*.tsv` from
`vg find -K w -E`), because this implementation drops in after k-mer extraction.
"""
import argparse
import sys

from .motif_ops import get_motif_pwm
from .res_writer import print_results, write_results
from .score_sequences import compute_results
from .utils import DEFAULT_OUTDIR, UNIF
from .workflow import Findmotif


def get_parser():
    p = argparse.ArgumentParser(prog="grafimo_b200", description="B200-native GRAFIMO motif scanning")
    sub = p.add_subparsers(dest="workflow")
    f = sub.add_parser("findmotif", help="scan pre-extracted variation-graph k-mers for motif occurrences")
    f.add_argument("-m", "--motif", nargs="+", required=True, metavar="MOTIF-FILE")
    f.add_argument("--kmers-dir", required=True, metavar="DIR", help="directory holding width_<w>/*.tsv k-mer files")
    f.add_argument("-k", "--bgfile", default=UNIF)
    f.add_argument("-p", "--pseudo", type=float, default=0.1)
    f.add_argument("-t", "--threshold", type=float, default=1e-4)
    f.add_argument("-q", "--no-qvalue", action="store_true", default=False, dest="no_qvalue")
    f.add_argument("-r", "--no-reverse", action="store_true", default=False, dest="no_reverse")
    f.add_argument("-f", "--text-only", action="store_true", default=False, dest="text_only")
    f.add_argument("--recomb", action="store_true", default=False)
    f.add_argument("--qvalueT", action="store_true", default=False, dest="qval_t")
    f.add_argument("-o", "--out", default=DEFAULT_OUTDIR)
    f.add_argument("-j", "--cores", type=int, default=1)
    f.add_argument("--verbose", action="store_true", default=False)
    f.add_argument("--debug", action="store_true", default=False)
    return p


def findmotif(wf: Findmotif, debug: bool) -> None:
    """The orchestration of src/grafimo/grafimo.py:80-190 minus scan_graph."""
    motifs = []
    for mf in wf.motif:
        motifs += get_motif_pwm(mf, wf, wf.cores, debug)
    for motif in motifs:
        res = compute_results(motif, wf.kmers_dir, debug, wf)
        if wf.text_only:
            print_results(res, debug)
        else:
            write_results(res, motif, len(motifs), wf, debug)


def main(argv=None):
    args = get_parser().parse_args(argv)
    if args.workflow != "findmotif":
        get_parser().print_help()
        return 1
    wf = Findmotif(motif=args.motif, kmers_dir=args.kmers_dir, bgfile=args.bgfile, pseudo=args.pseudo,
                   threshold=args.threshold, out=args.out, cores=args.cores, recomb=args.recomb,
                   no_qvalue=args.no_qvalue, no_reverse=args.no_reverse, text_only=args.text_only, qval_t=args.qval_t,
                   verbose=args.verbose)
    findmotif(wf, args.debug)
    return 0


if __name__ == "__main__":
    sys.exit(main())
