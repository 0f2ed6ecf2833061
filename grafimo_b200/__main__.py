"""`python -m grafimo_b200 findmotif ...` -- the findmotif command line for the motif-scanning path.

Keeps the findmotif flags that drive the path in the reference (src/grafimo/__main__.py:119-413): -m/--motif,
-k/--bgfile, -p/--pseudo, -t/--threshold, -q/--no-qvalue, -r/--no-reverse, -f/--text-only, --recomb, --qvalueT,
-o/--out, -j/--cores, --verbose, --debug.  The variation-graph arguments (-g/-d/-b, `vg` subprocesses) are
replaced by --kmers-dir: the directory `scan_graph` would have produced (`width_<w>/*.tsv` from
`vg find -K w -E`), because this implementation drops in after k-mer extraction -- or by -l/-v/-b (the inputs of
`grafimo buildvg` plus the BED file): then the graph is built here and its k-mers are extracted and scored on the GPU
without `vg` and without any text in between.
"""
import argparse
import sys

from .motif_ops import get_motif_pwm
from .res_writer import print_results, write_results, write_results_device
from .score_sequences import compute_results, scan_dir_device, scan_rows_device
from .utils import DEFAULT_OUTDIR, UNIF
from .workflow import Findmotif


def get_parser():
    p = argparse.ArgumentParser(prog="grafimo_b200", description="B200-native GRAFIMO motif scanning")
    sub = p.add_subparsers(dest="workflow")
    f = sub.add_parser("findmotif", help="scan pre-extracted variation-graph k-mers for motif occurrences")
    f.add_argument("-m", "--motif", nargs="+", required=True, metavar="MOTIF-FILE")
    f.add_argument("--kmers-dir", default="", metavar="DIR", help="directory holding width_<w>/*.tsv k-mer files")
    f.add_argument("-l", "--linear-genome", default="", metavar="REFERENCE-FASTA", dest="linear_genome",
                   help="reference genome: with -v and -b the variation graph is built and scanned on the GPU")
    f.add_argument("-v", "--vcf", default="", metavar="VCF", help="phased variants (VCF or VCF.gz)")
    f.add_argument("-b", "--bedfile", default="", metavar="BEDFILE", help="regions to scan (UCSC BED)")
    f.add_argument("--chroms-prefix-find", default="", dest="chroms_prefix", metavar="PREFIX",
                   help="prefix of the sequence names in the FASTA/VCF (e.g. chr)")
    f.add_argument("-k", "--bgfile", default=UNIF)
    f.add_argument("-p", "--pseudo", type=float, default=0.1)
    f.add_argument("-t", "--threshold", type=float, default=1e-4)
    f.add_argument("-q", "--no-qvalue", action="store_true", default=False, dest="no_qvalue")
    f.add_argument("-r", "--no-reverse", action="store_true", default=False, dest="no_reverse")
    f.add_argument("-f", "--text-only", action="store_true", default=False, dest="text_only")
    f.add_argument("--recomb", action="store_true", default=False)
    f.add_argument("--qvalueT", action="store_true", default=False, dest="qval_t")
    f.add_argument("-o", "--out", default=DEFAULT_OUTDIR)
    f.add_argument("-j", "--cores", type=int, default=1,
                   help="host threads for the motif files (the scan itself runs on the GPU whatever this says)")
    f.add_argument("--gpus", type=int, default=1, metavar="N",
                   help="GPUs of this node to use (one process per GPU, launched here with torch.distributed.run): a motif "
                        "collection is sharded by motif, a single motif by chromosome / k-mer file with one all-reduce of "
                        "the score histogram so that the q-values stay global")
    f.add_argument("--verbose", action="store_true", default=False)
    f.add_argument("--debug", action="store_true", default=False)
    return p


def findmotif(wf: Findmotif, debug: bool) -> None:
    """The orchestration of src/grafimo/grafimo.py:80-190 minus scan_graph.  The reference scores one motif after the
    other (grafimo.py:177-183); here a motif collection over a graph goes through ONE many-motif scan
    (score_sequences.scan_rows_device_many), and with --gpus N the collection is split over the GPUs by motif (greedy by
    width x k-mers; every rank writes the reports of its own motifs) -- a single motif is split by rows instead and
    the ranks exchange the score histogram (compute_results / compute_results_rows)."""
    from . import dist as gdist
    from . import score_sequences as ss
    motifs = []
    for mf in wf.motif:
        motifs += get_motif_pwm(mf, wf, wf.cores, debug)
    n_all = len(motifs)
    world, rank = ss._dist_world()
    by_motif = world > 1 and n_all > 1
    graphs = load_graphs(wf, debug) if wf.has_graph_inputs() else None
    if by_motif:  # shard the collection: longest scans first, always to the least loaded rank
        mine = gdist.assign_chromosomes([m.width for m in motifs], world)[rank]
        motifs = [motifs[i] for i in mine]
    import contextlib
    with (ss.local_only() if by_motif else contextlib.nullcontext()):
        if graphs is not None:  # scan_graph + compute_results + writers without text or a DataFrame in between
            if world > 1 and not by_motif:  # one motif, many GPUs: chromosomes over the ranks, global q-values
                graphs = [graphs[i] for i in gdist.assign_chromosomes([sum(b - a for a, b in sp) for _, sp in graphs], world)[rank]]
                for motif in motifs:
                    rows = [dg.extract(spans, motif.width) for dg, spans in graphs]
                    res = ss.compute_results_rows(motif, rows, debug, wf)
                    if rank == 0:
                        print_results(res, debug) if wf.text_only else write_results(res, motif, n_all, wf, debug)
                return
            # one k-mer set per distinct motif width, like scan_graph (src/grafimo/extract_regions.py:131-134)
            rows_of_width = {w: [dg.extract(spans, w) for dg, spans in graphs] for w in sorted({m.width for m in motifs})}
            for motif, report in zip(motifs, ss.scan_rows_device_many(motifs, rows_of_width, debug, wf)):
                if wf.text_only:
                    print_results(report.to_df(), debug)
                else:
                    write_results_device(report, motif, n_all, wf, debug)
            return
        for motif in motifs:
            if not wf.text_only:  # files straight from the device columns (K8) when the input allows it
                report = scan_dir_device(motif, wf.kmers_dir, debug, wf)
                if report is not None:
                    write_results_device(report, motif, n_all, wf, debug)
                    continue
            res = compute_results(motif, wf.kmers_dir, debug, wf)
            if world > 1 and not by_motif and rank != 0:
                continue  # every rank holds the merged table; rank 0 writes it
            if wf.text_only:
                print_results(res, debug)
            else:
                write_results(res, motif, n_all, wf, debug)


def load_graphs(wf: Findmotif, debug: bool):
    """One device-resident graph per BED chromosome (what `grafimo buildvg` + scan_graph set up with vg,
    src/grafimo/grafimo.py:32-78, extract_regions.py:55-237).  Regions are named like the reference names them:
    `<chromosome without prefix>:<start>-<stop>` (extract_regions.py:164-170)."""
    from . import score_sequences as ss
    from .extract_regions import DeviceGraph, get_regions_bed
    from .utils import exception_handler
    from .vgraph import read_fasta
    regions, n = get_regions_bed(wf.bedfile, debug)
    if n == 0:
        exception_handler(ValueError, f"No region found in {wf.bedfile}.\n", debug)
    seqs = read_fasta(wf.linear_genome)
    ctx = ss._context()
    out = []
    parsed = None
    if wf.vcf and len(regions) > 1:  # one pass over the VCF for all BED chromosomes, not one per chromosome
        from .vgraph import read_vcf_device
        parsed, _ = read_vcf_device(ctx, wf.vcf, None, by_chrom=True)
    for chrom, spans in regions.items():
        key = chrom[3:] if chrom.startswith("chr") else chrom
        name = wf.chroms_prefix + key
        if name not in seqs:
            exception_handler(KeyError, f"{name} is not a sequence of {wf.linear_genome}. Consider --chroms-prefix-find.\n", debug)
        dg = DeviceGraph.from_files(ctx, seqs, wf.vcf, name, display_name=key, parsed_vcf=parsed)
        out.append((dg, [(int(a), int(b)) for a, b in spans]))
    return out


def _launch_ranks(gpus: int, argv) -> int:
    """`--gpus N` from a plain command line: the same command under torch.distributed.run, one process per GPU."""
    import os
    import subprocess
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={gpus}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + os.getpid() % 400), "-m", "grafimo_b200"] + list(argv)
    return subprocess.call(cmd)


def main(argv=None):
    import os
    argv = list(sys.argv[1:] if argv is None else argv)
    args = get_parser().parse_args(argv)
    if args.workflow != "findmotif":
        get_parser().print_help()
        return 1
    wf = Findmotif(motif=args.motif, kmers_dir=args.kmers_dir, bgfile=args.bgfile, pseudo=args.pseudo,
                   threshold=args.threshold, out=args.out, cores=args.cores, recomb=args.recomb,
                   no_qvalue=args.no_qvalue, no_reverse=args.no_reverse, text_only=args.text_only, qval_t=args.qval_t,
                   verbose=args.verbose, gpus=args.gpus, linear_genome=args.linear_genome, vcf=args.vcf, bedfile=args.bedfile,
                   chroms_prefix=args.chroms_prefix)
    if not wf.kmers_dir and not wf.has_graph_inputs():
        get_parser().error("give --kmers-dir, or -l/--linear-genome with -b/--bedfile (and -v/--vcf)")
    if wf.gpus > 1 and "RANK" not in os.environ:
        return _launch_ranks(wf.gpus, argv)
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:  # one of the ranks of a --gpus N run
        import torch
        from . import dist as gdist
        from . import score_sequences as ss
        info = gdist.init_from_env("nccl")
        torch.cuda.set_device(info["local"])
        gdist.init_comm(ss._context())  # the library's own communicator: the histogram all-reduce runs inside the C ABI
    findmotif(wf, args.debug)
    return 0


if __name__ == "__main__":
    sys.exit(main())
