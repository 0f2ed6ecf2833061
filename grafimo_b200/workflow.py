"""Argument container of `grafimo findmotif` for the motif-scanning path.

Keeps the attribute names the path reads from the reference's `Findmotif` workflow object
(src/grafimo/workflow.py:233-632; read at src/grafimo/score_sequences.py:93-99, src/grafimo/motif_ops.py:1161-1176
and src/grafimo/res_writer.py:98-101).  Graph construction / k-mer extraction arguments are not part of this path:
the k-mers come from a directory of `vg find` TSVs (`kmers_dir`), or -- SURVEY.md 8f-1 -- from a variation graph
built on the fly from the inputs of `grafimo buildvg` (-l/--linear-genome FASTA, -v/--vcf VCF; src/grafimo/__main__.py:
198-217) and scanned on the GPU over the regions of -b/--bedfile.
"""
from .utils import DEFAULT_OUTDIR, UNIF


class Findmotif(object):
    def __init__(self, motif=None, kmers_dir="", bgfile=UNIF, pseudo=0.1, threshold=1e-4, out=DEFAULT_OUTDIR, cores=1,
                 recomb=False, top_graphs=0, no_qvalue=False, no_reverse=False, text_only=False, qval_t=False,
                 verbose=False, gpus=1, linear_genome="", vcf="", bedfile="", chroms_prefix=""):
        def expect(value, kind, name):
            if not isinstance(value, kind):
                raise TypeError(f"\n\nERROR: commandline parsing failed. Type mismatch: expected {kind.__name__}, "
                                f"got {type(value).__name__} instance ({name}).\n")
        motif = [] if motif is None else motif
        expect(motif, list, "motif"); expect(kmers_dir, str, "kmers_dir"); expect(bgfile, str, "bgfile")
        expect(pseudo, float, "pseudo"); expect(threshold, float, "threshold"); expect(out, str, "out")
        expect(cores, int, "cores"); expect(recomb, bool, "recomb"); expect(top_graphs, int, "top_graphs")
        expect(no_qvalue, bool, "no_qvalue"); expect(no_reverse, bool, "no_reverse"); expect(text_only, bool, "text_only")
        expect(qval_t, bool, "qval_t"); expect(verbose, bool, "verbose"); expect(gpus, int, "gpus")
        expect(linear_genome, str, "linear_genome"); expect(vcf, str, "vcf"); expect(bedfile, str, "bedfile")
        expect(chroms_prefix, str, "chroms_prefix")
        if bool(linear_genome) != bool(bedfile) or (vcf and not linear_genome):
            raise ValueError("\n\nERROR: scanning a graph built on the fly needs -l/--linear-genome and -b/--bedfile "
                             "(and -v/--vcf for the variants).\n")
        if gpus < 1:
            raise ValueError("\n\nERROR: --gpus must be at least 1.\n")
        if not (0 < threshold <= 1):
            raise ValueError("\n\nERROR: the threshold must be in (0, 1].\n")
        if qval_t and no_qvalue:
            raise ValueError("\n\nERROR: unable to apply the threshold on q-values if they are not computed.\n")
        self._motif, self._kmers_dir, self._bgfile, self._pseudo = motif, kmers_dir, bgfile, pseudo
        self._thresh, self._outdir, self._cores, self._recomb = threshold, out, cores, recomb
        self._top_graphs, self._no_qvalue, self._no_rev, self._text_only = top_graphs, no_qvalue, no_reverse, text_only
        self._qvalueT, self._verbose, self._gpus = qval_t, verbose, gpus
        self._linear_genome, self._vcf, self._bedfile, self._chroms_prefix = linear_genome, vcf, bedfile, chroms_prefix

    motif = property(lambda self: self._motif)
    kmers_dir = property(lambda self: self._kmers_dir)
    bgfile = property(lambda self: self._bgfile)
    pseudo = property(lambda self: self._pseudo)
    threshold = property(lambda self: self._thresh)
    outdir = property(lambda self: self._outdir)
    cores = property(lambda self: self._cores)
    recomb = property(lambda self: self._recomb)
    top_graphs = property(lambda self: self._top_graphs)
    noqvalue = property(lambda self: self._no_qvalue)
    noreverse = property(lambda self: self._no_rev)
    text_only = property(lambda self: self._text_only)
    qvalueT = property(lambda self: self._qvalueT)
    verbose = property(lambda self: self._verbose)
    gpus = property(lambda self: self._gpus)
    linear_genome = property(lambda self: self._linear_genome)
    vcf = property(lambda self: self._vcf)
    bedfile = property(lambda self: self._bedfile)
    chroms_prefix = property(lambda self: self._chroms_prefix)

    def has_graph_inputs(self):
        return bool(self._linear_genome)
