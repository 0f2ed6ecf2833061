"""Report writers: TSV, HTML and GFF3 with the reference's byte layout.

`write_results` / `writeGFF3` / `print_results` follow src/grafimo/res_writer.py:41-208,213-303,415-437.
The GFF3 attribute string reproduces the reference's quirks (`pvalue==`, `sequence==...=;`) because
downstream tooling parses it as is.  The `--top-graphs` PNG rendering needs the external `vg` and `dot`
binaries and is outside the accelerated path: requesting it raises.
"""
import os
import time

import numpy as np
import pandas as pd

from .grafimo_errors import FileWriteError, VGError
from .motif import Motif
from .utils import DEFAULT_OUTDIR, PHASE, SOURCE, TP, dftolist, exception_handler


def _format_unique(values, fmt):
    """Formats a float column through its distinct values: p, q and score take one value per score bin, so a
    report of millions of rows has only a few thousand distinct numbers to format."""
    arr = np.asarray(values, dtype=np.float64)
    uniq, inv = np.unique(arr, return_inverse=True)
    return np.array([fmt(u) for u in uniq.tolist()], dtype=object)[inv]


def gff3_lines(data: pd.DataFrame, no_qvalue: bool, debug: bool = False):
    """The GFF3 body lines (src/grafimo/res_writer.py:262-298), built column-wise."""
    cols = dftolist(data, no_qvalue, debug)
    if not no_qvalue and len(cols) != 12:
        exception_handler(ValueError, "Q-values columns seems to be missing.\n", debug)
    motif_ids, motif_names, seqnames, starts, stops, strands, scores, pvalues, seqs, _freqs, refs = cols[:11]
    score_s = _format_unique(scores, lambda v: str(round(v, 1)))
    p_s = _format_unique(pvalues, lambda v: str(np.format_float_scientific(v, exp_digits=2)))
    q_s = _format_unique(cols[11], lambda v: str(np.format_float_scientific(v, exp_digits=2))) if not no_qvalue else None
    out = []
    for i in range(len(seqnames)):
        seqname, strand = seqnames[i], strands[i]
        # '-' rows carry start > stop; GFF3 wants forward coordinates
        first, second = (stops[i], starts[i]) if strand == "-" else (starts[i], stops[i])
        qpart = f"qvalue={q_s[i]};" if q_s is not None else ""
        out.append(f"{seqname.split(':')[0]}\t{SOURCE}\t{TP}\t{first}\t{second}\t{score_s[i]}\t{strand}\t{PHASE}\t"
                   f"Name={motif_ids[i]}_{seqname}{strand}:{refs[i]};Alias={motif_names[i]};"
                   f"ID={motif_ids[i]}=-={motif_names[i]}=-={seqname};pvalue=={p_s[i]};{qpart}sequence=={seqs[i]}=;\n")
    return out


def writeGFF3(prefix: str, data: pd.DataFrame, no_qvalue: bool, debug: bool) -> None:
    if not isinstance(prefix, str):
        exception_handler(TypeError, f"Expected str, got {type(prefix).__name__}.\n", debug)
    if not isinstance(data, pd.DataFrame):
        exception_handler(TypeError, f"Expected DataFrame, got {type(data).__name__}.\n", debug)
    if not isinstance(no_qvalue, bool):
        exception_handler(TypeError, f"Expected bool, got {type(no_qvalue).__name__}.\n", debug)
    gfffn = ".".join([prefix, "gff"])
    try:
        with open(gfffn, mode="w+") as out:
            out.write("##gff-version 3\n")
            for line in gff3_lines(data, no_qvalue, debug):
                out.write(line)
    except OSError:
        exception_handler(FileWriteError, f"An error ocurred while writing {gfffn}.\n", debug)


def write_results(results: pd.DataFrame, motif: Motif, motif_num: int, args_obj, debug: bool) -> None:
    """TSV (`DataFrame.to_csv(sep="\\t")`, index column included), HTML and GFF3 in the output directory
    (default `grafimo_out_<PID>_<motifID>`, src/grafimo/res_writer.py:103-148)."""
    if not isinstance(results, pd.DataFrame):
        exception_handler(TypeError, f"Expected DataFrame, got {type(results).__name__}.\n", debug)
    if len(results) == 0:
        exception_handler(ValueError, "No potential motif occurrence retreived.\n", debug)
    if not isinstance(motif, Motif):
        exception_handler(TypeError, f"Expected Motif, got {type(motif).__name__}.\n", debug)
    if not isinstance(motif_num, int) or motif_num <= 0:
        exception_handler(ValueError, "No motif searched. Probably something went wrong.\n", debug)
    outdir, no_qvalue, verbose = args_obj.outdir, args_obj.noqvalue, args_obj.verbose
    if getattr(args_obj, "top_graphs", 0) > 0:
        exception_handler(VGError, "--top-graphs needs the external vg and dot binaries and is not part of the "
                          "B200 motif-scanning path.\n", debug)
    default_name = outdir == DEFAULT_OUTDIR
    if default_name:
        outdir = "_".join(["grafimo_out", str(os.getpid()), motif.motif_id])
    os.makedirs(outdir, exist_ok=True)
    print(f"\nWriting results in {outdir}.\n")
    prefix = "_".join(["grafimo_out", motif.motif_id]) if (not default_name and motif_num > 1) else "grafimo_out"
    base = os.path.join(outdir, prefix)
    t0 = time.time()
    results.to_csv(base + ".tsv", sep="\t", encoding="utf-8")
    if verbose:
        print("%s.tsv written in %.2fs" % (prefix, time.time() - t0))
    if not getattr(args_obj, "text_only", False):
        t0 = time.time()
        results.to_html(base + ".html")
        if verbose:
            print("%s.html written in %.2fs" % (prefix, time.time() - t0))
    t0 = time.time()
    writeGFF3(base, results, no_qvalue, debug)
    if verbose:
        print("%s.gff written in %.2fs" % (prefix, time.time() - t0))


def print_results(results: pd.DataFrame, debug: bool) -> None:
    """--text-only: the table on stdout (src/grafimo/res_writer.py:415-437)."""
    if not isinstance(results, pd.DataFrame):
        exception_handler(TypeError, f"Expected DataFrame, got {type(results).__name__}.\n", debug)
    pd.set_option("display.max_rows", len(results))
    print()
    print(results)
    pd.reset_option("display.max_rows")
