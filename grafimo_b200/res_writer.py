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
from .motif import Motif, is_motif
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
    if not is_motif(motif):
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


# ---------------------------------------------------------------------------------------------------------
# K8: the same two files written from device-resident hit columns (csrc/report.cu), no DataFrame in between
# ---------------------------------------------------------------------------------------------------------
HTML_ROW_LIMIT = 200_000  # write_results_device leaves the HTML table out above this many rows


def _tsv_float_strings(values):
    """Text pandas' `to_csv` writes for these float64 values (the reference's TSV writer): asked from pandas itself."""
    import io
    if len(values) == 0:
        return []
    buf = io.StringIO()
    pd.DataFrame({"v": np.asarray(values, dtype=np.float64)}).to_csv(buf, sep="\t", header=False, index=False)
    out = buf.getvalue().split("\n")[:len(values)]
    return out


class DeviceReport:
    """Hit columns on the device + everything K8 needs to print them.  Built by score_sequences.scan_rows_device."""

    def __init__(self, ctx, motif, width, want_q, kmer, strand, start, stop, freq, ref, bin_, name, seqnames, score_by_bin,
                 p_by_bin, q_by_bin):
        self.ctx, self.motif, self.width, self.want_q = ctx, motif, int(width), bool(want_q)
        self.kmer, self.strand, self.start, self.stop, self.freq, self.ref, self.bin, self.name = (
            kmer, strand, start, stop, freq, ref, bin_, name)
        self.seqnames = list(seqnames)
        self.score_by_bin, self.p_by_bin, self.q_by_bin = score_by_bin, p_by_bin, q_by_bin
        self.n = int(kmer.shape[0])

    def _tables(self, layout):
        sci = lambda v: str(np.format_float_scientific(v, exp_digits=2))  # noqa: E731
        if layout == 0:
            score = _tsv_float_strings(self.score_by_bin)
            p = _tsv_float_strings(self.p_by_bin)
            q = _tsv_float_strings(self.q_by_bin) if self.want_q else []
        else:
            score = [str(round(v, 1)) for v in np.asarray(self.score_by_bin, dtype=np.float64).tolist()]
            p = [sci(v) for v in np.asarray(self.p_by_bin, dtype=np.float64).tolist()]
            q = [sci(v) for v in np.asarray(self.q_by_bin, dtype=np.float64).tolist()] if self.want_q else []
        chroms = [s.split(":")[0] for s in self.seqnames]
        consts = [self.motif.motif_id, self.motif.motif_name, "ref", "non.ref", f"\t{SOURCE}\t{TP}\t", f"\t{PHASE}\tName=", ";Alias=",
                  ";ID=", "=-=", ";pvalue==", ";qvalue=", ";sequence==", "=;\n"]
        tables = [score, p, q, self.seqnames, chroms, consts]
        first, strings = [], []
        for t in tables:
            first.append(len(strings))
            strings.extend(t)
        enc = [s.encode("utf-8") for s in strings]
        off = np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.uint32)
        blob = np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8)
        return blob, off, first

    def render(self, layout):
        """-> uint8 array (a bytes-like view of pinned memory) with the body rows of the TSV (layout 0) or GFF3
        (layout 1) file."""
        import ctypes

        import torch

        from ._lib import Report, check
        ctx = self.ctx
        if self.n == 0:
            return np.zeros(0, dtype=np.uint8)
        blob, off, first = self._tables(layout)
        with torch.cuda.stream(ctx.stream):
            d_blob = torch.from_numpy(blob.copy()).to(ctx.device)
            d_off = torch.from_numpy(off.astype(np.int32)).to(ctx.device)
            row_off = torch.empty(self.n + 1, dtype=torch.int64, device=ctx.device)
        r = Report()
        r.n_rows, r.index_base, r.width, r.layout, r.want_q = self.n, 0, self.width, int(layout), int(self.want_q)
        for k, t in (("d_kmer", self.kmer), ("d_strand", self.strand), ("d_start", self.start), ("d_stop", self.stop),
                     ("d_freq", self.freq), ("d_ref", self.ref), ("d_bin", self.bin), ("d_name", self.name), ("d_strings", d_blob),
                     ("d_string_off", d_off)):
            setattr(r, k, t.data_ptr())
        r.first_score, r.first_p, r.first_q, r.first_name, r.first_chrom, r.first_const = first
        total = ctypes.c_uint64(0)
        ctx.enter()
        check(ctx.lib.gb2_report_measure(ctx.h, ctypes.byref(r), ctypes.c_void_p(row_off.data_ptr()), ctypes.byref(total)),
              "gb2_report_measure", ctx.h)
        nbytes = int(total.value)
        with torch.cuda.stream(ctx.stream):
            out = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=ctx.device)
        check(ctx.lib.gb2_report_write(ctx.h, ctypes.byref(r), ctypes.c_void_p(row_off.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                       nbytes), "gb2_report_write", ctx.h)
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)  # pinned: the copy back runs at PCIe speed
        with torch.cuda.stream(ctx.stream):
            host.copy_(out[:nbytes], non_blocking=True)
        ctx.sync()
        ctx.leave()
        return host.numpy()

    def to_df(self) -> pd.DataFrame:
        """The same rows as a DataFrame with the reference's columns (resultsTmp.py:269-301), in the device's order."""
        import torch
        with torch.cuda.stream(self.ctx.stream):
            h = {k: getattr(self, k).cpu().numpy() for k in ("kmer", "strand", "start", "stop", "freq", "ref", "bin", "name")}
        self.ctx.sync()
        n, w = self.n, self.width
        from .extract_regions import decode_kmers
        letters = decode_kmers(h["kmer"], w)
        seq = np.ascontiguousarray(letters).view(f"S{w}").ravel().astype(f"U{w}").astype(object) if n else np.array([], dtype=object)
        cols = {
            "motif_id": [self.motif.motif_id] * n, "motif_alt_id": [self.motif.motif_name] * n,
            "sequence_name": np.array(self.seqnames, dtype=object)[h["name"]] if n else np.array([], dtype=object),
            "start": h["start"].astype(np.int64), "stop": h["stop"].astype(np.int64),
            "strand": np.where(h["strand"] == 45, "-", "+").astype(object),
            "score": np.asarray(self.score_by_bin, dtype=np.float64)[h["bin"]],
            "p-value": np.asarray(self.p_by_bin, dtype=np.float64)[h["bin"]],
        }
        if self.want_q:
            cols["q-value"] = np.asarray(self.q_by_bin, dtype=np.float64)[h["bin"]]
        cols["matched_sequence"] = seq
        cols["haplotype_frequency"] = h["freq"].astype(np.int64)
        cols["reference"] = np.where(h["ref"] == 1, "ref", "non.ref").astype(object)
        return pd.DataFrame(cols)

    def tsv_header(self):
        cols = ["motif_id", "motif_alt_id", "sequence_name", "start", "stop", "strand", "score", "p-value"]
        if self.want_q:
            cols.append("q-value")
        cols += ["matched_sequence", "haplotype_frequency", "reference"]
        return ("\t" + "\t".join(cols) + "\n").encode("utf-8")


def write_results_device(report: DeviceReport, motif: Motif, motif_num: int, args_obj, debug: bool) -> None:
    """write_results for a DeviceReport: `<prefix>.tsv` and `<prefix>.gff` with the reference's byte layout
    (src/grafimo/res_writer.py:103-148,213-303), formatted by K8.  Same directory / prefix rules as write_results.  The HTML
    table (pandas `to_html`, res_writer.py:142) is only written through the DataFrame path: a table of millions of rows is
    not something a browser opens."""
    if report.n == 0:
        exception_handler(ValueError, "No potential motif occurrence retreived.\n", debug)
    if not isinstance(motif_num, int) or motif_num <= 0:
        exception_handler(ValueError, "No motif searched. Probably something went wrong.\n", debug)
    outdir, verbose = args_obj.outdir, args_obj.verbose
    default_name = outdir == DEFAULT_OUTDIR
    if default_name:
        outdir = "_".join(["grafimo_out", str(os.getpid()), motif.motif_id])
    os.makedirs(outdir, exist_ok=True)
    print(f"\nWriting results in {outdir}.\n")
    prefix = "_".join(["grafimo_out", motif.motif_id]) if (not default_name and motif_num > 1) else "grafimo_out"
    base = os.path.join(outdir, prefix)
    try:
        t0 = time.time()
        with open(base + ".tsv", "wb") as fh:
            fh.write(report.tsv_header())
            fh.write(report.render(0))
        if verbose:
            print("%s.tsv written in %.2fs" % (prefix, time.time() - t0))
        t0 = time.time()
        with open(base + ".gff", "wb") as fh:
            fh.write(b"##gff-version 3\n")
            fh.write(report.render(1))
        if verbose:
            print("%s.gff written in %.2fs" % (prefix, time.time() - t0))
        if not getattr(args_obj, "text_only", False):
            if report.n <= HTML_ROW_LIMIT:
                report.to_df().to_html(base + ".html")
            else:
                print(f"{prefix}.html not written: {report.n} rows (more than {HTML_ROW_LIMIT}).")
    except OSError:
        exception_handler(FileWriteError, f"An error ocurred while writing {base}.\n", debug)
