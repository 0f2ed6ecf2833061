"""Motif construction: four file formats -> probability matrix -> log-odds -> integer-scaled score matrix
-> score-distribution DP (GPU).

Public names follow the reference's `grafimo.motif_ops` (src/grafimo/motif_ops.py): `build_motif_jaspar`
(:51), `build_motif_meme` (:237), `build_motif_transfac` (:640), `build_motif_pfm` (:809),
`process_motif_for_logodds` (:971), `scale_pwm` (:1027), `get_motif_pwm` (:1116), `pseudo_bg` (:1189),
`norm_motif` (:1307).  The numeric steps use the same per-element expressions and summation orders as the
reference (they define the integer matrix the GPU consumes and are pinned by the reference's four
`test_motif_processing_*` goldens); the DP for all motifs of a file is one batched GPU launch instead of an
mp.Pool of Cython loops.
"""
import os
import time
from typing import Dict, List, Tuple

import numpy as np

from .grafimo_errors import BGFileError, MotifFileFormatError, MotifFileReadError
from .motif import Motif
from .motif_processing import (apply_pseudocount_jaspar_transfac_pfm, apply_pseudocount_meme, comp_pval_mat_batched,
                               compute_log_odds, get_uniform_bg, read_bg_file)
from .utils import (DNA_ALPHABET, PSEUDOBG, RANGE, REV_COMPL, UNIF, almost_equal, exception_handler, is_jaspar, is_meme,
                    is_pfm, is_transfac, isListEqual)


# ---------------------------------------------------------------------------------------------------------
# background (src/grafimo/motif_ops.py:1189-1302)
# ---------------------------------------------------------------------------------------------------------
def average_bg_with_rc(bgs: Dict[str, float], debug: bool) -> Dict[str, float]:
    """Average every base with its complement; the result is keyed A,T,C,G (insertion order matters for the
    summation order of norm_bg)."""
    out = {}
    for nuc in bgs.keys():
        rc = REV_COMPL[nuc.upper()]
        if REV_COMPL[rc] == nuc and ord(nuc) < ord(rc):
            avg = np.double((bgs[nuc] + bgs[rc]) / np.double(2))
            out[nuc] = avg
            out[rc] = avg
    return out


def norm_bg(bgs: Dict[str, float], debug: bool) -> Dict[str, float]:
    """(bg + 5e-7) / (sum(bg) + 4 * 5e-7), the sum taken in the dict's key order."""
    tot = np.double(len(bgs) * PSEUDOBG)
    for nuc in bgs.keys():
        tot += np.double(bgs[nuc])
    assert tot > 0
    return {nuc: np.double((bgs[nuc] + PSEUDOBG) / tot) for nuc in bgs.keys()}


def pseudo_bg(bgs: Dict[str, float], no_reverse: bool, debug: bool) -> Dict[str, float]:
    if not isinstance(bgs, dict):
        exception_handler(TypeError, f"Expected dict, got {type(bgs).__name__}.\n", debug)
    if not isinstance(no_reverse, bool):
        exception_handler(TypeError, f"Expected bool, got {type(no_reverse).__name__}.\n", debug)
    return norm_bg(bgs if no_reverse else average_bg_with_rc(bgs, debug), debug)


def _background(bg_file: str, alphabet: List[str], no_reverse: bool, debug: bool) -> Dict[str, float]:
    if bg_file == UNIF:
        bgs = get_uniform_bg(alphabet, debug)
    elif os.path.isfile(bg_file):
        bgs = read_bg_file(bg_file, debug)
    else:
        exception_handler(BGFileError, f"Unable to parse {bg_file}.\n", debug)
    return pseudo_bg(bgs, no_reverse, debug)


# ---------------------------------------------------------------------------------------------------------
# probability-matrix helpers
# ---------------------------------------------------------------------------------------------------------
def norm_motif(probs: np.ndarray, rows: List[str], motif_width: int, alphabet: List[str], debug: bool) -> np.ndarray:
    """Renormalise columns whose sum (accumulated in `alphabet` order) is off 1 by more than 1e-5
    (src/grafimo/motif_ops.py:1307-1362).  `rows` names the matrix rows."""
    if motif_width <= 0:
        exception_handler(ValueError, "Forbidden motif width.\n", debug)
    if any(nuc not in DNA_ALPHABET for nuc in alphabet):
        exception_handler(ValueError, "The motif is not built on DNA alphabet.\n", debug)
    probs = np.array(probs, dtype=np.float64)
    ridx = [rows.index(nuc) for nuc in alphabet]
    for j in range(motif_width):
        tot = np.double(0)
        for r in ridx:
            tot += probs[r, j]
        assert tot != 0
        if not almost_equal(1, tot, 0.00001):
            for r in ridx:
                probs[r, j] = np.double(probs[r, j] / tot)
    return probs


def _counts_to_probs(counts: np.ndarray) -> np.ndarray:
    """counts / column sums, rows added top to bottom (pandas' `df / df.sum(0)` in the reference)."""
    colsum = counts[0].copy()
    for r in range(1, counts.shape[0]):
        colsum = colsum + counts[r]
    return counts / colsum[None, :]


def _check_common(motif_file, bg_file, pseudocount, no_reverse, debug):
    if not isinstance(motif_file, str):
        exception_handler(TypeError, f"Expected str, got {type(motif_file).__name__}.\n", debug)
    if not os.path.isfile(motif_file):
        exception_handler(FileNotFoundError, f"Unable to locate {motif_file}.\n", debug)
    if not isinstance(bg_file, str):
        exception_handler(TypeError, f"Expected str, got {type(bg_file).__name__}.\n", debug)
    if bg_file != UNIF and not os.path.isfile(bg_file):
        exception_handler(FileNotFoundError, f"Unable to locate {bg_file}.\n", debug)
    if pseudocount <= 0:
        exception_handler(ValueError, "Pseudocount value must be positive.\n", debug)
    if not isinstance(no_reverse, bool):
        exception_handler(TypeError, f"Expected bool, got {no_reverse}.\n", debug)


def _motif_from_counts(counts, rows, motif_id, motif_name, bg_file, pseudocount, no_reverse, alphabet, debug):
    counts = np.asarray(counts, dtype=np.float64)
    if any(len(c) != counts.shape[1] for c in counts):
        exception_handler(ValueError, "Motif counts width mismatch.\n", debug)
    width = int(counts.shape[1])
    nucsmap = {nuc: i for i, nuc in enumerate(rows)}
    bgs = _background(bg_file, alphabet, no_reverse, debug)
    probs = norm_motif(_counts_to_probs(counts), rows, width, alphabet, debug)
    probs = apply_pseudocount_jaspar_transfac_pfm(counts, probs, pseudocount, bgs, width, alphabet, nucsmap, debug)
    motif = Motif(probs, width, alphabet, motif_id, motif_name, nucsmap)
    motif.set_bg(bgs)
    return motif


# ---------------------------------------------------------------------------------------------------------
# parsers
# ---------------------------------------------------------------------------------------------------------
def _read_jaspar(motif_file, bg_file, pseudocount, no_reverse, verbose, debug) -> Motif:
    nucs, counts = [], []
    try:
        with open(motif_file) as fh:
            header = fh.readline().strip()[1:]
            if not header:
                exception_handler(IOError, f"{motif_file} seems to empty.\n", debug)
            motif_id, motif_name = header.split("\t")[0:2]
            for line in fh:
                line = line.strip()
                if not line:
                    break
                nucs.append(line[:1].upper())
                counts.append([float(x) for x in line[1:].split()[1:][:-1]])  # drop "[" and "]"
        if not counts:
            exception_handler(IOError, f"{motif_file} seems to be empty.\n", debug)
    except (OSError, ValueError):
        exception_handler(MotifFileReadError, f"An error occurred while reading {motif_file}.\n", debug)
    return _motif_from_counts(counts, nucs, motif_id, motif_name, bg_file, pseudocount, no_reverse, sorted(nucs), debug)


def _read_transfac(motif_file, bg_file, pseudocount, no_reverse, verbose, debug) -> Motif:
    motif_id = motif_name = None
    nucs, rows = None, []
    try:
        with open(motif_file) as fh:
            for line in fh:
                line = line.strip()
                if not line:
                    continue
                parts = line.split(None, 1)
                field = parts[0].strip()
                if field == "AC":
                    motif_id = parts[1].strip()
                elif field == "ID":
                    motif_name = parts[1].strip()
                elif field in ("P0", "PO"):
                    nucs = parts[1].strip().split()[:4]
                    assert nucs == DNA_ALPHABET
                    for cline in fh:
                        cparts = cline.strip().split(None, 1)
                        try:
                            position = int(cparts[0].strip())
                        except (ValueError, IndexError):
                            break
                        if len(cparts) != 2:
                            exception_handler(ValueError, f"Invalid count line seen in {motif_file}", debug)
                        if position != len(rows) + 1:
                            exception_handler(ValueError, "Mismatching motif width and position.", debug)
                        vals = cparts[1].strip().split()[:4]
                        if len(vals) != 4:
                            exception_handler(ValueError, "Perhaps the input motif is not a DNA motif", debug)
                        rows.append([float(v) for v in vals])
    except (OSError, AssertionError):
        exception_handler(OSError, f"An error occurred while parsing {motif_file}.", debug)
    counts = np.asarray(rows, dtype=np.float64).T  # -> [4, w], rows in the P0 line's order
    return _motif_from_counts(counts, nucs, motif_id, motif_name, bg_file, pseudocount, no_reverse, sorted(nucs), debug)


def _read_pfm(motif_file, bg_file, pseudocount, no_reverse, verbose, debug) -> Motif:
    motif_id = motif_name = ""
    counts = []
    try:
        with open(motif_file) as fh:
            for line in fh:
                line = line.strip()
                if not line:
                    exception_handler(ValueError, f"{motif_file} seems empty.", debug)
                if line.startswith(">"):
                    motif_id, motif_name = line[1:].split()
                    continue
                counts.append([float(x) for x in line.split()])
        if len(counts) < 2:
            exception_handler(IOError, f"{motif_file} seems to be empty or that it has missing data.", debug)
    except (OSError, ValueError):
        exception_handler(OSError, f"An error occurred while parsing {motif_file}.", debug)
    assert len(counts) == 4
    if not motif_name and not motif_id:
        motif_id = motif_name = os.path.basename(motif_file)
    return _motif_from_counts(counts, list(DNA_ALPHABET), motif_id, motif_name, bg_file, pseudocount, no_reverse,
                              list(DNA_ALPHABET), debug)


def _read_meme(motif_file, bg_file, pseudocount, no_reverse, verbose, debug) -> List[Motif]:
    raw = []
    try:
        with open(motif_file) as fh:
            lines = fh.read().split("\n")
    except OSError:
        exception_handler(MotifFileReadError, f"An error occurred while reading {motif_file}.\n", debug)
    i = 0
    while i < len(lines) and not lines[i].startswith("ALPHABET"):
        i += 1
    if i == len(lines):
        exception_handler(EOFError, f"Unexpected EOF reached, unable to parse {motif_file}.\n", debug)
    if lines[i].strip().replace("ALPHABET= ", "") != "ACGT":
        exception_handler(ValueError, "The motif is not built on DNA alphabet.\n", debug)
    alphabet = sorted("ACGT")
    nucsmap = {a: k for k, a in enumerate(alphabet)}
    while True:
        while i < len(lines) and not lines[i].startswith("MOTIF"):
            i += 1
        if i >= len(lines):
            break
        ids = lines[i].split()
        motif_id, motif_name = (ids[1], ids[1]) if len(ids) == 2 else ids[1:3]
        while i < len(lines) and not lines[i].startswith("letter-probability matrix:"):
            i += 1
        if i >= len(lines):
            exception_handler(EOFError, f"Unexpected premature EOF in {motif_file}.\n", debug)
        stat = lines[i]
        try:
            width = int(stat.split("w=")[1].split()[0])
            nsites = int(stat.split("nsites=")[1].split()[0])
            float(stat.split("E=")[1].split()[0])
        except (IndexError, ValueError):
            exception_handler(MotifFileReadError, f"An error occurred while reading {motif_file}.\n", debug)
        i += 1
        cols = []
        while i < len(lines):
            freqs = lines[i].split()
            if len(freqs) != 4:
                break
            cols.append([np.double(f) for f in freqs])
            i += 1
        if len(cols) < width:
            exception_handler(EOFError, "Unexpected end of motif found.\n", debug)
        raw.append((motif_id, motif_name, width, nsites, np.asarray(cols, dtype=np.float64).T))
    bgs = _background(bg_file, alphabet, no_reverse, debug)
    motifs = []
    for motif_id, motif_name, width, nsites, probs in raw:
        probs = norm_motif(probs, alphabet, width, alphabet, debug)
        probs = apply_pseudocount_meme(probs, pseudocount, nsites, width, bgs, alphabet, nucsmap, debug)
        motif = Motif(probs, width, alphabet, motif_id, motif_name, nucsmap)
        motif.set_bg(bgs)
        motifs.append(motif)
    return motifs


# ---------------------------------------------------------------------------------------------------------
# log-odds -> integer scaling -> DP
# ---------------------------------------------------------------------------------------------------------
def scale_pwm(motif_matrix: np.ndarray, alphabet: List[str], motif_width: int, nucsmap: dict,
              debug: bool) -> Tuple[np.ndarray, int, int, int, np.double]:
    """Integer scaling into [0, 1000] (src/grafimo/motif_ops.py:1090-1111): lower = floor(min) (max - 1 when
    the matrix is constant), offset = round(floor(lower)), scale = floor(1000 / (max - lower)),
    scaled = round((x - offset) * scale) with numpy's round-half-to-even."""
    if not isinstance(motif_matrix, np.ndarray):
        exception_handler(TypeError, f"Expected ndarray, got {type(motif_matrix).__name__}.\n", debug)
    if motif_matrix.size == 0 or motif_matrix.sum() == 0:
        exception_handler(ValueError, "The motif log-odds natrix is empty.\n", debug)
    if not isListEqual(alphabet, DNA_ALPHABET):
        exception_handler(ValueError, "The motif is not built on DNA alphabet.\n", debug)
    if not isinstance(motif_width, int) or motif_width <= 0:
        exception_handler(ValueError, "Forbidden motif width.\n", debug)
    lower = motif_matrix.min()
    upper = motif_matrix.max()
    if lower == upper:
        lower = np.double(upper - 1)
    lower = np.floor(lower)
    offset = np.round(np.floor(lower))
    scale_factor = np.floor(RANGE / (upper - lower))
    scaled = np.round((motif_matrix - offset) * scale_factor).astype(int)
    return scaled, int(scaled.min()), int(scaled.max()), int(scale_factor), np.double(offset)


def _scale_motif(motif: Motif, debug: bool) -> Motif:
    log_odds = compute_log_odds(motif.count_matrix, motif.width, motif.bg, motif.alphabet, motif.nucsmap, debug)
    motif.set_motif_score_matrix(log_odds)
    scaled, min_val, max_val, scale, offset = scale_pwm(motif.score_matrix, motif.alphabet, motif.width, motif.nucsmap, debug)
    motif.set_motif_score_matrix(scaled)
    motif.set_is_scaled()
    motif.set_scale(scale)
    motif.set_min_val(min_val)
    motif.set_max_val(max_val)
    motif.set_offset(offset)
    return motif


def process_motifs_for_logodds(motifs: List[Motif], debug: bool) -> List[Motif]:
    """log-odds + scaling on the host, then ONE batched GPU launch of the DP for all motifs."""
    for m in motifs:
        if not isinstance(m, Motif):
            exception_handler(TypeError, f"Expected Motif, got {type(m).__name__}.\n", debug)
        _scale_motif(m, debug)
    for m, pv in zip(motifs, comp_pval_mat_batched(motifs, debug)):
        m.set_motif_pval_matrix(pv)
    return motifs


def process_motif_for_logodds(motif: Motif, debug: bool) -> Motif:
    """Single-motif form kept for drop-in compatibility (src/grafimo/motif_ops.py:971-1022)."""
    return process_motifs_for_logodds([motif], debug)[0]


# ---------------------------------------------------------------------------------------------------------
# public builders
# ---------------------------------------------------------------------------------------------------------
def _timed(verbose, label, fn):
    t0 = time.time()
    out = fn()
    if verbose:
        print("%s in %.2fs" % (label, time.time() - t0))
    return out


def build_motif_jaspar(motif_file, bg_file, pseudocount, no_reverse, verbose, debug) -> Motif:
    _check_common(motif_file, bg_file, pseudocount, no_reverse, debug)
    motif = _timed(verbose, "Read motif", lambda: _read_jaspar(motif_file, bg_file, pseudocount, no_reverse, verbose, debug))
    return _timed(verbose, f"Motif {motif.motif_id} processed", lambda: process_motif_for_logodds(motif, debug))


def build_motif_transfac(motif_file, bgfile, pseudocount, no_reverse, verbose, debug) -> Motif:
    _check_common(motif_file, bgfile, pseudocount, no_reverse, debug)
    motif = _timed(verbose, "Motif parsed", lambda: _read_transfac(motif_file, bgfile, pseudocount, no_reverse, verbose, debug))
    return _timed(verbose, f"Motif {motif.motif_id} processed", lambda: process_motif_for_logodds(motif, debug))


def build_motif_pfm(motif_file, bgfile, pseudocount, no_reverse, verbose, debug) -> Motif:
    _check_common(motif_file, bgfile, pseudocount, no_reverse, debug)
    motif = _timed(verbose, "Motif parsed", lambda: _read_pfm(motif_file, bgfile, pseudocount, no_reverse, verbose, debug))
    return _timed(verbose, f"Motif {motif.motif_id} processed", lambda: process_motif_for_logodds(motif, debug))


def build_motif_meme(motif_file, bg_file, pseudocount, no_reverse, cores, verbose, debug) -> List[Motif]:
    """All motifs of a MEME file; `cores` is accepted for signature compatibility (the DP is one GPU launch)."""
    _check_common(motif_file, bg_file, pseudocount, no_reverse, debug)
    if not isinstance(pseudocount, float):
        exception_handler(TypeError, f"Expected float, got {type(pseudocount).__name__}.\n", debug)
    motifs = _timed(verbose, f"Read all motifs in {motif_file}",
                    lambda: _read_meme(motif_file, bg_file, pseudocount, no_reverse, verbose, debug))
    print(f"\nRead {len(motifs)} motifs in {motif_file}")
    print("\nProcessing motifs\n")
    return _timed(verbose, f"Processed motif(s) in {motif_file}", lambda: process_motifs_for_logodds(motifs, debug))


def get_motif_pwm(motif_file: str, workflow, cores: int, debug: bool) -> List[Motif]:
    """Format dispatch (src/grafimo/motif_ops.py:1116-1184); always returns a list."""
    if not isinstance(motif_file, str):
        exception_handler(TypeError, f"Expected str, got {type(motif_file).__name__}.\n", debug)
    if not os.path.isfile(motif_file):
        exception_handler(FileNotFoundError, f"Unable to locate {motif_file}.\n", debug)
    args = (workflow.bgfile, workflow.pseudo, workflow.noreverse)
    if is_jaspar(motif_file, debug):
        motif = build_motif_jaspar(motif_file, *args, workflow.verbose, debug)
    elif is_meme(motif_file, debug):
        motif = build_motif_meme(motif_file, *args, cores, workflow.verbose, debug)
    elif is_transfac(motif_file, debug):
        motif = build_motif_transfac(motif_file, *args, workflow.verbose, debug)
    elif is_pfm(motif_file, debug):
        motif = build_motif_pfm(motif_file, *args, workflow.verbose, debug)
    else:
        exception_handler(MotifFileFormatError, "GRAFIMO accepts motifs in JASPAR, MEME, TRANSFAC, or PFM formats.", debug)
    return motif if isinstance(motif, list) else [motif]
