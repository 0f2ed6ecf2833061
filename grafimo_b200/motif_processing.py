"""Host-side PWM maths and the B2 seam `comp_pval_mat`.

Mirrors the public functions of the reference's Cython module `motif_processing`
(src/grafimo/motif_processing.pyx): background readers (:40-188), pseudocounts (:192-440), log-odds
(:444-548) stay on the host -- they are O(4w) and define the integer matrix, so they are written with the
same per-element expressions -- while the score-distribution DP `comp_pval_mat` (:552-632) runs on the GPU
(K3, grafimo_b200/csrc/pval.cu) and is available batched over many motifs.
"""
from typing import Dict, List

import numpy as np

from .grafimo_errors import BGFileError, MotifProcessingError
from .motif import Motif
from .utils import DNA_ALPHABET, RANGE, exception_handler, isListEqual, lg2

_ctx = None


def _context():
    """Process-wide GPU context for the DP (created on first use; fails loudly without a GPU)."""
    global _ctx
    if _ctx is None:
        from .engine import Context
        _ctx = Context()
    return _ctx


# ---- background ---------------------------------------------------------------------------------------
def read_bg_file(bg_file: str, debug: bool) -> Dict[str, float]:
    """0-order MEME background file -> {nuc: prob} in file order (motif_processing.pyx:40-101)."""
    bg, seen = {}, set()
    try:
        with open(bg_file) as fh:
            for line in fh:
                if not line.strip() or line[0] == "#":
                    continue
                if line[0].upper() not in DNA_ALPHABET:
                    exception_handler(ValueError, f"Found symbol not part of the DNA alphabet: {line[0]}\n", debug)
                nuc, prob_str = line.split()[:2]
                prob = float(prob_str)
                assert prob > 0
                if nuc.upper() in seen:
                    exception_handler(BGFileError, f"Found two times {nuc}.\n", debug)
                bg[nuc.upper()] = prob
                seen.add(nuc.upper())
                if len(seen) == len(DNA_ALPHABET):
                    break
    except (OSError, AssertionError, ValueError):
        exception_handler(BGFileError, f"An error occurred while parsing {bg_file}", debug)
    return bg


def get_uniform_bg(alphabet: List[str], debug: bool) -> Dict[str, float]:
    unifp = 1.0 / float(len(alphabet))
    return {a: unifp for a in alphabet}


# ---- pseudocounts --------------------------------------------------------------------------------------
def _bg_column(bgs, alphabet, nucsmap):
    col = np.empty(len(alphabet), dtype=np.float64)
    for nuc in alphabet:
        assert bgs[nuc] > 0
        col[nucsmap[nuc]] = bgs[nuc]
    return col[:, None]


def apply_pseudocount_jaspar_transfac_pfm(counts_matrix, probs_matrix, pseudocount, bgs, width, alphabet, nucsmap, debug):
    """(p * site_counts + pseudo * bg) / (site_counts + pseudo), site_counts = int(sum of the column counts)
    (motif_processing.pyx:192-263)."""
    counts = np.asarray(counts_matrix, dtype=np.float64)
    probs = np.asarray(probs_matrix, dtype=np.float64)
    if counts.size == 0 or counts.sum() == 0:
        exception_handler(ValueError, "Motif counts matrix is empty.\n", debug)
    if probs.size == 0 or probs.sum() == 0:
        exception_handler(ValueError, "Motif probability matrix is empty.\n", debug)
    if not isListEqual(alphabet, DNA_ALPHABET):
        exception_handler(ValueError, "The motif is not built on DNA alphabet.\n", debug)
    if pseudocount <= 0:
        exception_handler(ValueError, "Pseudocount values must be > 0.\n", debug)
    if width <= 0:
        exception_handler(ValueError, "Forbidden motif width.\n", debug)
    pseudo = float(pseudocount)
    site = np.empty(width, dtype=np.float64)
    for j in range(width):
        s = 0
        for v in counts[:, j]:  # python sum(): 0 + c0 + c1 + c2 + c3, then C-int truncation
            s = s + v
        site[j] = float(int(s))
    total = site + pseudo
    return ((probs * site[None, :]) + (pseudo * _bg_column(bgs, alphabet, nucsmap))) / total[None, :]


def apply_pseudocount_meme(probs_matrix, pseudocount, site_counts, width, bgs, alphabet, nucsmap, debug):
    """(p * nsites + pseudo * bg) / (nsites + pseudo) (motif_processing.pyx:313-386)."""
    probs = np.asarray(probs_matrix, dtype=np.float64)
    if probs.size == 0 or probs.sum() == 0:
        exception_handler(ValueError, "The probability matrix is empty.\n", debug)
    if pseudocount <= 0:
        exception_handler(ValueError, "The pseudocount must be > 0.", debug)
    if site_counts <= 0:
        exception_handler(ValueError, "The site counts must be > 0.\n", debug)
    if width <= 0:
        exception_handler(ValueError, "Forbidden motif width.\n", debug)
    if not isListEqual(alphabet, DNA_ALPHABET):
        exception_handler(ValueError, "The motif is not built on DNA alphabet.\n", debug)
    pseudo = float(pseudocount)
    site = float(int(site_counts))
    total = site + pseudo
    return ((probs * site) + (pseudo * _bg_column(bgs, alphabet, nucsmap))) / total


# ---- log-odds -------------------------------------------------------------------------------------------
def compute_log_odds(probs_matrix, width, bgs, alphabet, nucsmap, debug):
    """lo[n, j] = ln(prob / bg) * 1.44269504, evaluated one scalar at a time (motif_processing.pyx:444-507)."""
    probs = np.asarray(probs_matrix, dtype=np.float64)
    if probs.size == 0 or probs.sum() == 0:
        exception_handler(ValueError, "The motif probability matrix is empty.\n", debug)
    if width <= 0:
        exception_handler(ValueError, "Forbidden motif width.\n", debug)
    if not isListEqual(alphabet, DNA_ALPHABET):
        exception_handler(ValueError, "The motif is not built on DNA alphabet.\n", debug)
    out = np.zeros(probs.shape, dtype=np.float64)
    tot_bg = 0.0
    tot_fg = 0.0
    for nuc in alphabet:
        idx = nucsmap[nuc]
        bg = float(bgs[nuc])
        assert bg > 0
        tot_bg += bg
        for j in range(width):
            prob = float(probs[idx, j])
            assert prob > 0
            tot_fg += prob
            out[idx, j] = lg2(prob / bg)
    assert tot_bg - 1.0 < 0.001
    assert tot_fg - width < 0.001
    return out


# ---- B2: score-distribution DP on the GPU ----------------------------------------------------------------
def comp_pval_mat(motif: Motif, debug: bool) -> np.ndarray:
    """Drop-in for `motif_processing.comp_pval_mat` (motif_processing.pyx:608-632): float64[1000*w+1],
    bit-exact to the reference.  Runs K3 on the GPU; raises when no GPU is present."""
    return comp_pval_mat_batched([motif], debug)[0]


def comp_pval_mat_batched(motifs: List[Motif], debug: bool) -> List[np.ndarray]:
    """K3 over many motifs in one launch (one CTA per motif)."""
    for m in motifs:
        if not m.is_scaled:
            exception_handler(MotifProcessingError, "The motif score matrix has not been scaled yet.\n", debug)
        if m.width * RANGE + 1 <= 0:
            exception_handler(MotifProcessingError, "Forbidden motif width.\n", debug)
    ctx = _context()
    from .motif import bg_acgt, score_matrix_acgt
    return ctx.pval_dp_batched([score_matrix_acgt(m) for m in motifs], [bg_acgt(m) for m in motifs])
