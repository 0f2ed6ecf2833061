"""The Motif object -- the type on the drop-in boundary.

Same constructor, setters and read-only properties as the reference's `grafimo.motif.Motif`
(src/grafimo/motif.py:18-483): the scoring path reads `is_scaled, score_matrix, pval_matrix, min_val,
scale, width, offset` (src/grafimo/score_sequences.py:262-268), the report `motif_id, motif_name`
(src/grafimo/resultsTmp.py:272-273) and the DP `bg, alphabet, nucsmap` (src/grafimo/motif_processing.pyx:583-587).
"""
from typing import Dict, List

import numpy as np
import pandas as pd

from .grafimo_errors import NotValidMotifMatrixError
from .utils import DNA_ALPHABET, isListEqual


def _expect(value, kind, what):
    if not isinstance(value, kind):
        raise TypeError(f"\n\nERROR: Expected {getattr(kind, '__name__', kind)}, got {type(value).__name__} ({what}).\n")


class Motif(object):
    def __init__(self, count_matrix: np.ndarray, width: int, alphabet: List[str], motif_id: str, motif_name: str,
                 nucsmap: dict):
        _expect(count_matrix, np.ndarray, "count_matrix")
        if count_matrix.size == 0 or count_matrix.sum() == 0:
            raise NotValidMotifMatrixError("\n\nERROR: Empty motif count matrix.\n")
        _expect(width, int, "width")
        if width <= 0:
            raise ValueError(f"\n\nERROR: Forbidden motif width ({width}).\n")
        _expect(motif_id, str, "motif_id")
        if not motif_id:
            raise ValueError("\n\nERROR: Not valid motif ID.\n")
        _expect(motif_name, str, "motif_name")
        if not motif_name:
            raise ValueError("\n\nERROR: Not valid motif name.\n")
        _expect(alphabet, list, "alphabet")
        if not isListEqual(alphabet, DNA_ALPHABET):
            raise ValueError("\n\nERROR: The motif is not built on DNA alphabet.\n")
        _expect(nucsmap, dict, "nucsmap")
        self._count_matrix = count_matrix
        self._score_matrix = None
        self._pval_matrix = None
        self._min_val = None
        self._max_val = None
        self._scale = None
        self._offset = None
        self._bg = None
        self._width = width
        self._motif_id = motif_id
        self._motif_name = motif_name
        self._alphabet = alphabet
        self._nucsmap = nucsmap
        self._is_scaled = False

    # ---- setters (type rules of src/grafimo/motif.py:189-313) --------------------------------------
    def set_motif_matrix(self, motif_matrix: pd.DataFrame) -> None:
        _expect(motif_matrix, pd.DataFrame, "motif_matrix")
        if motif_matrix.empty:
            raise ValueError("\n\nERROR: Empty motif matrix.\n")
        self._count_matrix = motif_matrix

    def set_motif_score_matrix(self, score_matrix: np.ndarray) -> None:
        _expect(score_matrix, np.ndarray, "score_matrix")
        if score_matrix.size == 0 or score_matrix.sum() == 0:
            raise ValueError("\n\nERROR: Empty motif score matrix.\n")
        self._score_matrix = score_matrix

    def set_motif_pval_matrix(self, pval_mat: np.ndarray) -> None:
        _expect(pval_mat, np.ndarray, "pval_mat")
        if len(pval_mat) == 0:
            raise ValueError("\n\nERROR: Empty motif p-value matrix.\n")
        if pval_mat.sum() == 0:
            raise ValueError("\n\nERROR: Not valid motif p-value matrix.\n")
        self._pval_matrix = pval_mat

    def set_min_val(self, min_val: int) -> None:
        _expect(min_val, int, "min_val")
        self._min_val = min_val

    def set_max_val(self, max_val: int) -> None:
        _expect(max_val, int, "max_val")
        self._max_val = max_val

    def set_scale(self, scale: int) -> None:
        _expect(scale, int, "scale")
        if scale <= 0:
            raise ValueError("\n\nERROR: Scaling factor must be positive integer number.\n")
        self._scale = scale

    def set_offset(self, offset: np.double) -> None:
        _expect(offset, np.double, "offset")  # must stay numpy.float64
        self._offset = offset

    def set_bg(self, bgs: Dict[str, float]) -> None:
        _expect(bgs, dict, "bgs")
        self._bg = bgs

    def set_width(self, width: int) -> None:
        _expect(width, int, "width")
        if width <= 0:
            raise ValueError("\n\nERROR: Not valid motif width.\n")
        self._width = width

    def set_motif_id(self, motif_id: str) -> None:
        _expect(motif_id, str, "motif_id")
        if not motif_id:
            raise ValueError("\n\nERROR: Not valid motif ID.\n")
        self._motif_id = motif_id

    def set_motif_name(self, motif_name: str) -> None:
        _expect(motif_name, str, "motif_name")
        if not motif_name:
            raise ValueError("\n\nERROR: Not valid motif name.\n")
        self._motif_name = motif_name

    def set_alphabet(self, alphabet: List[str]) -> None:
        _expect(alphabet, list, "alphabet")
        if not isListEqual(alphabet, DNA_ALPHABET):
            raise ValueError("\n\nERROR: The motif is not built on DNA alphabet.\n")
        self._alphabet = alphabet

    def set_is_scaled(self) -> None:
        if self._is_scaled:
            raise AssertionError("\n\nERROR: The motif matrix has already been scaled.\n")
        self._is_scaled = True

    # ---- read-only views -------------------------------------------------------------------------------
    def _need(self, value, name):
        if value is None:
            raise AttributeError(f"\n\nERROR: \"self._{name}\" is empty.\n")
        return value

    @property
    def count_matrix(self):
        return self._need(self._count_matrix, "count_matrix")

    @property
    def score_matrix(self):
        return self._need(self._score_matrix, "score_matrix")

    @property
    def pval_matrix(self):
        return self._need(self._pval_matrix, "pval_matrix")

    @property
    def min_val(self):
        return self._need(self._min_val, "min_val")

    @property
    def max_val(self):
        return self._need(self._max_val, "max_val")

    @property
    def scale(self):
        return self._need(self._scale, "scale")

    @property
    def nucsmap(self):
        return self._need(self._nucsmap, "nucsmap")

    @property
    def offset(self):
        return self._need(self._offset, "offset")

    @property
    def bg(self):
        return self._need(self._bg, "bg")

    @property
    def width(self):
        return self._width

    @property
    def motif_id(self):
        return self._motif_id

    @property
    def motif_name(self):
        return self._motif_name

    @property
    def alphabet(self):
        return self._alphabet

    @property
    def is_scaled(self):
        return self._is_scaled

    # ---- helpers used by the B200 path ---------------------------------------------------------------
    def score_matrix_acgt(self) -> np.ndarray:
        """Integer matrix with rows in A,C,G,T order whatever row order the motif file used (nucsmap)."""
        sm = np.asarray(self.score_matrix)
        return np.ascontiguousarray(np.stack([sm[self._nucsmap[n]] for n in DNA_ALPHABET]), dtype=np.int64)

    def bg_acgt(self) -> np.ndarray:
        return np.array([self.bg[n] for n in DNA_ALPHABET], dtype=np.float64)

    def print(self, matrix: str) -> None:
        chosen = {"raw_counts": self._count_matrix, "score_matrix": self._score_matrix, "pval_matrix": self._pval_matrix}
        if matrix not in chosen:
            raise ValueError("\n\nERROR: unable to print the requested matrix.\n")
        print(chosen[matrix])


# ---- the reference's own Motif objects (src/grafimo/motif.py:18-483) carry the same properties: the seams accept them --
_MOTIF_PROPS = ("score_matrix", "pval_matrix", "min_val", "scale", "width", "offset", "is_scaled", "motif_id", "motif_name",
                "bg", "nucsmap")


def is_motif(obj) -> bool:
    """True for this package's Motif and for any object with the properties the path reads from the reference's
    Motif (score_sequences.py:262-268, resultsTmp.py:272-273, motif_processing.pyx:583-587)."""
    return isinstance(obj, Motif) or all(hasattr(obj, a) for a in _MOTIF_PROPS)


def score_matrix_acgt(motif) -> np.ndarray:
    """Integer matrix with rows in A,C,G,T order whatever row order the motif file used (nucsmap)."""
    sm = np.asarray(motif.score_matrix)
    return np.ascontiguousarray(np.stack([sm[motif.nucsmap[n]] for n in DNA_ALPHABET]), dtype=np.int64)


def bg_acgt(motif) -> np.ndarray:
    return np.array([motif.bg[n] for n in DNA_ALPHABET], dtype=np.float64)
