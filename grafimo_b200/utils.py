"""Constants and small helpers shared by the host side.

The constants are parity-relevant and mirror src/grafimo/utils.py:19-32 of the reference (RANGE, PSEUDOBG and
the truncated log2 factor all enter the integer score matrix).
"""
import os
import sys

import numpy as np

DNA_ALPHABET = ["A", "C", "G", "T"]
REV_COMPL = {"A": "T", "C": "G", "G": "C", "T": "A"}
NOMAP = "NOMAP"
ALL_CHROMS = "use_all_chroms"
UNIF = "unfrm_dst"
PSEUDOBG = np.double(0.0000005)
LOG_FACTOR = 1.44269504  # sic: the reference multiplies ln(x) by this truncated 1/ln(2)
RANGE = 1000
DEFAULT_OUTDIR = "default_out_dir_name"
SOURCE = "grafimo"
TP = "nucleotide_motif"
PHASE = "."


def die(code):
    sys.exit(code)


def sigint_handler():
    print("\nCaught SIGINT. GRAFIMO will exit")
    die(2)


def exception_handler(exception_type, exception, debug):
    """src/grafimo/utils.py:63-78: raise when debugging, otherwise a one-line error on stderr and exit code 1."""
    if debug:
        raise exception_type(f"\n\n{exception}")
    sys.stderr.write("\n\nERROR: " + f"{exception}")
    die(1)


def isListEqual(lst1, lst2):
    return len(lst1) == len(lst2) and sorted(lst1) == sorted(lst2)


def almost_equal(value1, value2, slope):
    return not ((value1 - slope) > value2 or (value1 + slope) < value2)


def lg2(value):
    """src/grafimo/utils.py:479-493: ln(x) * 1.44269504 (one scalar at a time, like the reference)."""
    return np.log(value) * LOG_FACTOR


def is_numeric(s):
    try:
        float(s)
    except ValueError:
        return False
    return True


def _readable(motif_file, debug):
    if not isinstance(motif_file, str):
        exception_handler(TypeError, f"Expected str, got {type(motif_file).__name__}.\n", debug)
    if not os.path.isfile(motif_file):
        exception_handler(FileNotFoundError, f"Unable to locate {motif_file}.\n", debug)
    if os.stat(motif_file).st_size == 0:
        exception_handler(EOFError, f"{motif_file} seems to be empty.\n", debug)


def is_jaspar(motif_file, debug=False):
    """Format sniffers: same acceptance rules as src/grafimo/utils.py:212-405."""
    _readable(motif_file, debug)
    if motif_file.split(".")[-1] != "jaspar":
        return False
    with open(motif_file) as fh:
        if not fh.readline().strip().startswith(">"):
            return False
        for line in fh:
            tok = line.strip().split()
            if not tok or len(tok) < 3 or tok[1] != "[" or tok[-1] != "]":
                return False
            if not all(is_numeric(c) for c in tok[2:-1]):
                return False
    return True


def is_meme(motif_file, debug=False):
    _readable(motif_file, debug)
    with open(motif_file) as fh:
        return any(line.startswith("MEME version") for line in fh)


def is_transfac(motif_file, debug=False):
    _readable(motif_file, debug)
    seen = {"AC": False, "ID": False, "PO": False}
    width = 0
    with open(motif_file) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            parts = line.split(None, 1)
            field = parts[0].strip()
            if len(field) != 2:
                return False
            if len(parts) != 2:
                continue
            value = parts[1].strip()
            if field in seen:
                if not value:
                    return False
                if field in ("P0", "PO") and value.split()[:4] != DNA_ALPHABET:
                    return False
                seen[field] = True
            try:
                position = int(field)
            except ValueError:
                continue
            if width == 0 and position == 0:
                return False
            width += 1
            if width != position:
                return False
    return sum(seen.values()) == 3


def is_pfm(motif_file, debug=False):
    _readable(motif_file, debug)
    with open(motif_file) as fh:
        for line in fh:
            if line.startswith(">"):
                continue
            if not all(is_numeric(c) for c in line.strip().split()):
                return False
    return True


def dftolist(data, no_qvalue, debug=False):
    """Column lists in the order the GFF3 writer indexes them (src/grafimo/utils.py:498-575)."""
    cols = ["motif_id", "motif_alt_id", "sequence_name", "start", "stop", "strand", "score", "p-value",
            "matched_sequence", "haplotype_frequency", "reference"]
    if len(data) == 0:
        exception_handler(ValueError, "Empty DataFrames cannot be converted to lists of values.\n", debug)
    out = [data[c].tolist() for c in cols]
    if not no_qvalue:
        out.append(data["q-value"].tolist())
    return out
