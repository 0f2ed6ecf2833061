"""ctypes front-end of the CPU oracle (oracle/oracle.c) plus a restatement of the reference's
row handling and result-table semantics.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under grafimo_b200/ imports this module.

Parity status: pinned (see oracle.c header and tests/test_oracle_golden.py).
Reference paths are relative to /root/reference.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
RANGE = 1000  # src/grafimo/utils.py:26

_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_pval_dp.argtypes = [_i64p, ctypes.c_int, _f64p, _f64p]
        L.orc_pval_dp.restype = ctypes.c_int
        L.orc_pvalue.argtypes = [_f64p, ctypes.c_int64, ctypes.c_int64]
        L.orc_pvalue.restype = ctypes.c_double
        L.orc_pvalue_table.argtypes = [_f64p, ctypes.c_int64, _f64p]
        L.orc_pvalue_table.restype = ctypes.c_int
        L.orc_bh.argtypes = [_f64p, ctypes.c_int64, _f64p]
        L.orc_bh.restype = ctypes.c_int
        L.orc_score_rows.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, _i64p, _f64p, ctypes.c_int64,
            ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
        ]
        L.orc_score_rows.restype = ctypes.c_int
        _LIB = L
    return _LIB


# ----------------------------------------------------------------------------------------------
def pval_dp(score_matrix, bg_acgt):
    """src/grafimo/motif_processing.pyx:552-603 -> float64[1000*w+1]."""
    sm = np.ascontiguousarray(score_matrix, dtype=np.int64)
    w = sm.shape[1]
    out = np.zeros(RANGE * w + 1, dtype=np.float64)
    rc = lib().orc_pval_dp(sm, w, np.ascontiguousarray(bg_acgt, dtype=np.float64), out)
    if rc != 0:
        raise ValueError(f"orc_pval_dp failed ({rc})")
    return out


def pvalue(pval_mat, score):
    pm = np.ascontiguousarray(pval_mat, dtype=np.float64)
    return lib().orc_pvalue(pm, pm.shape[0], int(score))


def pvalue_table(pval_mat):
    pm = np.ascontiguousarray(pval_mat, dtype=np.float64)
    out = np.empty_like(pm)
    lib().orc_pvalue_table(pm, pm.shape[0], out)
    return out


def bh(pvalues):
    """src/grafimo/score_sequences.py:401-428 (statsmodels fdr_bh)."""
    p = np.ascontiguousarray(pvalues, dtype=np.float64)
    q = np.empty_like(p)
    lib().orc_bh(p, p.shape[0], q)
    return q


def kmers_to_matrix(seqs, width):
    """list of str -> uint8[n, width] ASCII matrix."""
    joined = "".join(seqs).encode("ascii")
    a = np.frombuffer(joined, dtype=np.uint8)
    if a.size != len(seqs) * width:
        raise ValueError("k-mer length mismatch")
    return a.reshape(len(seqs), width).copy()


def score_rows(ascii_rows, score_matrix, pval_mat, min_val, scale, offset, nthreads=1, want_p=True):
    """src/grafimo/score_sequences.py:331-396 over an ASCII k-mer matrix uint8[n,w].
    Returns (int score, log-odds, p-value)."""
    rows = np.ascontiguousarray(ascii_rows, dtype=np.uint8)
    n, w = rows.shape
    sm = np.ascontiguousarray(score_matrix, dtype=np.int64)
    pm = np.ascontiguousarray(pval_mat, dtype=np.float64)
    assert sm.shape == (4, w) and pm.shape[0] == RANGE * w + 1
    isc = np.empty(n, dtype=np.int64)
    lo = np.empty(n, dtype=np.float64)
    pv = np.empty(n, dtype=np.float64) if want_p else None
    rc = lib().orc_score_rows(
        rows.ctypes.data, n, w, rows.strides[0], sm, pm, int(min_val), int(scale), float(offset),
        isc.ctypes.data, lo.ctypes.data, pv.ctypes.data if want_p else None, int(nthreads))
    if rc < 0:
        raise MemoryError("orc_score_rows")
    return isc, lo, pv


# ----------------------------------------------------------------------------------------------
def parse_rows(lines, noreverse=False):
    """Field handling of src/grafimo/score_sequences.py:279-293,305-307 for vg-find TSV rows."""
    seqname, seq, start, stop, strand, freq, ref = [], [], [], [], [], [], []
    for line in lines:
        data = line.strip().split()
        if not data:
            continue
        st = data[2][-1]
        if noreverse and st == "-":
            continue
        seqname.append(data[0])
        seq.append(data[1])
        start.append(int(data[2].split(":")[1][:-1]))
        stop.append(int(data[3].split(":")[1][:-1]))
        strand.append(st)
        freq.append(int(data[4]))
        ref.append(data[5])
    return dict(seqname=seqname, seq=seq, start=np.array(start, dtype=np.int64),
                stop=np.array(stop, dtype=np.int64), strand=strand, freq=np.array(freq, dtype=np.int64), ref=ref)


def compute_results(motif, lines, threshold=1e-4, noqvalue=False, qvalueT=False, noreverse=False, recomb=False,
                    nthreads=1):
    """Restates compute_results -> score_seqs -> compute_qvalues -> ResultTmp.to_df
    (src/grafimo/score_sequences.py:44-211,216-326,401-428; src/grafimo/resultsTmp.py:241-314).

    `motif` is a dict with score_matrix, pval_mat, min_val, scale, offset, width, motif_id, motif_name.
    Returns a dict of column arrays in the reference's column order; rows sorted by
    (p-value, start, stop, strand, sequence) -- the reference's own tie order is undefined."""
    w = int(motif["width"])
    r = parse_rows(lines, noreverse)
    n = len(r["seq"])
    if n == 0:
        raise ValueError("No result retrieved. Unable to proceed.")
    rows = kmers_to_matrix(r["seq"], w)
    isc, lo, pv = score_rows(rows, motif["score_matrix"], motif["pval_mat"], motif["min_val"], motif["scale"],
                             motif["offset"], nthreads)
    ref = np.array(r["ref"], dtype=object)
    dist = np.abs(r["stop"] - r["start"])
    ref[(ref == "ref") & (dist != w)] = "non.ref"  # score_sequences.py:305-307
    q = None if noqvalue else bh(pv)  # on ALL rows, before any filter (score_sequences.py:194-198)
    keep = (q < threshold) if qvalueT else (pv < threshold)  # resultsTmp.py:303-307 (strict)
    if not recomb:
        keep &= r["freq"] > 0  # resultsTmp.py:309-310
    idx = np.nonzero(keep)[0]
    seq = np.array(r["seq"], dtype=object)
    strand = np.array(r["strand"], dtype=object)
    order = np.lexsort((seq[idx].astype(str), strand[idx].astype(str), r["stop"][idx], r["start"][idx], pv[idx]))
    idx = idx[order]
    out = {
        "motif_id": np.array([motif["motif_id"]] * len(idx), dtype=object),
        "motif_alt_id": np.array([motif["motif_name"]] * len(idx), dtype=object),
        "sequence_name": np.array(r["seqname"], dtype=object)[idx],
        "start": r["start"][idx],
        "stop": r["stop"][idx],
        "strand": strand[idx],
        "score": lo[idx],
        "p-value": pv[idx],
    }
    if not noqvalue:
        out["q-value"] = q[idx]
    out["matched_sequence"] = seq[idx]
    out["haplotype_frequency"] = r["freq"][idx]
    out["reference"] = ref[idx]
    out["_int_score"] = isc[idx]
    out["_scanned"] = n
    return out
