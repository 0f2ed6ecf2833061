"""Benjamini-Hochberg restatement of statsmodels' `multipletests(..., method="fdr_bh")`
(published algorithm: sort ascending, divide by the empirical CDF k/n, reverse running
minimum, clip at 1, undo the sort).  Pinned bit-for-bit against the q-value column of the
reference's own golden table (704 rows produced by the reference authors with the real
statsmodels) in tests/test_oracle_golden.py.  Test infrastructure only."""
import numpy as np


def multipletests(pvals, alpha=0.05, method="fdr_bh", is_sorted=False, returnsorted=False):
    if method not in ("fdr_bh", "indep", "p", "poscorr"):
        raise NotImplementedError(method)
    p = np.asarray(pvals, dtype=np.float64)
    n = p.shape[0]
    order = np.argsort(p, kind="stable")
    ps = p[order]
    ecdf = np.arange(1, n + 1) / float(n)
    raw = ps / ecdf
    corrected = np.minimum.accumulate(raw[::-1])[::-1]
    corrected[corrected > 1] = 1
    out = np.empty_like(corrected)
    out[order] = corrected
    reject = out <= alpha
    return reject, out, None, None
