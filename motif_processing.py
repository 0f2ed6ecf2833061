"""Top-level `motif_processing` module, the name the reference imports its Cython extension under
(/root/reference setup.py:53; `from motif_processing import ...` at src/grafimo/motif_ops.py:29-35).

With this repository on sys.path ahead of the reference's compiled extension, the reference's own `motif_ops`
picks up the GPU score-distribution DP (`comp_pval_mat` -> gb2_pval_dp_batched, bit-exact) and the host-side PWM
helpers without a source change.  The objects handed in may be the reference's own `Motif` instances: only the
properties both classes share are read.  No CPU fallback: `comp_pval_mat` raises without the CUDA library / a GPU.
"""
from grafimo_b200.motif_processing import (  # noqa: F401
    apply_pseudocount_jaspar_transfac_pfm,
    apply_pseudocount_meme,
    comp_pval_mat,
    comp_pval_mat_batched,
    compute_log_odds,
    get_uniform_bg,
    read_bg_file,
)
