"""grafimo_b200 -- B200-native implementation of GRAFIMO's motif-scanning hot path.

Host side: Python mirroring the reference's call sites (`compute_results`, `comp_pval_mat`, the `Motif`
object, the TSV/GFF3 writers).  Compute: hand-written sm_100a CUDA kernels behind a C ABI
(include/grafimo_b200.h, grafimo_b200/csrc).  No CPU fallback, no Triton, no multi-backend dispatch.
"""
__version__ = "0.1.0"
