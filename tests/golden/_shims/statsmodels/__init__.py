"""Stand-in package: only `statsmodels.stats.multitest.multipletests(method="fdr_bh")`
is provided (the single statsmodels entry point the reference's scoring path calls).
Test infrastructure for tests/golden/make_golden.py only."""
__version__ = "0.0-shim"
