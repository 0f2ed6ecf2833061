"""Randomised GPU parity: arbitrary integer score matrices of every width 1..64 (including degenerate ones) with
their oracle-computed score distributions; DP, p-table, both-strand scores, histogram, hits and q-values must be
bit-exact against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    from grafimo_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _random_motif(rng, w, kind):
    if kind == "uniformish":
        sm = rng.integers(0, 1001, size=(4, w))
    elif kind == "spiky":  # one strong base per column, like real PWMs
        sm = rng.integers(0, 120, size=(4, w))
        sm[rng.integers(0, 4, size=w), np.arange(w)] = rng.integers(800, 1001, size=w)
    elif kind == "flat":  # constant columns: span collapses
        sm = np.repeat(rng.integers(1, 50, size=(1, w)), 4, axis=0)
        sm[0, 0] += 1
    else:  # tiny values
        sm = rng.integers(0, 3, size=(4, w))
        sm[0, 0] = 5
    bg = rng.dirichlet([5, 5, 5, 5])
    return sm.astype(np.int64), bg


@pytest.mark.parametrize("seed", list(range(6)) + ["wide0", "wide1"])
def test_random_motifs_all_widths(ctx, seed):
    from grafimo_b200.engine import Scan
    from oracle import oracle as orc
    wide = isinstance(seed, str)  # widths 33..64: two packed words per k-mer, K2's wide kernel
    if wide:
        seed = 6 + int(seed[-1])
    rng = np.random.default_rng(1000 + seed)
    widths = list(range(33, 65)) if wide else list(range(1, 33))
    rng.shuffle(widths)
    if wide:
        widths = sorted(set(widths[:8] + ([33, 64] if seed == 6 else [34, 63])))
    motifs = []
    for i, w in enumerate(widths if wide or not seed else widths[:12]):
        kind = ["uniformish", "spiky", "flat", "tiny"][(i + seed) % 4]
        if 28 < w <= 32 and kind == "uniformish":
            kind = "spiky"  # the span would not fit shared memory; above 32 such spans are counted in global memory
        motifs.append((w, kind) + _random_motif(rng, w, kind))
    # K3 batched, all motifs in one launch
    pvs = ctx.pval_dp_batched([m[2] for m in motifs], [m[3] for m in motifs])
    for (w, kind, sm, bg), pv in zip(motifs, pvs):
        exp = orc.pval_dp(sm, bg)
        assert np.array_equal(pv, exp), (w, kind)
        min_val, scale, offset = int(sm.min()), int(rng.integers(1, 200)), float(-rng.integers(0, 20))
        dm = ctx.motif(sm, pv, min_val, scale, offset)
        tab = orc.pvalue_table(pv)
        assert np.array_equal(dm.ptable, tab[dm.lo:dm.hi + 1]), (w, kind)
        n = int(rng.integers(200, 3000))
        seqs = ["".join(rng.choice(list("ACGT"), size=w)) for _ in range(n)]
        for k in rng.integers(0, n, size=3):
            pos = int(rng.integers(0, w))
            seqs[k] = seqs[k][:pos] + "N" + seqs[k][pos + 1:]
        a = orc.kmers_to_matrix(seqs, w)
        packed, nmask, _ = ctx.encode(torch.from_numpy(a).cuda())
        thr = float(rng.choice([1.0, 0.3, 0.05]))
        sc = Scan(ctx, dm, strands=2, threshold=thr, hit_capacity=2 * n + 8)
        sc.score(packed, nmask)
        out = sc.finalize()
        comp = str.maketrans("ACGTN", "TGCAN")
        a_r = orc.kmers_to_matrix([s.translate(comp)[::-1] for s in seqs], w)
        isf, lof, pf = orc.score_rows(a, sm, pv, min_val, scale, offset)
        isr, lor, pr = orc.score_rows(a_r, sm, pv, min_val, scale, offset)
        p_all, lo_all, is_all = np.concatenate([pf, pr]), np.concatenate([lof, lor]), np.concatenate([isf, isr])
        q_all = orc.bh(p_all)
        idx = out["row"].astype(np.int64) + n * out["strand"].astype(np.int64)
        assert sorted(idx.tolist()) == np.nonzero(p_all < thr)[0].tolist(), (w, kind, thr)
        assert np.array_equal(out["int_score"], is_all[idx]) and np.array_equal(out["score"], lo_all[idx])
        assert np.array_equal(out["p-value"], p_all[idx]) and np.array_equal(out["q-value"], q_all[idx])
        # the dense form (K2 dense scores -> two partition passes) must give the same table, row for row: degenerate score
        # distributions (one bin, tied p-values, non-monotone p-tables) and both k-mer forms included
        den = Scan(ctx, dm, strands=2, threshold=thr, dense_rows=n)
        den.score(packed, nmask)
        dout = den.finalize()
        for k in ("row", "strand", "int_score", "score", "p-value", "q-value"):
            assert np.array_equal(dout[k], out[k]), (w, kind, thr, k)
