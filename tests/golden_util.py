"""Loaders for the committed golden bundles (tests/golden/, produced by make_golden.py)."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixtures():
    with open(os.path.join(GOLDEN, "fixtures.json")) as fh:
        return json.load(fh)


def motif_tags():
    return sorted(os.path.basename(p)[len("motif_"):-4] for p in glob.glob(os.path.join(GOLDEN, "cases", "motif_*.npz")))


def scoring_tags():
    return sorted(os.path.basename(p)[len("scoring_"):-4] for p in glob.glob(os.path.join(GOLDEN, "cases", "scoring_*.npz")))


def load_motif(tag):
    z = np.load(os.path.join(GOLDEN, "cases", f"motif_{tag}.npz"))
    d = {k: z[k] for k in z.files}
    return dict(
        tag=tag,
        width=int(d["width"]), motif_id=str(d["motif_id"][0]), motif_name=str(d["motif_name"][0]),
        count_matrix=d["count_matrix"], score_matrix=d["score_matrix"], pval_mat=d["pval_mat"],
        min_val=int(d["min_val"]), max_val=int(d["max_val"]), scale=int(d["scale"]), offset=float(d["offset"]),
        bg_acgt=d["bg_acgt"], bg_key_order=[str(x) for x in d["bg_key_order"]],
        source=str(d["source"][0]), fmt=str(d["fmt"][0]), bgfile=str(d["bgfile"][0]),
        no_reverse=bool(d["no_reverse"]), pseudo=float(d["pseudo"]),
    )


def load_scoring(tag):
    z = np.load(os.path.join(GOLDEN, "cases", f"scoring_{tag}.npz"))
    cols = [str(c) for c in z["columns"]]
    table = {}
    for c in cols:
        a = z["col_" + c.replace("-", "_")]
        table[c] = a.astype(object) if a.dtype.kind == "U" else a
    files = [[str(x) for x in z[f"file_{k}"]] for k in range(int(z["n_files"]))]
    out = dict(tag=tag, columns=cols, table=table, files=files, motif_tag=str(z["motif_tag"][0]),
               options=json.loads(str(z["options_json"][0])), stdout=str(z["stdout"][0]))
    for k in ("gff3_head25", "tsv_head25"):
        if k in z.files:
            out[k] = str(z[k][0])
    return out


def canonical_order(table):
    """Row order used for comparisons: the reference's tie order is undefined."""
    keys = (table["matched_sequence"].astype(str), table["strand"].astype(str), table["haplotype_frequency"],
            table["stop"], table["start"], table["p-value"])
    return np.lexsort(keys)


def assert_tables_equal(got, exp, cols, qtol=1e-12, exact_q=True):
    assert len(got["start"]) == len(exp["start"]), (len(got["start"]), len(exp["start"]))
    og, oe = canonical_order(got), canonical_order(exp)
    for c in cols:
        g, e = np.asarray(got[c])[og], np.asarray(exp[c])[oe]
        if c == "q-value":
            if exact_q:
                assert np.array_equal(g.astype(np.float64), e.astype(np.float64)), c
            else:
                np.testing.assert_allclose(g.astype(np.float64), e.astype(np.float64), rtol=qtol, atol=0)
        elif np.asarray(e).dtype.kind == "f":
            assert np.array_equal(g.astype(np.float64), e.astype(np.float64)), c  # bit-exact
        elif np.asarray(e).dtype.kind in "iu":
            assert np.array_equal(g.astype(np.int64), e.astype(np.int64)), c
        else:
            assert list(map(str, g)) == list(map(str, e)), c
