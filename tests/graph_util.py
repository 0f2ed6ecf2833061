"""Helpers of the k-mer extraction tests: seeded random variant sets and a plain-Python reading of the flat
VariationGraph arrays (used on the CPU to check the host-side builder against oracle/graph_oracle.py)."""
import numpy as np

from oracle import graph_oracle as go


def random_case(seed, length=400, n_var=40, n_hap=12, indel=0.35, multi=0.15, overlap=0.1, n_frac=0.0):
    """-> (ref str, variants [(pos0, ref, alt)], gt uint8 [n_var, n_hap]).  SNPs, insertions, deletions, MNPs/complex
    alleles, some multi-allelic sites and some variants that overlap an earlier deletion."""
    rng = np.random.default_rng(seed)
    ref = "".join(rng.choice(list("ACGT"), size=length))
    if n_frac:
        ref = list(ref)
        for i in np.nonzero(rng.random(length) < n_frac)[0]:
            ref[i] = "N"
        ref = "".join(ref)
    variants, seen = [], set()
    pos = np.sort(rng.choice(np.arange(1, length - 12), size=n_var, replace=False))
    for p in pos:
        p = int(p)
        kinds = 1 + int(rng.random() < multi)
        for _ in range(kinds):
            u = rng.random()
            if u > indel:
                r = ref[p]
                a = str(rng.choice([c for c in "ACGT" if c != r]))
            elif u < indel * 0.4:
                r, a = "", "".join(rng.choice(list("ACGT"), size=int(rng.integers(1, 6))))
            elif u < indel * 0.8:
                r, a = ref[p:p + int(rng.integers(1, 7))], ""
            else:
                r = ref[p:p + int(rng.integers(2, 5))]
                a = "".join(rng.choice(list("ACGT"), size=int(rng.integers(1, 5))))
            s, r, a = go.reduce_variant(p, r, a)
            if r == a or (s, r, a) in seen or "N" in r:
                continue
            seen.add((s, r, a))
            variants.append((s, r, a))
            if len(r) > 2 and rng.random() < overlap * 5:  # a SNP inside the span just deleted/replaced
                q = s + 1
                if ref[q] != "N":
                    alt = str(rng.choice([c for c in "ACGT" if c != ref[q]]))
                    if (q, ref[q], alt) not in seen:
                        seen.add((q, ref[q], alt))
                        variants.append((q, ref[q], alt))
    af = rng.random(len(variants)) ** 2
    gt = (rng.random((len(variants), n_hap)) < af[:, None]).astype(np.uint8)
    return ref, variants, gt


def rows_from_arrays(g, region, w):
    """Walk enumeration + popcount(AND of edge sets) over the arrays of grafimo_b200.vgraph.VariationGraph, in plain
    Python -- the algorithm of csrc/graph.cu without a GPU.  -> sorted list of (start, stop, seq, freq, isref, nodes)."""
    rs, re = region
    full = 0xFFFFFFFF
    out = []
    codes = "ACGTN"

    def freq_of(cons):
        if g.n_hap == 0:
            return 0
        cons = [c for c in cons if c != full]
        if not cons:
            return g.n_hap
        acc = g.cons_bits[cons[0]].copy()
        for c in cons[1:]:
            acc &= g.cons_bits[c]
        return int(sum(bin(int(x)).count("1") for x in acc))

    def rec(nodes, cons, seq, n, o, start):
        b0, b1 = int(g.node_off[n]), int(g.node_off[n + 1])
        take = min(b1 - b0 - o, w - len(seq))
        seq2 = seq + "".join(codes[c] for c in g.seq[b0 + o:b0 + o + take])
        if len(seq2) == w:
            stop = min(int(g.node_a0[n]) + o + take, int(g.node_clamp[n]))
            if stop <= re:
                f = freq_of(cons[1:] if len(nodes) > 1 else cons[:1])
                out.append((start, stop, seq2, f, all(g.node_flags[x] & 1 for x in nodes), tuple(int(x) + 1 for x in nodes)))
            return
        for e in range(int(g.edge_off[n]), int(g.edge_off[n + 1])):
            t = int(g.edge_to[e])
            rec(nodes + [t], cons + [int(g.edge_cons[e])], seq2, t, 0, start)

    lo, hi = g.region_nodes(rs, re)
    for n in range(lo, hi):
        for j in range(int(g.node_off[n + 1]) - int(g.node_off[n])):
            start = min(int(g.node_a0[n]) + j, int(g.node_clamp[n]))
            if rs <= start < re:
                rec([n], [int(g.node_cons[n])], "", n, j, start)
    return sorted(out)


def oracle_rows(ref, variants, gt, region, w, max_node=32):
    g = go.build_graph(ref, variants, max_node=max_node)
    rows = go.extract_rows(g, None if gt is None else [list(map(int, r)) for r in gt], region, w)
    return sorted((r["start"], r["stop"], r["seq"], r["freq"], bool(r["ref"]), tuple(r["nodes"])) for r in rows)
