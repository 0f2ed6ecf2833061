"""Reconstruction of the local variation graph behind the reference's real-`vg` k-mer fixture
(tests/test_data/input/width_19/scoring_test_input.tsv of the reference tree: 704 rows that `vg find -K 19 -E -H gbwt`
printed for 22:19723256-19723526 of the 1000-Genomes graph, 5096 haplotypes; committed as
tests/golden/fixtures.json["scoring_input_tsv"]) FROM ITS OWN ROWS -- test infrastructure.

Nothing else in the reference tree holds that region's FASTA / VCF, but the rows determine it:
  * reference bases      from the '+' rows flagged `ref` that span exactly 19 bp (sequence == reference slice);
  * node layout          from column 7 (vg node ids) and the start coordinates of the rows that begin in each node;
  * SNP alleles          from the rows through nodes that never occur in a `ref` row (alternative-allele nodes);
  * the 2-bp deletion    from the rows that span 21 bp (an edge that skips a reference node);
  * haplotype counts     from column 5: the frequency of a row through exactly one alternative allele is that allele's
                         carrier count; the one pair of sites closer than 19 bp fixes the joint counts (their
                         alternative alleles are never on one haplotype: the recombinant row has frequency 0).
The reconstruction yields (reference string, variants, genotype matrix, coordinate offset, node-id offset, region) in
the form oracle/graph_oracle.py and grafimo_b200.vgraph take; the pinning tests then demand that the oracle and K7
print exactly the fixture's 704 rows -- every field, node paths included.
"""
import collections

import numpy as np


def _coord(x):
    return int(x.split(":")[1][:-1])


def _path(field):
    return [int(n[:-1]) for n in field.strip(",").split(",")]


def parse_rows(text):
    rows = []
    for ln in text.split("\n"):
        if not ln.strip():
            continue
        f = ln.split("\t")
        rows.append(dict(region=f[0], seq=f[1], strand=f[2][-1], start=_coord(f[2]), stop=_coord(f[3]), freq=int(f[4]),
                         ref=f[5], nodes=_path(f[6]), line=ln))
    return rows


def reconstruct(text, n_hap=5096, max_node=32):
    rows = parse_rows(text)
    plus = [r for r in rows if r["strand"] == "+"]
    chrom, span = rows[0]["region"].split(":")
    rs, re = (int(x) for x in span.split("-"))
    w = len(plus[0]["seq"])
    # reference bases
    ref = {}
    for r in plus:
        if r["ref"] == "ref" and r["stop"] - r["start"] == w:
            for i, ch in enumerate(r["seq"]):
                assert ref.setdefault(r["start"] + i, ch) == ch
    assert sorted(ref) == list(range(rs, max(ref) + 1))
    ref_end = max(ref) + 1
    ref_nodes = {n for r in plus if r["ref"] == "ref" for n in r["nodes"]}
    all_nodes = {n for r in plus for n in r["nodes"]}
    alt_nodes = all_nodes - ref_nodes
    # first reference coordinate of every node = the smallest start among the rows that begin in it (every node but the
    # first has a row starting at its first base: the region covers it from its first base on)
    a0 = {}
    for r in plus:
        n = r["nodes"][0]
        a0[n] = min(a0.get(n, r["start"]), r["start"])
    first_node = min(all_nodes)
    # alternative alleles: single-base nodes here (SNPs); the base comes from a row that starts IN the node
    succ = collections.defaultdict(set)
    for r in plus:
        for a, b in zip(r["nodes"], r["nodes"][1:]):
            succ[a].add(b)
    variants = []  # (pos, ref allele, alt allele, alt node or None)
    for n in sorted(alt_nodes):
        starts = [r for r in plus if r["nodes"][0] == n]
        if starts:
            pos, base = a0[n], starts[0]["seq"][0]
        else:  # the node is only ever entered from the left: find it inside a row that begins in the previous node
            r = next(r for r in plus if n in r["nodes"][1:] and r["nodes"][0] in a0 and r["nodes"][0] in ref_nodes
                     and r["nodes"].index(n) == 1)
            prev = r["nodes"][0]
            nxt_ref = min(x for x in succ[prev] if x in ref_nodes and x in a0 or x in ref_nodes)
            # length of `prev` from its start to the site = (start of the reference sibling) - a0[prev]; the sibling may
            # have no row starting in it either, so take the position from the row itself
            sib = [x for x in succ[prev] if x != n]
            pos = None
            for q in plus:
                if q["nodes"][:2] == [prev, sib[0]] and q["ref"] == "ref" and q["stop"] - q["start"] == w:
                    # the reference row with the same node prefix: first base where the two sequences differ
                    if q["start"] == r["start"]:
                        d = [i for i in range(w) if q["seq"][i] != r["seq"][i]]
                        pos, base = r["start"] + d[0], r["seq"][d[0]]
                        break
            assert pos is not None
        assert ref[pos] != base
        variants.append([pos, ref[pos], base, n])
    # deletions: an edge between two reference nodes that skips reference nodes
    ref_sorted = sorted(ref_nodes)
    node_len = {}
    for r in plus:  # length of a node = distance to the next node's a0 along a reference row
        ns = r["nodes"]
        for a, b in zip(ns, ns[1:]):
            if a in ref_nodes and b in ref_nodes and a in a0 and b in a0 and b == min(x for x in succ[a] if x in ref_nodes):
                node_len[a] = a0[b] - a0[a]
    for a in ref_sorted:
        for b in succ[a]:
            if b in ref_nodes and b in a0 and a in node_len and a0[b] > a0[a] + node_len[a]:
                gap = (a0[a] + node_len[a], a0[b])
                if any(v[1] != "" and v[0] == gap[0] and len(v[1]) == 1 for v in variants) and gap[1] - gap[0] == 1:
                    continue  # that is the reference/alternative pair of a SNP, not a deletion
                variants.append([gap[0], "".join(ref[p] for p in range(*gap)), "", None])
    variants.sort(key=lambda v: (v[0], v[3] is None))
    # carrier counts: frequency of a row that passes exactly one alternative allele (a deletion: a 21-bp span)
    def alleles_of(r):
        got = set()
        for vi, v in enumerate(variants):
            if v[3] is not None and v[3] in r["nodes"]:
                got.add(vi)
            if v[3] is None and r["stop"] - r["start"] > w and r["start"] < v[0] and r["stop"] > v[0] + len(v[1]):
                got.add(vi)
        return got

    count = {}
    for r in plus:
        al = alleles_of(r)
        if len(al) == 1:
            vi = next(iter(al))
            assert count.setdefault(vi, r["freq"]) == r["freq"], (variants[vi], count[vi], r["freq"])
    assert len(count) == len(variants)
    # genotypes: carriers of different alleles are disjoint unless a row proves otherwise (none does here)
    gt = np.zeros((len(variants), n_hap), dtype=np.uint8)
    nxt = 0
    for vi in range(len(variants)):
        gt[vi, nxt:nxt + count[vi]] = 1
        nxt += count[vi]
    for r in plus:  # joint frequencies of the rows that pass two alternative alleles must hold too
        al = alleles_of(r)
        if len(al) >= 2:
            joint = int(np.all(gt[sorted(al)], axis=0).sum())
            assert joint == r["freq"], (sorted(al), joint, r["freq"])
    # local coordinates: the first node is cut by the region start; give it the bases it had before so that the 32-base
    # chaining of vg's nodes falls where it fell in the real graph
    visible = a0[min(x for x in succ[first_node] if x in ref_nodes)] - rs  # bases of the first node inside the region
    pad = max_node - visible
    offset = rs - pad
    local_ref = "A" * pad + "".join(ref[p] for p in range(rs, ref_end))
    local_variants = [(v[0] - offset, v[1], v[2]) for v in variants]
    return dict(rows=rows, chrom=chrom, region=(rs, re), w=w, offset=offset, ref=local_ref,
                variants=local_variants, gt=gt, node_offset=first_node - 1, n_hap=n_hap,
                local_region=(rs - offset, min(re, ref_end) - offset))


def shift_lines(lines, case):
    """local vg-TSV lines (coordinates and node ids of the reconstructed graph) -> the fixture's coordinates and ids"""
    out = []
    off, noff = case["offset"], case["node_offset"]
    name = f"{case['chrom']}:{case['region'][0]}-{case['region'][1]}"
    for ln in lines:
        f = ln.split("\t")
        c, s = f[2].split(":")
        c2, e = f[3].split(":")
        f[0] = name
        f[2] = f"{case['chrom']}:{int(s[:-1]) + off}{s[-1]}"
        f[3] = f"{case['chrom']}:{int(e[:-1]) + off}{e[-1]}"
        f[6] = "".join(f"{int(n[:-1]) + noff}{n[-1]}," for n in f[6].strip(",").split(","))
        out.append("\t".join(f))
    return out
