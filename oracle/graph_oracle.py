"""CPU oracle of the k-mer extraction step -- TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's cpu_baseline).

Nothing under grafimo_b200/ imports this file.

What it restates.  GRAFIMO obtains its k-mers by shelling out to the external `vg` binary
(src/grafimo/extract_regions.py:180,225: `vg find -p REGION -x XG -H GBWT -K w -E`) on a graph it built with
`vg construct -r REF -v VCF -R chr -C -a -p` + `vg index -G gbwt -x xg` (src/grafimo/constructVG.py:332,394-396).
`vg` is a third-party program whose source is NOT under /root/reference and whose binary is absent here, so this
file restates the *published behaviour* of those two commands in the most naive way possible (dict-of-lists graph,
recursion, every haplotype spelled out base by base) and is pinned on what the reference tree holds of them:

  * tests/test_data/input/test.fa + test.vcf.gz  ->  tests/test_data/expected_results/expected_seqs.tsv
    (`vg find -x test.xg -E -p x:0-20 -K 19`, tests/grafimo_run_test.py:49-63): all 32 lines, every field,
    including the node path column -- committed as tests/golden/fixtures.json["expected_seqs_tsv"];
  * tests/test_data/input/width_19/scoring_test_input.tsv -- REAL `vg find -K 19 -E -H gbwt` output for
    22:19723256-19723526 of the 1000-Genomes graph (704 rows, 5096 haplotypes, five SNPs and a 2-bp deletion; the
    reference's own test_scoring input).  tests/fixture_graph.py reconstructs that region's reference bases, alleles,
    carrier counts and node layout from the rows themselves, and this oracle then prints EXACTLY those 704 lines:
    sequence, start, stop, strand, haplotype frequency (5096 ... 1 and the 0 of the recombinant walk), `ref` flag
    (the 36 walks through the deletion edge span 21 bp and are flagged `ref`, as vg does), and vg's node ids
    (tests/test_graph_cpu.py::test_oracle_reproduces_real_vg_kmer_fixture; K7 is held to the same lines in
    tests/test_gpu_graph.py::test_extract_equals_real_vg_kmer_fixture).

Parity status: PINNED on real vg output for SNPs, deletions (coordinates of walks through a deletion edge, their `ref`
flag) and haplotype frequencies including 0.  NOT pinned by anything in the reference tree (no fixture holds one): the
coordinates reported for walks that begin or end INSIDE an inserted / multi-base alternative allele -- a documented
choice (nearest reference position, clamped to the allele's reference span).

Model.  Variants are (pos0, ref_allele, alt_allele), already reduced (shared prefix/suffix removed); the graph is
the reference cut at every allele boundary, one node per reference segment and per non-empty alternative allele
(nodes longer than `max_node` are chained), edges between everything that ends and everything that starts at a
breakpoint; a deletion is an edge.  Node ids follow vg construct's order: at a breakpoint the alternative alleles
come first, then the reference segment (this reproduces the fixture's node paths).
"""
from collections import Counter

COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def revcomp(s):
    return "".join(COMP.get(c, "N") for c in reversed(s))


def reduce_variant(pos0, ref, alt):
    """Removes the shared prefix, then the shared suffix, of a VCF REF/ALT pair."""
    ref, alt = ref.upper(), alt.upper()
    while ref and alt and ref[0] == alt[0]:
        ref, alt, pos0 = ref[1:], alt[1:], pos0 + 1
    while ref and alt and ref[-1] == alt[-1]:
        ref, alt = ref[:-1], alt[:-1]
    return pos0, ref, alt


class Node:
    def __init__(self, nid, seq, a0, clamp, isref, bp):
        self.id, self.seq, self.a0, self.clamp, self.isref, self.bp = nid, seq, a0, clamp, isref, bp

    def start_of(self, j):  # reference coordinate reported for a walk STARTING at base j
        return min(self.a0 + j, self.clamp)

    def stop_of(self, j):   # reference coordinate reported for a walk ENDING at base j (exclusive end)
        return min(self.a0 + j + 1, self.clamp)


class Graph:
    def __init__(self):
        self.nodes = []      # Node, id = index + 1
        self.out = {}        # id -> sorted list of ids
        self.first = {}      # ("ref", bp) / ("alt", variant index) -> first node id of the chain
        self.last = {}       # same keys -> last node id of the chain
        self.variants = []
        self.bps = []
        self.length = 0


def build_graph(ref, variants, max_node=32):
    ref = ref.upper()
    L = len(ref)
    g = Graph()
    g.length = L
    g.variants = variants = [(int(s), r.upper(), a.upper()) for s, r, a in variants]
    for s, r, a in variants:
        assert ref[s:s + len(r)] == r, f"REF allele mismatch at {s}"
        assert r != a
    bps = sorted({0, L} | {s for s, r, a in variants} | {s + len(r) for s, r, a in variants})
    g.bps = bps
    at = {b: [] for b in bps}  # variant indices by start, input order
    for vi, (s, r, a) in enumerate(variants):
        at[s].append(vi)
    g.at = at

    def chain(key, seq, a0, clamp, isref, bp):
        ids = []
        for c0 in range(0, len(seq), max_node):
            nid = len(g.nodes) + 1
            piece = seq[c0:c0 + max_node]
            g.nodes.append(Node(nid, piece, a0 + c0, clamp if clamp is not None else a0 + c0 + len(piece), isref, bp))
            g.out[nid] = []
            if ids:
                g.out[ids[-1]].append(nid)
            ids.append(nid)
        g.first[key], g.last[key] = ids[0], ids[-1]

    for i, b in enumerate(bps[:-1]):
        for vi in at[b]:
            s, r, a = variants[vi]
            if a:
                chain(("alt", vi), a, s, s + len(r), False, b)
        chain(("ref", b), ref[b:bps[i + 1]], b, None, True, b)

    # what ends at a breakpoint (through any chain of pure deletions) and what starts there; an insertion sits
    # between the two sides of its breakpoint
    ins_first = {b: [g.first[("alt", vi)] for vi in at[b] if not variants[vi][1]] for b in bps}
    ins_last = {b: [g.last[("alt", vi)] for vi in at[b] if not variants[vi][1]] for b in bps}
    ends = {b: [] for b in bps}
    for i, b in enumerate(bps[1:]):
        ends[b].append(g.last[("ref", bps[i])])
    for vi, (s, r, a) in enumerate(variants):
        if a and r:
            ends[s + len(r)].append(g.last[("alt", vi)])
    for b in bps:  # ascending: the sources of a deletion's start are complete when it is visited
        for vi in at[b]:
            s, r, a = variants[vi]
            if not a:
                ends[s + len(r)].extend(ends[s] + ins_last[s])
    for b in bps[:-1]:
        starts = [g.first[("ref", b)]] + [g.first[("alt", vi)] for vi in at[b] if variants[vi][1] and variants[vi][2]]
        for u in ends[b]:
            g.out[u].extend(ins_first[b] + starts)
        for u in ins_last[b]:
            g.out[u].extend(starts)
    for u in g.out:
        g.out[u] = sorted(set(g.out[u]))
    return g


def haplotype_path(g, carried):
    """Node ids of one haplotype; `carried` = set of variant indices it holds.  A variant that begins inside an
    allele the haplotype already took is skipped (the haplotype never arrives at its breakpoint)."""
    path, b = [], 0

    def chain(key):
        nid = g.first[key]
        while True:
            path.append(nid)
            if nid == g.last[key]:
                return
            nid = [x for x in g.out[nid]][0]

    nxt = {b0: b1 for b0, b1 in zip(g.bps[:-1], g.bps[1:])}
    while b < g.length:
        here = [vi for vi in g.at[b] if vi in carried]
        ins = [vi for vi in here if not g.variants[vi][1]]
        rep = [vi for vi in here if g.variants[vi][1]]
        if ins:
            chain(("alt", ins[0]))
        if rep:
            s, r, a = g.variants[rep[0]]
            if a:
                chain(("alt", rep[0]))
            b = s + len(r)
        else:
            chain(("ref", b))
            b = nxt[b]
    return path


def haplotype_walk_counts(g, gt, w):
    """Counter over (node-id tuple, offset in the first node) of every w-base window of every haplotype."""
    cnt = Counter()
    if gt is None:
        return cnt
    n_hap = len(gt[0]) if len(gt) else 0
    for h in range(n_hap):
        carried = {vi for vi in range(len(g.variants)) if gt[vi][h]}
        bases = [(nid, j) for nid in haplotype_path(g, carried) for j in range(len(g.nodes[nid - 1].seq))]
        for i in range(len(bases) - w + 1):
            win = bases[i:i + w]
            nodes = []
            for nid, _ in win:
                if not nodes or nodes[-1] != nid:
                    nodes.append(nid)
            cnt[(tuple(nodes), win[0][1])] += 1
    return cnt


def enumerate_walks(g, w):
    """Every walk of exactly w bases: (node-id tuple, offset in first node, sequence, offset of the last base in the
    last node), in (first node, offset, depth-first over ascending targets) order."""
    out = []

    def rec(nodes, off0, seq, nid, off):
        node = g.nodes[nid - 1]
        take = min(len(node.seq) - off, w - len(seq))
        seq2 = seq + node.seq[off:off + take]
        if len(seq2) == w:
            out.append((tuple(nodes), off0, seq2, off + take - 1))
            return
        for t in g.out[nid]:
            rec(nodes + [t], off0, seq2, t, 0)

    for node in g.nodes:
        for j in range(len(node.seq)):
            rec([node.id], j, "", node.id, j)
    return out


def extract_rows(g, gt, region, w, n_hap=None):
    """Rows of `vg find -p chr:start-stop -K w -E [-H gbwt]`, forward strand only:
    list of dicts(seq, start, stop, freq, ref, nodes).  freq is 0 for every row when gt is None (no GBWT)."""
    rs, re = region
    cnt = haplotype_walk_counts(g, gt, w)
    rows = []
    for nodes, off0, seq, off_last in enumerate_walks(g, w):
        start = g.nodes[nodes[0] - 1].start_of(off0)
        stop = g.nodes[nodes[-1] - 1].stop_of(off_last)
        if start < rs or stop > re:
            continue
        rows.append(dict(seq=seq, start=start, stop=stop, freq=cnt.get((nodes, off0), 0),
                         ref=all(g.nodes[n - 1].isref for n in nodes), nodes=nodes))
    return rows


def vg_tsv_lines(rows, chrom, region):
    """The 7-column text `vg find` prints, both orientations ('-' rows: reverse complement, start/stop swapped,
    node path reversed -- SURVEY.md F1)."""
    name = f"{chrom}:{region[0]}-{region[1]}"
    lines = []
    for r in rows:
        flag = "ref" if r["ref"] else "non.ref"
        fw = "".join(f"{n}+," for n in r["nodes"])
        rv = "".join(f"{n}-," for n in reversed(r["nodes"]))
        lines.append(f"{name}\t{r['seq']}\t{chrom}:{r['start']}+\t{chrom}:{r['stop']}+\t{r['freq']}\t{flag}\t{fw}")
        lines.append(f"{name}\t{revcomp(r['seq'])}\t{chrom}:{r['stop']}-\t{chrom}:{r['start']}-\t{r['freq']}\t{flag}\t{rv}")
    return lines


def parse_vcf_text(text, chrom=None):
    """Minimal VCF reader for the oracle: -> (variants [(pos0, ref, alt)], gt [variant][haplotype] of 0/1)."""
    variants, gt = [], []
    for line in text.splitlines():
        if not line or line[0] == "#":
            continue
        f = line.split("\t")
        if chrom is not None and f[0] != chrom:
            continue
        alts = f[4].split(",")
        calls = []
        for s in f[9:]:
            g0 = s.split(":")[0].replace("/", "|").split("|")
            calls.extend(int(x) if x.isdigit() else 0 for x in g0)
        for k, alt in enumerate(alts, start=1):
            if alt in (".", "*") or alt.startswith("<"):
                continue
            s, r, a = reduce_variant(int(f[1]) - 1, f[3], alt)
            if r == a:
                continue
            variants.append((s, r, a))
            gt.append([1 if c == k else 0 for c in calls])
    return variants, gt
