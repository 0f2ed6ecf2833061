"""Variation graph of one chromosome, laid out for the GPU k-mer extractor (csrc/graph.cu).

This is the first "next" row of SURVEY.md 8(f): the reference obtains its k-mers by running the external `vg`
program once per BED region (`vg find -p REGION -x XG -H GBWT -K w -E > width_w/REGION.tsv`,
src/grafimo/extract_regions.py:180,225,326) on a graph built by `vg construct -r REF -v VCF -C -a` and
`vg index -G gbwt -x xg` (src/grafimo/constructVG.py:332,394-396), and then parses the text again
(src/grafimo/score_sequences.py:273-293).  Here the same inputs -- reference sequence, phased variants -- become
flat arrays that live in HBM, and the walks of every region are enumerated there straight into packed k-mers
(DeviceGraph.extract), so no text is written or parsed between the graph and the scoring kernel.

Graph model (what `vg construct` builds, restated; pinned on the reference's own vg fixture, see
oracle/graph_oracle.py):
  * the reference is cut at every allele boundary; one node per reference segment and per non-empty alternative
    allele, nodes longer than `max_node_len` (vg's default 32) are chained; a deletion is an edge;
  * at a breakpoint everything that ends there is joined to everything that starts there; an insertion sits between
    the two sides;
  * node ids: at a breakpoint the alternative alleles first (input order), then the reference segment;
  * haplotypes (two per VCF sample, what `vg index -G` threads into the GBWT) follow their alleles; a variant that
    begins inside an allele the haplotype already took is ignored for that haplotype.
Per walk the extractor reports what `vg find -K -E -H` prints: sequence, start/stop on the reference path,
the number of haplotypes that contain the walk's node sequence (0 for walks no haplotype follows -- the rows
`--recomb` is about), and whether every node lies on the reference path.

Haplotype support is stored as bit sets: one row per node (haplotypes through the node) and one per edge
(haplotypes that take the edge).  In a DAG a haplotype contains the node sequence n1..nk exactly when it takes
every edge (n_i -> n_i+1), so the frequency of a walk is popcount(AND of its edge rows) -- or of the node row for
a walk inside one node.  Rows equal to "every haplotype" are not stored (sentinel NO_CONS).
"""
import gzip

import numpy as np

NO_CONS = 0xFFFFFFFF
_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[_c + 32] = _i  # lower case


def read_fasta(path):
    """-> {name: str}.  Plain or gzipped FASTA."""
    op = gzip.open if str(path).endswith(".gz") else open
    seqs, name, parts = {}, None, []
    with op(path, "rt") as fh:
        for line in fh:
            if line.startswith(">"):
                if name is not None:
                    seqs[name] = "".join(parts)
                name, parts = line[1:].split()[0], []
            else:
                parts.append(line.strip())
    if name is not None:
        seqs[name] = "".join(parts)
    return seqs


def reduce_allele(pos0, ref, alt):
    """VCF REF/ALT -> (pos0, ref, alt) without the shared prefix (first) and suffix."""
    ref, alt = ref.upper(), alt.upper()
    k = 0
    n = min(len(ref), len(alt))
    while k < n and ref[k] == alt[k]:
        k += 1
    ref, alt, pos0 = ref[k:], alt[k:], pos0 + k
    k = 0
    n = min(len(ref), len(alt))
    while k < n and ref[len(ref) - 1 - k] == alt[len(alt) - 1 - k]:
        k += 1
    if k:
        ref, alt = ref[:len(ref) - k], alt[:len(alt) - k]
    return pos0, ref, alt


def read_vcf(path, chrom=None):
    """Phased VCF (plain or .gz) -> (variants [(pos0, ref, alt)], gt uint8 [n_variants, n_haplotypes], samples).
    One entry per ALT allele; symbolic / '*' alleles are skipped; a missing call counts as reference; '/' is read
    like '|'.  Haplotype 2i, 2i+1 = the two alleles of sample i (one column for haploid calls)."""
    op = gzip.open if str(path).endswith(".gz") else open
    variants, rows, samples, ploidy = [], [], [], None
    with op(path, "rt") as fh:
        for line in fh:
            if line.startswith("##") or not line.strip():
                continue
            f = line.rstrip("\n").split("\t")
            if line.startswith("#"):
                samples = f[9:]
                continue
            if chrom is not None and f[0] != chrom:
                continue
            calls = []
            for s in f[9:]:
                g = s.split(":", 1)[0].replace("/", "|").split("|")
                if ploidy is None:
                    ploidy = len(g)
                g = (g + ["0"] * ploidy)[:ploidy]
                calls.extend(int(x) if x.isdigit() else 0 for x in g)
            calls = np.asarray(calls, dtype=np.int32)
            for k, alt in enumerate(f[4].split(","), start=1):
                if alt in (".", "*") or alt.startswith("<") or "[" in alt or "]" in alt:
                    continue
                s, r, a = reduce_allele(int(f[1]) - 1, f[3], alt)
                if r == a:
                    continue
                variants.append((s, r, a))
                rows.append((calls == k).astype(np.uint8))
    n_hap = len(samples) * (ploidy or 2)
    gt = np.stack(rows) if rows else np.zeros((0, n_hap), dtype=np.uint8)
    return variants, gt, samples


def pack_bits(gt, words):
    """uint8/bool [n, n_hap] -> uint32 [n, words] (haplotype h = bit h & 31 of word h >> 5)."""
    gt = np.ascontiguousarray(gt, dtype=np.uint8)
    n, h = gt.shape
    out = np.zeros((n, words * 4), dtype=np.uint8)
    if h:
        b = np.packbits(gt, axis=1, bitorder="little")
        out[:, :b.shape[1]] = b
    return out.view(np.uint32)


class VariationGraph:
    """Flat arrays of one chromosome's graph (host side).  Node index = vg node id - 1."""

    def __init__(self):
        self.chrom = ""
        self.length = 0
        self.n_hap = 0
        self.words = 0          # uint32 words per haplotype bit set (multiple of 4)
        self.max_node_len = 32

    # ---------------------------------------------------------------------------------------------------
    @staticmethod
    def build(chrom, ref, variants, gt=None, max_node_len=32):
        """ref: str/bytes; variants: [(pos0, ref_allele, alt_allele)] (reduced); gt: uint8 [n_variants, n_hap] or
        None (no haplotype index: every frequency is reported as 0, as `vg find` does without -H)."""
        g = VariationGraph()
        g.chrom, g.max_node_len = str(chrom), int(max_node_len)
        refb = ref.encode("ascii") if isinstance(ref, str) else bytes(ref)
        L = len(refb)
        g.length = L
        ref_codes = _CODE[np.frombuffer(refb, dtype=np.uint8)]
        nv = len(variants)
        vs = np.array([v[0] for v in variants], dtype=np.int64).reshape(nv)
        vr = np.array([len(v[1]) for v in variants], dtype=np.int64).reshape(nv)
        va = np.array([len(v[2]) for v in variants], dtype=np.int64).reshape(nv)
        if nv and (vs.min() < 0 or (vs + vr).max() > L):
            raise ValueError("variant outside the reference sequence")
        order = np.argsort(vs, kind="stable")  # by start, input order within a start
        refu = refb.upper()
        for i in order:
            s, r, a = variants[i]
            if refu[s:s + len(r)] != r.upper().encode("ascii"):
                raise ValueError(f"REF allele of variant at {s} does not match the reference sequence")
            if r.upper() == a.upper():
                raise ValueError(f"variant at {s} has identical alleles")
        alt_codes = [_CODE[np.frombuffer(variants[i][2].encode("ascii"), dtype=np.uint8)] for i in range(nv)]
        alt_off = np.concatenate([[0], np.cumsum(va)]).astype(np.int64)
        all_codes = np.concatenate([ref_codes] + alt_codes) if nv else ref_codes

        bps = np.unique(np.concatenate([[0, L], vs, vs + vr])).astype(np.int64)
        nb = len(bps) - 1
        # ---- items in node-id order: per breakpoint the non-empty alternative alleles, then the reference segment
        alt_v = order[va[order] > 0]
        it_bp = np.concatenate([np.searchsorted(bps, vs[alt_v]), np.arange(nb)])
        it_isref = np.concatenate([np.zeros(len(alt_v), np.int64), np.ones(nb, np.int64)])
        it_var = np.concatenate([alt_v, np.full(nb, -1)])
        it_rank = np.concatenate([np.arange(len(alt_v)), np.arange(nb)])
        o = np.lexsort((it_rank, it_isref, it_bp))
        it_bp, it_isref, it_var = it_bp[o], it_isref[o], it_var[o]
        n_items = len(o)
        isref = it_isref == 1
        safe_var = np.where(isref, 0, it_var)
        it_len = np.where(isref, bps[np.minimum(it_bp + 1, nb)] - bps[it_bp], va[safe_var])
        it_a0 = np.where(isref, bps[it_bp], vs[safe_var])
        it_clamp = np.where(isref, -1, vs[safe_var] + vr[safe_var])
        it_src = np.where(isref, bps[it_bp], L + alt_off[safe_var])
        M = g.max_node_len
        it_chunks = (it_len + M - 1) // M
        it_first = np.concatenate([[0], np.cumsum(it_chunks)]).astype(np.int64)
        n_nodes = int(it_first[-1])
        nd_item = np.repeat(np.arange(n_items), it_chunks)
        nd_k = np.arange(n_nodes) - it_first[nd_item]
        nd_len = np.minimum(M, it_len[nd_item] - nd_k * M)
        nd_a0 = it_a0[nd_item] + nd_k * M
        nd_isref = isref[nd_item]
        nd_clamp = np.where(nd_isref, nd_a0 + nd_len, it_clamp[nd_item])
        node_off = np.concatenate([[0], np.cumsum(nd_len)]).astype(np.int64)
        total = int(node_off[-1])
        if total >= 2 ** 32 or n_nodes >= 2 ** 31:
            raise ValueError("graph too large for 32-bit base offsets")
        src0 = it_src[nd_item] + nd_k * M
        gather = np.repeat(src0 - node_off[:-1], nd_len) + np.arange(total)
        g.seq = np.ascontiguousarray(all_codes[gather], dtype=np.uint8)
        g.node_off = node_off.astype(np.uint32)
        g.node_a0 = nd_a0.astype(np.int64)
        g.node_clamp = nd_clamp.astype(np.int64)
        g.node_flags = nd_isref.astype(np.uint8)
        g.node_key = np.where(nd_isref, nd_a0, it_a0[nd_item]).astype(np.int64)  # non-decreasing in the node index
        g.n_nodes = n_nodes
        g.max_ref_allele = int(vr.max()) if nv else 0

        # first / last node of every item; items by (breakpoint, kind)
        first_node = it_first[:-1]
        last_node = it_first[1:] - 1
        ref_item_of_bp = np.nonzero(isref)[0]  # one per breakpoint 0..nb-1, ascending
        item_of_var = np.full(nv, -1, dtype=np.int64)
        item_of_var[it_var[~isref]] = np.nonzero(~isref)[0]

        # ---- haplotype bit sets
        n_hap = 0 if gt is None else int(np.asarray(gt).shape[1])
        g.n_hap = n_hap
        W = max(4, ((n_hap + 31) // 32 + 3) // 4 * 4)
        g.words = W
        have_h = gt is not None and n_hap > 0
        if have_h:
            if np.asarray(gt).shape[0] != nv:
                raise ValueError("gt must have one row per variant")
            G = pack_bits(gt, W)
            full = pack_bits(np.ones((1, n_hap), np.uint8), W)[0]
        zero = np.zeros(W, dtype=np.uint32)
        cons_rows = []

        cons_index = {}  # identical sets share a row (an allele's node and the edges into it, ...)

        def cons_id(bits):
            if not have_h or np.array_equal(bits, full):
                return NO_CONS
            key = bits.tobytes()
            k = cons_index.get(key)
            if k is None:
                k = cons_index[key] = len(cons_rows)
                cons_rows.append(bits)
            return k

        node_cons = np.full(n_nodes, NO_CONS, dtype=np.uint32)
        e_src, e_dst, e_cons = [], [], []
        # chain-internal edges: everything on a node continues to the next node of its item
        inner = np.nonzero(nd_k > 0)[0]
        chain_src, chain_dst = inner - 1, inner

        # variants by breakpoint index
        v_bp = np.searchsorted(bps, vs[order])
        v_lo = np.searchsorted(v_bp, np.arange(nb + 1), side="left")
        arrive = [None] * (nb + 1)       # bit set of haplotypes arriving at breakpoint i
        sources = [[] for _ in range(nb + 1)]  # [(last node, bit set)] that end at breakpoint i
        if have_h:
            arrive[0] = full.copy()
        item_set = [None] * n_items
        for i in range(nb):
            here = order[v_lo[i]:v_lo[i + 1]]
            A = arrive[i] if have_h and arrive[i] is not None else zero
            src = sources[i]
            # insertions: between what ends here and what starts here
            ins = [v for v in here if vr[v] == 0]
            if ins:
                left = A.copy() if have_h else zero
                nsrc = []
                for v in ins:
                    it = item_of_var[v]
                    took = (left & G[v]) if have_h else zero
                    if have_h:
                        left = left & ~took
                    item_set[it] = took
                    for u, S in src:
                        e_src.append(u); e_dst.append(first_node[it]); e_cons.append(cons_id(S & took) if have_h else NO_CONS)
                    nsrc.append((last_node[it], took))
                src = [(u, (S & left) if have_h else zero) for u, S in src] + nsrc
            # replacements and deletions that start here, then the reference segment
            left = A.copy() if have_h else zero
            for v in here:
                if vr[v] == 0:
                    continue
                took = (left & G[v]) if have_h else zero
                if have_h:
                    left = left & ~took
                j = int(np.searchsorted(bps, vs[v] + vr[v]))
                if have_h:
                    arrive[j] = took.copy() if arrive[j] is None else (arrive[j] | took)
                if va[v] > 0:
                    it = item_of_var[v]
                    item_set[it] = took
                    for u, S in src:
                        e_src.append(u); e_dst.append(first_node[it]); e_cons.append(cons_id(S & took) if have_h else NO_CONS)
                    sources[j].append((last_node[it], took))
                else:  # deletion: whatever ended here now ends at its far side
                    sources[j].extend((u, (S & took) if have_h else zero) for u, S in src)
            it = ref_item_of_bp[i]
            item_set[it] = left
            for u, S in src:
                e_src.append(u); e_dst.append(first_node[it]); e_cons.append(cons_id(S & left) if have_h else NO_CONS)
            sources[i + 1].append((last_node[it], left))
            if have_h:
                arrive[i + 1] = left.copy() if arrive[i + 1] is None else (arrive[i + 1] | left)
            arrive[i] = None
            sources[i] = None
        if have_h:
            it_cons = np.array([cons_id(s if s is not None else zero) for s in item_set], dtype=np.uint32)
            node_cons = it_cons[nd_item]
        g.node_cons = node_cons.astype(np.uint32)
        # structural edges may repeat (two deletions with the same ends): merge, OR-ing their haplotype sets
        es = np.concatenate([np.asarray(e_src, dtype=np.int64), chain_src])
        ed = np.concatenate([np.asarray(e_dst, dtype=np.int64), chain_dst])
        ec = np.concatenate([np.asarray(e_cons, dtype=np.uint32), node_cons[chain_src].astype(np.uint32)])
        o = np.lexsort((ed, es))
        es, ed, ec = es[o], ed[o], ec[o]
        if len(es) > 1:
            dup = np.nonzero((es[1:] == es[:-1]) & (ed[1:] == ed[:-1]))[0] + 1
            if len(dup):
                for k in dup:  # rare
                    a, b = int(ec[k - 1]), int(ec[k])
                    if a == NO_CONS or b == NO_CONS:
                        ec[k] = NO_CONS
                    else:
                        ec[k] = cons_id(cons_rows[a] | cons_rows[b])
                keep = np.ones(len(es), dtype=bool)
                keep[dup - 1] = False
                es, ed, ec = es[keep], ed[keep], ec[keep]
        g.edge_off = np.searchsorted(es, np.arange(n_nodes + 1), side="left").astype(np.uint32)
        g.edge_to = ed.astype(np.uint32)
        g.edge_cons = ec.astype(np.uint32)
        g.n_edges = len(ed)
        g.cons_bits = (np.stack(cons_rows) if cons_rows else np.zeros((1, W), dtype=np.uint32)).astype(np.uint32)
        g.n_cons = len(cons_rows)
        return g

    @staticmethod
    def from_files(fasta, vcf, chrom, max_node_len=32, use_haplotypes=True):
        seqs = read_fasta(fasta)
        if chrom not in seqs:
            raise KeyError(f"{chrom} is not a sequence of {fasta}")
        variants, gt, _ = read_vcf(vcf, chrom)
        return VariationGraph.build(chrom, seqs[chrom], variants, gt if use_haplotypes else None, max_node_len)

    # ---------------------------------------------------------------------------------------------------
    def region_nodes(self, start, stop):
        """Node index range [lo, hi) that can hold the first base of a walk reported inside [start, stop); scalars
        or arrays."""
        span = max(self.max_node_len, self.max_ref_allele)
        start, stop = np.asarray(start, dtype=np.int64), np.asarray(stop, dtype=np.int64)
        lo = np.searchsorted(self.node_key, start - span, side="left").astype(np.int64)
        hi = np.maximum(lo, np.searchsorted(self.node_key, stop, side="left").astype(np.int64))
        if lo.ndim == 0:
            return int(lo), int(hi)
        return np.ascontiguousarray(lo), np.ascontiguousarray(hi)

    def to_device(self, ctx):
        from .extract_regions import DeviceGraph
        return DeviceGraph(ctx, self)


# ---------------------------------------------------------------------------------------------------------
# K9: the VCF read on the device (csrc/vcf.cu) -- same result as read_vcf, for files of 1000-Genomes size
# ---------------------------------------------------------------------------------------------------------
def _bgzf_blocks(mm):
    """Block table of a BGZF file (the blocked gzip `bgzip` writes; tabix-indexed VCFs such as the 1000 Genomes files are
    BGZF): [(offset of the deflate data, its length, uncompressed size)], or None when the file is plain gzip."""
    import struct
    n = len(mm)
    out, o = [], 0
    while o < n:
        if n - o < 18 or mm[o:o + 4] != b"\x1f\x8b\x08\x04":
            return None
        xlen = struct.unpack_from("<H", mm, o + 10)[0]
        p, end, bsize = o + 12, o + 12 + xlen, None
        while p + 4 <= end:
            si1, si2, slen = mm[p], mm[p + 1], struct.unpack_from("<H", mm, p + 2)[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", mm, p + 4)[0] + 1
            p += 4 + slen
        if bsize is None or o + bsize > n:
            return None
        isize = struct.unpack_from("<I", mm, o + bsize - 4)[0]
        out.append((end, bsize - (12 + xlen) - 8, isize))
        o += bsize
    return out


def _vcf_chunks(path, chunk_bytes, threads=None):
    """Yields uint8 tensors (views of ONE reusable pinned buffer) holding whole lines of a plain, gzipped or BGZF VCF; the
    consumer must be done with a chunk before it asks for the next.  BGZF blocks are independent deflate streams: they
    are inflated by a thread pool (zlib releases the GIL) straight to their place in the buffer -- a single gzip stream
    (~0.2 GB/s) would otherwise bound the whole graph path on real, compressed VCFs."""
    import os
    import torch
    pin = torch.cuda.is_available()
    gz = str(path).endswith(".gz")
    size = None if gz else os.stat(path).st_size
    blocks = mm = None
    if gz and os.stat(path).st_size > 0:
        import mmap
        fh_mm = open(path, "rb")
        mm = mmap.mmap(fh_mm.fileno(), 0, access=mmap.ACCESS_READ)
        blocks = _bgzf_blocks(mm)
        if blocks is None:
            mm.close()
            fh_mm.close()
            mm = None
        else:
            size = sum(b[2] for b in blocks)
    cap = int(chunk_bytes if size is None else min(chunk_bytes, size + 2))
    buf = torch.empty(max(cap, 2), dtype=torch.uint8, pin_memory=pin)
    view = buf.numpy()
    mv = memoryview(view)
    fill = 0

    def cut_and_carry(fill, last):
        """-> (bytes to yield, bytes carried to the front of the buffer)"""
        if last:
            if view[fill - 1] != 10:
                view[fill] = 10
                fill += 1
            return fill, 0
        lo = max(0, fill - (64 << 20))
        cut = bytes(mv[lo:fill]).rfind(b"\n")
        if cut < 0:
            raise ValueError(f"{path}: a line longer than the chunk size / 64 MiB")
        cut += lo + 1
        return cut, fill - cut

    if blocks is not None:
        import zlib
        from concurrent.futures import ThreadPoolExecutor
        nthreads = int(threads or min(32, os.cpu_count() or 1))

        def inflate(job):
            (c0, clen, isize), dst = job
            if isize:
                raw = zlib.decompress(mm[c0:c0 + clen], -15)
                if len(raw) != isize:
                    raise ValueError(f"{path}: corrupt BGZF block at byte {c0}")
                view[dst:dst + isize] = np.frombuffer(raw, dtype=np.uint8)

        try:
            with ThreadPoolExecutor(max_workers=nthreads) as ex:
                k = 0
                while k < len(blocks):
                    jobs = []
                    while k < len(blocks) and fill + blocks[k][2] <= view.shape[0] - 1:
                        jobs.append((blocks[k], fill))
                        fill += blocks[k][2]
                        k += 1
                    if not jobs:
                        raise ValueError(f"{path}: a line longer than the chunk size")
                    list(ex.map(inflate, jobs))
                    if fill == 0:
                        continue
                    n_out, rest = cut_and_carry(fill, k == len(blocks))
                    yield buf[:n_out]
                    if rest:
                        view[:rest] = view[n_out:n_out + rest].copy()
                    fill = rest
        finally:
            mm.close()
            fh_mm.close()
        return

    op = gzip.open if gz else open
    with op(path, "rb") as fh:
        while True:
            room = view.shape[0] - fill - 1
            got = fh.readinto(mv[fill:fill + room]) if room > 0 else 0
            if got:
                fill += got
                if fill < view.shape[0] - 1:
                    continue  # keep filling (gzip returns short reads)
            if fill == 0:
                break
            n_out, rest = cut_and_carry(fill, not got)
            yield buf[:n_out]
            if not got:
                break
            view[:rest] = view[n_out:n_out + rest].copy()
            fill = rest


VCF_MAX_ALT = 16  # csrc/vcf.cu: ALT alleles per line whose genotype rows the kernel builds


def _host_genotype_rows(line: bytes, n_alts: int, ploidy: int, n_hap: int, words: int):
    """Haplotype bit rows of one VCF data line, on the host (the general rule of gb2_vcf_parse_genotypes: bit
    sample * ploidy + j of row a-1 is set when the j-th allele of the sample's call is a).  -> (uint32 [n_alts, words],
    calls out of range)."""
    rows = np.zeros((n_alts, words), dtype=np.uint32)
    bad = 0
    f = line.rstrip(b"\r\n").split(b"\t")
    if len(f) < 10 or not f[8].startswith(b"GT"):
        return rows, bad
    for s_i, col in enumerate(f[9:]):
        call = col.split(b":")[0].replace(b"/", b"|").split(b"|")
        for j, a in enumerate(call):
            if not a.isdigit():
                continue
            a = int(a)
            if a == 0:
                continue
            hbit = s_i * ploidy + j
            if a > n_alts or hbit >= n_hap or j >= ploidy:
                bad += 1
                continue
            rows[a - 1, hbit >> 5] |= np.uint32(1 << (hbit & 31))
    return rows, bad


def read_vcf_device(ctx, path, chrom=None, chunk_bytes=1 << 30, by_chrom=False):
    """Phased VCF -> ({pos int64, ref_len int32, alt_off int64[n+1], alt uint8}, (gt_bits uint32 [n, words], n_hap),
    samples): the arrays DeviceGraph.build / gb2_graph_build take, alleles reduced, one entry per ALT allele, in file
    order (stable-sorted by position).  The text is tokenised on the GPU (gb2_tsv_index_lines, gb2_vcf_parse_fields,
    gb2_vcf_parse_genotypes); the host only slices the allele strings.  Same conventions as read_vcf."""
    import ctypes

    import torch

    from ._lib import check

    def ptr(t):
        return ctypes.c_void_p(t.data_ptr())

    samples, ploidy = None, None
    want = None if chrom is None else np.frombuffer(str(chrom).encode("ascii"), dtype=np.uint8)
    out_pos, out_rlen, out_alt, out_bits, out_ref = [], [], [], [], []
    out_chrom, chrom_ids = [], {}  # by_chrom: chromosome id of every kept row (names in order of first appearance)
    n_hap = words = 0
    skipped_many = bad_calls = 0
    ctx.enter()
    for text in _vcf_chunks(path, chunk_bytes):
        host = text.numpy()
        with torch.cuda.stream(ctx.stream):
            d_text = text.to(ctx.device, non_blocking=True)
        line_off, n = ctx.index_lines(d_text, False, 2)
        if n == 0:
            continue
        i32 = lambda: ctx.empty(n, torch.int32)  # noqa: E731
        kind, clen, pos = ctx.empty(n, torch.uint8), i32(), ctx.empty(n, torch.int64)
        roff, rlen, aoff, alen, nalt, soff, llen = i32(), i32(), i32(), i32(), i32(), i32(), i32()
        check(ctx.lib.gb2_vcf_parse_fields(ctx.h, ptr(d_text), d_text.shape[0], ptr(line_off), n, ptr(kind), ptr(clen), ptr(pos),
                                           ptr(roff), ptr(rlen), ptr(aoff), ptr(alen), ptr(nalt), ptr(soff), ptr(llen)),
              "gb2_vcf_parse_fields", ctx.h)
        with torch.cuda.stream(ctx.stream):
            h = {k: v.cpu().numpy() for k, v in dict(kind=kind, clen=clen, pos=pos, roff=roff, rlen=rlen, aoff=aoff, alen=alen,
                                                     nalt=nalt, soff=soff, llen=llen, off=line_off).items()}
        ctx.sync()
        if (h["kind"] == 2).any():
            bad = int(np.nonzero(h["kind"] == 2)[0][0])
            lo = int(h["off"][bad])
            raise ValueError(f"{path}: malformed VCF line: {bytes(host[lo:lo + 80])!r}")
        if samples is None:  # the column header precedes the first data line
            for i in np.nonzero(h["kind"] == 0)[0]:
                lo = int(h["off"][i])
                if bytes(host[lo:lo + 6]) == b"#CHROM":
                    samples = bytes(host[lo:lo + int(h["llen"][i])]).decode("ascii").rstrip("\r").split("\t")[9:]
        data = h["kind"] == 1
        if want is not None and data.any():
            same = data & (h["clen"] == len(want))
            idx = np.nonzero(same)[0]
            if len(idx):
                names = host[h["off"][idx][:, None] + np.arange(len(want))[None, :]]
                same[idx] = (names == want[None, :]).all(axis=1)
            data = same
        sel = np.nonzero(data & (h["nalt"] > 0))[0]
        if len(sel) == 0:
            continue
        if samples is None:
            samples = []
        if ploidy is None:
            with_s = sel[h["soff"][sel] >= 0]
            ploidy = 2
            if len(with_s):
                lo = int(h["off"][with_s[0]] + h["soff"][with_s[0]])
                call = bytes(host[lo:lo + 64]).split(b"\t")[0].split(b"\n")[0].split(b":")[0]
                ploidy = call.count(b"|") + call.count(b"/") + 1
            n_hap = len(samples) * ploidy
            words = max(4, ((n_hap + 31) // 32 + 3) // 4 * 4)
        row_base = np.full(n, -1, dtype=np.int64)
        counts_alt = h["nalt"][sel].astype(np.int64)
        row_base[sel] = np.concatenate([[0], np.cumsum(counts_alt)[:-1]])
        n_rows = int(counts_alt.sum())
        with torch.cuda.stream(ctx.stream):
            d_bits = torch.zeros((n_rows, words), dtype=torch.int32, device=ctx.device)
            d_base = torch.from_numpy(row_base).to(ctx.device)
            d_counts = torch.zeros(2, dtype=torch.int64, device=ctx.device)
        check(ctx.lib.gb2_vcf_parse_genotypes(ctx.h, ptr(d_text), d_text.shape[0], ptr(line_off), n, ptr(soff), ptr(llen), ptr(nalt), ptr(d_base),
                                              int(ploidy), n_hap, words, ptr(d_bits), ptr(d_counts)), "gb2_vcf_parse_genotypes", ctx.h)
        with torch.cuda.stream(ctx.stream):
            bits = d_bits.cpu().numpy().view(np.uint32)
            cnt = d_counts.cpu().numpy()
        ctx.sync()
        bad_calls += int(cnt[1])
        if int(cnt[0]):  # lines with more ALT alleles than the kernel builds rows for (VCF_MAX_ALT): their genotype rows are
            # filled here, on the host, so that no allele ever enters the graph with an empty haplotype set
            bits = bits.copy()
            for k in np.nonzero(h["nalt"][sel] > VCF_MAX_ALT)[0]:
                i = int(sel[k])
                lo = int(h["off"][i])
                line = bytes(host[lo:lo + int(h["llen"][i])])
                rows_k, bad_k = _host_genotype_rows(line, int(h["nalt"][i]), int(ploidy), n_hap, words)
                b0 = int(row_base[i])
                bits[b0:b0 + rows_k.shape[0]] = rows_k
                bad_calls += bad_k
                skipped_many += 1
        # alleles: single-base REF/ALT lines need no trimming and no per-line Python
        off = h["off"][sel].astype(np.int64)
        simple = (h["rlen"][sel] == 1) & (h["alen"][sel] == 1)
        rb = row_base[sel]
        pos0 = np.zeros(n_rows, dtype=np.int64); rl = np.zeros(n_rows, dtype=np.int32)
        alts = [None] * n_rows
        refs = [b""] * n_rows
        keep = np.zeros(n_rows, dtype=bool)
        si = np.nonzero(simple)[0]
        if len(si):
            r = host[off[si] + h["roff"][sel][si]]
            a_ = host[off[si] + h["aoff"][sel][si]]
            ok = ((a_ & 0xDF) >= 65) & ((a_ & 0xDF) <= 90) & ((a_ & 0xDF) != (r & 0xDF))  # a letter other than REF
            rows_i = rb[si]
            pos0[rows_i] = h["pos"][sel][si] - 1
            rl[rows_i] = 1
            keep[rows_i] = ok
            letters = [bytes([c]) for c in range(256)]
            for row, ch, rc in zip(rows_i[ok].tolist(), (a_ & 0xDF)[ok].tolist(), (r & 0xDF)[ok].tolist()):
                alts[row] = letters[ch]
                refs[row] = letters[rc]
        for k in np.nonzero(~simple)[0]:
            lo = int(off[k])
            ref_s = bytes(host[lo + int(h["roff"][sel][k]):lo + int(h["roff"][sel][k]) + int(h["rlen"][sel][k])]).decode("ascii")
            alt_s = bytes(host[lo + int(h["aoff"][sel][k]):lo + int(h["aoff"][sel][k]) + int(h["alen"][sel][k])]).decode("ascii")
            for j, alt in enumerate(alt_s.split(",")):
                if alt in (".", "*") or alt.startswith("<") or "[" in alt or "]" in alt:
                    continue
                s, r, a_ = reduce_allele(int(h["pos"][sel][k]) - 1, ref_s, alt)
                if r == a_:
                    continue
                row = int(rb[k]) + j
                pos0[row], rl[row], alts[row], keep[row] = s, len(r), a_.encode("ascii"), True
                refs[row] = r.upper().encode("ascii")
        kept = np.nonzero(keep)[0]
        if by_chrom:
            cl = h["clen"][sel].astype(np.int64)
            mx = int(cl.max())
            col = np.arange(mx, dtype=np.int64)[None, :]
            nm = host[np.minimum(off[:, None] + col, len(host) - 1)].copy()
            nm[col >= cl[:, None]] = 0
            uniq, inv = np.unique(np.ascontiguousarray(nm).view(f"S{mx}").ravel(), return_inverse=True)
            gid = np.array([chrom_ids.setdefault(u.decode("ascii"), len(chrom_ids)) for u in uniq], dtype=np.int64)
            out_chrom.append(np.repeat(gid[inv], counts_alt)[kept])
        out_pos.append(pos0[kept]); out_rlen.append(rl[kept]); out_alt.extend(alts[i] for i in kept.tolist())
        out_ref.extend(refs[i] for i in kept.tolist())
        out_bits.append(bits[kept])
    ctx.leave()
    if samples is None:
        samples = []
    if bad_calls:
        import warnings
        warnings.warn(f"{path}: {bad_calls} genotype calls name an allele the line does not have (or a haplotype beyond the "
                      "header's samples); they were read as the reference allele")

    def assemble(rows):
        """variants dict + genotype bit rows of the kept rows `rows` (indices into the concatenated outputs; None = all)"""
        if not out_pos or (rows is not None and len(rows) == 0):
            empty = {"pos": np.zeros(0, np.int64), "ref_len": np.zeros(0, np.int32), "alt_off": np.zeros(1, np.int64),
                     "alt": np.zeros(0, np.uint8), "ref": np.zeros(0, np.uint8)}
            return empty, (np.zeros((0, max(words, 4)), np.uint32), n_hap)
        pos = np.concatenate(out_pos); rlen = np.concatenate(out_rlen); bits = np.concatenate(out_bits)
        idx = np.arange(len(pos)) if rows is None else rows
        order = idx[np.argsort(pos[idx], kind="stable")]
        alt_list = [out_alt[i] for i in order.tolist()]
        alt_off = np.concatenate([[0], np.cumsum([len(a) for a in alt_list])]).astype(np.int64)
        alt = np.frombuffer(b"".join(alt_list), dtype=np.uint8) if alt_list else np.zeros(0, np.uint8)
        ref_cat = b"".join(out_ref[i] for i in order.tolist())  # REF alleles (reduced), back to back: checked against the FASTA
        variants = {"pos": np.ascontiguousarray(pos[order]), "ref_len": np.ascontiguousarray(rlen[order]), "alt_off": alt_off,
                    "alt": alt, "ref": np.frombuffer(ref_cat, dtype=np.uint8), "lines_with_many_alts_read_on_host": skipped_many,
                    "calls_out_of_range": bad_calls}
        return variants, (np.ascontiguousarray(bits[order]), n_hap)

    if by_chrom:  # ONE pass over the file for all chromosomes: {name: (variants, (bits, n_hap))}
        ids = np.concatenate(out_chrom) if out_chrom else np.zeros(0, np.int64)
        return {name: assemble(np.nonzero(ids == k)[0]) for name, k in chrom_ids.items()}, samples
    variants, gtb = assemble(None)
    return variants, gtb, samples
