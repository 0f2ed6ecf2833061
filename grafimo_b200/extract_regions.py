"""K-mer extraction from the variation graph, on the GPU.

The reference's `scan_graph` (src/grafimo/extract_regions.py:55-237) runs, for every BED region and motif width,
the external command `vg find -p REGION -x XG -H GBWT -K w -E > width_w/REGION.tsv` in an `mp.Pool`
(:128,180,225,275,326) and hands the directory of TSVs to `compute_results`.  Here the graph lives on the device
(`DeviceGraph`, built from the same inputs `grafimo buildvg` takes: reference FASTA + phased VCF, see vgraph.py) and
all regions of a chromosome are enumerated in two kernel launches (`gb2_graph_prepare` / `gb2_graph_extract`,
csrc/graph.cu) into `GraphRows` -- packed k-mers and side arrays that `score_sequences.compute_results_rows` scores
without any text in between.  `GraphRows.to_vg_tsv` / `scan_graph` still write vg's 7-column text for callers (and
parity tests) that want the reference's file interface.

No CPU fallback: DeviceGraph needs the CUDA library and a GPU.
"""
import ctypes
import gzip
import os
from typing import Dict, List, Tuple

import numpy as np
import torch

from ._lib import MAX_WIDTH, NARROW_WIDTH, GraphInfo, check
from .utils import exception_handler
from .vgraph import VariationGraph


def get_regions_bed(bedfile: str, debug: bool) -> Tuple[Dict[str, List], int]:
    """BED reader with the reference's behaviour (src/grafimo/extract_regions.py:371-433): only lines that start
    with "chr" are data, the first three fields are kept as strings, regions are grouped by chromosome in file
    order."""
    if not isinstance(bedfile, str):
        exception_handler(TypeError, f"Expected str, got {type(bedfile).__name__}.\n", debug)
    if not os.path.isfile(bedfile):
        exception_handler(FileNotFoundError, f"Unable to locate {bedfile}.\n", debug)
    if os.stat(bedfile).st_size == 0:
        exception_handler(IOError, f"{bedfile} is empty.\n", debug)
    regions, n = {}, 0
    op = gzip.open if bedfile.split(".")[-1] == "gz" else open
    with op(bedfile, "rt") as fh:
        for line in fh:
            if line.startswith("chr"):
                chrom, start, stop = line.strip().split()[:3]
                regions.setdefault(chrom, []).append((start, stop))
                n += 1
    return regions, n


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class DeviceGraph:
    """A variation graph resident in HBM: from host arrays (gb2_graph_create, `DeviceGraph(ctx, VariationGraph)`) or
    built by the library from reference + alleles + genotype bit sets (gb2_graph_build, `DeviceGraph.build`)."""

    def __init__(self, ctx, graph: VariationGraph = None, handle=None, chrom=None):
        self.ctx, self.graph = ctx, graph
        if handle is not None:
            self.h, self.chrom = handle, str(chrom)
        else:
            g = graph
            self.chrom = g.chrom
            h = ctypes.c_void_p()
            arrays = dict(node_off=np.ascontiguousarray(g.node_off, np.uint32), seq=np.ascontiguousarray(g.seq, np.uint8),
                          a0=np.ascontiguousarray(g.node_a0, np.int64), clamp=np.ascontiguousarray(g.node_clamp, np.int64),
                          flags=np.ascontiguousarray(g.node_flags, np.uint8), ncons=np.ascontiguousarray(g.node_cons, np.uint32),
                          eoff=np.ascontiguousarray(g.edge_off, np.uint32), eto=np.ascontiguousarray(g.edge_to, np.uint32),
                          econs=np.ascontiguousarray(g.edge_cons, np.uint32), bits=np.ascontiguousarray(g.cons_bits, np.uint32))
            check(ctx.lib.gb2_graph_create(ctx.h, g.n_nodes, _np_ptr(arrays["node_off"]), _np_ptr(arrays["seq"]),
                                           _np_ptr(arrays["a0"]), _np_ptr(arrays["clamp"]), _np_ptr(arrays["flags"]),
                                           _np_ptr(arrays["ncons"]), g.n_edges, _np_ptr(arrays["eoff"]), _np_ptr(arrays["eto"]),
                                           _np_ptr(arrays["econs"]), g.n_hap, g.words, g.n_cons, _np_ptr(arrays["bits"]),
                                           ctypes.byref(h)), "gb2_graph_create", ctx.h)
            self.h = h
        info = GraphInfo()
        check(ctx.lib.gb2_graph_get_info(self.h, ctypes.byref(info)), "gb2_graph_get_info")
        self.info = info
        self.n_hap = int(info.n_hap)

    @staticmethod
    def _build_inputs(ref, variants, gt=None, gt_bits=None):
        """Host arrays of gb2_graph_build: (refa, pos, rlen, alt_off, alt, n_hap, words, bits)."""
        from .vgraph import pack_bits
        if isinstance(ref, str):
            ref = ref.encode("ascii")
        refa = np.frombuffer(ref, dtype=np.uint8) if isinstance(ref, (bytes, bytearray)) else np.ascontiguousarray(ref, np.uint8)
        if isinstance(variants, dict):
            pos = np.ascontiguousarray(variants["pos"], np.int64)
            rlen = np.ascontiguousarray(variants["ref_len"], np.int32)
            alt_off = np.ascontiguousarray(variants["alt_off"], np.int64)
            alt = np.ascontiguousarray(variants["alt"], np.uint8)
            if len(pos) > 1 and (np.diff(pos) < 0).any():
                raise ValueError("variant arrays must be sorted by position")
            order = None
            if variants.get("ref") is not None and len(pos):
                # the REF alleles the VCF states must be what the FASTA holds there (a wrong assembly or an off-by-one in
                # the chromosome naming would otherwise build -- and scan -- a wrong graph silently; `vg construct` rejects it)
                rl = rlen.astype(np.int64)
                ref_cat = np.ascontiguousarray(variants["ref"], np.uint8)
                if len(ref_cat) != int(rl.sum()):
                    raise ValueError("REF alleles do not match the ref_len column")
                if (pos < 0).any() or (pos + rl > len(refa)).any():
                    bad = int(np.nonzero((pos < 0) | (pos + rl > len(refa)))[0][0])
                    raise ValueError(f"variant at position {int(pos[bad]) + 1} lies outside the reference sequence ({len(refa)} bp)")
                starts = np.concatenate([[0], np.cumsum(rl)[:-1]])
                idx = np.repeat(pos, rl) + (np.arange(len(ref_cat), dtype=np.int64) - np.repeat(starts, rl))
                diff = (refa[idx] & 0xDF) != (ref_cat & 0xDF)
                if diff.any():
                    k = int(np.searchsorted(starts, np.nonzero(diff)[0][0], side="right") - 1)
                    got = bytes(refa[int(pos[k]):int(pos[k]) + int(rl[k])]).decode("ascii", "replace")
                    want = bytes(ref_cat[int(starts[k]):int(starts[k]) + int(rl[k])]).decode("ascii", "replace")
                    raise ValueError(f"REF allele of the variant at position {int(pos[k]) + 1} is {want!r} in the VCF but the "
                                     f"reference sequence holds {got!r} there ({int(diff.sum())} mismatching bases in all): "
                                     "wrong assembly or chromosome?")
        else:
            nv = len(variants)
            pos = np.array([v[0] for v in variants], dtype=np.int64).reshape(nv)
            order = np.argsort(pos, kind="stable")
            up = refa.tobytes().upper()
            for s_, r_, a_ in variants:
                if up[s_:s_ + len(r_)] != r_.upper().encode("ascii"):
                    raise ValueError(f"REF allele of variant at {s_} does not match the reference sequence")
            pos = np.ascontiguousarray(pos[order])
            rlen = np.array([len(variants[i][1]) for i in order], dtype=np.int32).reshape(nv)
            alts = [variants[i][2].encode("ascii") for i in order]
            alt_off = np.concatenate([[0], np.cumsum([len(a) for a in alts])]).astype(np.int64)
            alt = np.frombuffer(b"".join(alts), dtype=np.uint8) if alts else np.zeros(0, np.uint8)
        nv = len(pos)
        n_hap, bits = 0, None
        if gt_bits is not None:
            bits, n_hap = gt_bits
            bits = np.ascontiguousarray(bits, np.uint32)
        elif gt is not None:
            gt = np.asarray(gt)
            n_hap = int(gt.shape[1])
            words = max(4, ((n_hap + 31) // 32 + 3) // 4 * 4)
            bits = pack_bits(gt if order is None else gt[order], words)
        words = bits.shape[1] if bits is not None else 4
        if bits is not None and bits.shape[0] != nv:
            raise ValueError("one genotype row per variant is required")
        alt_p = alt if len(alt) else np.zeros(1, np.uint8)
        return refa, pos, rlen, alt_off, alt_p, int(n_hap), int(words), bits

    @staticmethod
    def build(ctx, chrom, ref, variants, gt=None, gt_bits=None, max_node_len=32):
        """Native builder (gb2_graph_build).  ref: str/bytes/uint8 array; variants: [(pos0, ref_allele, alt_allele)]
        (reduced) or a dict of arrays {pos int64, ref_len int32, alt_off int64[n+1], alt uint8}; genotypes as
        gt uint8 [n_variants, n_hap] or gt_bits uint32 [n_variants, words] + n_hap (tuple) -- None: no haplotype index.
        Variants are put in position order (stable) like vgraph.VariationGraph.build does."""
        refa, pos, rlen, alt_off, alt_p, n_hap, words, bits = DeviceGraph._build_inputs(ref, variants, gt, gt_bits)
        h = ctypes.c_void_p()
        check(ctx.lib.gb2_graph_build(ctx.h, _np_ptr(refa), len(refa), len(pos), _np_ptr(pos), _np_ptr(rlen), _np_ptr(alt_off),
                                      _np_ptr(alt_p), n_hap, words, _np_ptr(bits) if bits is not None else None,
                                      int(max_node_len), ctypes.byref(h)), "gb2_graph_build", ctx.h)
        return DeviceGraph(ctx, None, handle=h, chrom=chrom)

    @staticmethod
    def build_many(ctx, items, max_node_len=32, n_threads=0):
        """Several chromosomes at once (gb2_graph_build_batch): items = [(chrom, ref, variants, gt, gt_bits)] with the
        argument forms of build(); the host passes run on n_threads worker threads of the library (0 = all hardware
        threads), uploads happen as graphs finish.  -> [DeviceGraph] in item order."""
        from ._lib import GraphInput
        n = len(items)
        if n == 0:
            return []
        keep, arr = [], (GraphInput * n)()
        for i, (chrom, ref, variants, gt, gt_bits) in enumerate(items):
            a = DeviceGraph._build_inputs(ref, variants, gt, gt_bits)
            keep.append(a)  # the arrays must outlive the call
            refa, pos, rlen, alt_off, alt_p, n_hap, words, bits = a
            arr[i] = GraphInput(refa.ctypes.data, len(refa), len(pos), pos.ctypes.data, rlen.ctypes.data, alt_off.ctypes.data,
                                alt_p.ctypes.data, n_hap, words, bits.ctypes.data if bits is not None else None,
                                int(max_node_len), 0)
        outs = (ctypes.c_void_p * n)()
        check(ctx.lib.gb2_graph_build_batch(ctx.h, n, ctypes.cast(arr, ctypes.c_void_p), int(n_threads),
                                            ctypes.cast(outs, ctypes.POINTER(ctypes.c_void_p))), "gb2_graph_build_batch", ctx.h)
        return [DeviceGraph(ctx, None, handle=ctypes.c_void_p(outs[i]), chrom=items[i][0]) for i in range(n)]

    @staticmethod
    def from_files(ctx, fasta, vcf, chrom, display_name=None, max_node_len=32, use_haplotypes=True, parsed_vcf=None):
        """Reference FASTA + phased VCF (the inputs of `grafimo buildvg`, src/grafimo/__main__.py:198-217) -> graph on
        the device: the VCF is tokenised on the GPU (K9), the graph is built by the library (gb2_graph_build).
        parsed_vcf: the {chromosome: (variants, genotype bits)} dict of read_vcf_device(..., by_chrom=True) -- callers
        that build several chromosomes read the file ONCE instead of once per chromosome."""
        from .vgraph import read_fasta, read_vcf_device
        seqs = fasta if isinstance(fasta, dict) else read_fasta(fasta)  # a dict {name: sequence} is taken as is
        if chrom not in seqs:
            raise KeyError(f"{chrom} is not a sequence of {fasta}")
        if parsed_vcf is not None:
            variants, gtb = parsed_vcf.get(chrom, ([], None))
        elif vcf:
            variants, gtb, _ = read_vcf_device(ctx, vcf, chrom)
        else:
            variants, gtb = [], None
        if not use_haplotypes or (gtb is not None and gtb[1] == 0):
            gtb = None
        return DeviceGraph.build(ctx, display_name if display_name is not None else chrom, seqs[chrom], variants, gt_bits=gtb,
                                 max_node_len=max_node_len)

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.gb2_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def extract(self, regions, width, want_walks=False):
        """regions: [(start, stop)] 0-based half-open on the reference path (what `-p chr:start-stop` names).
        -> GraphRows with every walk of `width` bases reported inside its region, forward strand."""
        ctx, g = self.ctx, self.graph
        width = int(width)
        if not 1 <= width <= MAX_WIDTH:
            raise ValueError(f"motif width {width} outside [1, {MAX_WIDTH}]")
        rs = np.array([int(r[0]) for r in regions], dtype=np.int64)
        re = np.array([int(r[1]) for r in regions], dtype=np.int64)
        n_rows = ctypes.c_uint64(0)
        ctx.enter()
        check(ctx.lib.gb2_graph_prepare(ctx.h, self.h, len(rs), _np_ptr(rs), _np_ptr(re), width, ctypes.byref(n_rows)),
              "gb2_graph_prepare", ctx.h)
        n = int(n_rows.value)
        rows = GraphRows(ctx, self.chrom, [(int(a), int(b)) for a, b in zip(rs, re)], width, n, want_walks, self.n_hap)
        rows.graph = g
        if n:
            check(ctx.lib.gb2_graph_extract(ctx.h, self.h, n, _ptr(rows.packed), _ptr(rows.nmask), _ptr(rows.start),
                                            _ptr(rows.stop), _ptr(rows.freq), _ptr(rows.isref), _ptr(rows.region),
                                            _ptr(rows.walk), _ptr(rows.walk_len), _ptr(rows.walk_off), _ptr(rows.counts)),
                  "gb2_graph_extract", ctx.h)
        ctx.leave()
        return rows


_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = {65: 84, 67: 71, 71: 67, 84: 65, 78: 78}


def decode_kmers(packed: np.ndarray, width: int) -> np.ndarray:
    """uint64/int64 [n] (or [n, 2] for widths above 32) packed k-mers -> uint8 [n, width] ASCII."""
    p = packed.astype(np.uint64)
    if p.ndim == 2:
        lo = decode_kmers(p[:, 0], NARROW_WIDTH)
        return np.concatenate([lo, decode_kmers(p[:, 1], width - NARROW_WIDTH)], axis=1)
    sh = (2 * np.arange(width, dtype=np.uint64))[None, :]
    return _ASCII[((p[:, None] >> sh) & np.uint64(3)).astype(np.intp)]


def walk_stride(width: int) -> int:
    """Nodes reserved per row of the walk array of gb2_graph_extract."""
    return NARROW_WIDTH if width <= NARROW_WIDTH else MAX_WIDTH


class GraphRows:
    """Device-resident rows of an extraction: what one `vg find -K w -E -H` call per region would have printed
    (forward strand), as arrays.  Row order: (region, first base, depth-first)."""

    def __init__(self, ctx, chrom, regions, width, n, want_walks, n_hap):
        self.ctx, self.chrom, self.regions, self.width, self.n, self.n_hap = ctx, chrom, regions, width, n, n_hap
        m = max(n, 1)
        from .engine import packed_rows
        self.packed = packed_rows(ctx, m, width)
        self.nmask = ctx.zeros((m + 31) // 32, torch.int32)
        self.start = ctx.empty(m, torch.int64)
        self.stop = ctx.empty(m, torch.int64)
        self.freq = ctx.empty(m, torch.int32)
        self.isref = ctx.empty(m, torch.uint8)
        self.region = ctx.empty(m, torch.int32)
        self.counts = ctx.zeros(2, torch.int64)
        self.walk = ctx.empty(m * walk_stride(width), torch.int32) if want_walks else None
        self.walk_len = ctx.empty(m, torch.uint8) if want_walks else None
        self.walk_off = ctx.empty(m, torch.uint8) if want_walks else None

    def region_name(self, r):
        a, b = self.regions[r]
        return f"{self.chrom}:{a}-{b}"

    def n_masked(self):
        self.ctx.sync()
        return int(self.counts[0].item())

    def host(self):
        """All columns as numpy arrays (tests, TSV writer)."""
        n = self.n
        with torch.cuda.stream(self.ctx.stream):
            out = dict(packed=self.packed[:n].cpu().numpy(), start=self.start[:n].cpu().numpy(), stop=self.stop[:n].cpu().numpy(),
                       freq=self.freq[:n].cpu().numpy(), isref=self.isref[:n].cpu().numpy(), region=self.region[:n].cpu().numpy(),
                       nmask=self.nmask.cpu().numpy())
            if self.walk is not None:
                out["walk"] = self.walk.cpu().numpy().reshape(-1, walk_stride(self.width))[:n]
                out["walk_len"] = self.walk_len[:n].cpu().numpy()
                out["walk_off"] = self.walk_off[:n].cpu().numpy()
        self.ctx.sync()
        return out

    def to_vg_tsv(self, both_strands=True):
        """-> {region index: [lines]} in vg's 7-column layout (pinned by the reference's expected_seqs.tsv).  Rows
        holding a non-ACGT base print 'N' there.  The node path column needs want_walks=True (else it is empty)."""
        h = self.host()
        n, w = self.n, self.width
        asc = decode_kmers(h["packed"], w) if n else np.zeros((0, w), np.uint8)
        if n and h["nmask"].any():
            # the packed form has no room for N: the (rare) masked rows are spelled again from their walks
            bad = np.nonzero((h["nmask"].view(np.uint32)[np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1)[0]
            for i in bad:
                asc[i] = self._ascii_of_walk(h, i)
        comp = np.zeros(256, np.uint8)
        for k, v in _COMP.items():
            comp[k] = v
        out = {r: [] for r in range(len(self.regions))}
        for i in range(n):
            r = int(h["region"][i])
            seq = asc[i].tobytes().decode("ascii")
            flag = "ref" if h["isref"][i] else "non.ref"
            if "walk" in h:
                nodes = h["walk"][i, :h["walk_len"][i]].astype(np.int64) + 1
                fw = "".join(f"{x}+," for x in nodes)
                rv = "".join(f"{x}-," for x in nodes[::-1])
            else:
                fw = rv = ""
            c, name = self.chrom, self.region_name(r)
            out[r].append(f"{name}\t{seq}\t{c}:{h['start'][i]}+\t{c}:{h['stop'][i]}+\t{h['freq'][i]}\t{flag}\t{fw}")
            if both_strands:
                rc = comp[asc[i]][::-1].tobytes().decode("ascii")
                out[r].append(f"{name}\t{rc}\t{c}:{h['stop'][i]}-\t{c}:{h['start'][i]}-\t{h['freq'][i]}\t{flag}\t{rv}")
        return out

    def _ascii_of_walk(self, h, i):
        if "walk" not in h or self.graph is None:
            raise ValueError("rows with non-ACGT bases can only be printed when the walks were kept (want_walks=True) "
                             "and the graph came from host arrays (vgraph.VariationGraph)")
        g = self.graph
        seq, need = [], self.width
        off = int(h["walk_off"][i])
        codes = np.frombuffer(b"ACGTN", dtype=np.uint8)
        for d, nd in enumerate(h["walk"][i, :h["walk_len"][i]]):
            b0, b1 = int(g.node_off[nd]), int(g.node_off[nd + 1])
            o = off if d == 0 else 0
            take = min(b1 - b0 - o, need)
            seq.append(codes[g.seq[b0 + o:b0 + o + take]])
            need -= take
        return np.concatenate(seq)


def scan_graph(graphs: Dict[str, DeviceGraph], bedfile: str, widths, outdir: str, debug: bool = False, chroms_prefix: str = "") -> str:
    """File-interface twin of the reference's scan_graph (src/grafimo/extract_regions.py:55-237): writes
    `<outdir>/width_<w>/<chrom>_<start>-<stop>.tsv` for every BED region and width, in vg's layout, and returns
    outdir -- the `sequence_loc` compute_results takes.  `graphs` maps the BED chromosome name without its "chr"
    prefix (what the reference puts in the region name, :150-170) to a DeviceGraph."""
    regions, _ = get_regions_bed(bedfile, debug)
    for w in widths:
        d = os.path.join(outdir, f"width_{int(w)}")
        os.makedirs(d, exist_ok=True)
        for chrom, spans in regions.items():
            key = chrom[3:] if chrom.startswith("chr") else chrom
            name = chroms_prefix + key
            dg = graphs.get(name) or graphs.get(key) or graphs.get(chrom)
            if dg is None:
                exception_handler(KeyError, f"No variation graph for chromosome {chrom}.\n", debug)
            rows = dg.extract([(int(a), int(b)) for a, b in spans], int(w), want_walks=True)
            for r, lines in rows.to_vg_tsv().items():
                a, b = rows.regions[r]
                with open(os.path.join(d, f"{rows.chrom}_{a}-{b}.tsv"), "w") as fh:
                    fh.write("".join(line + "\n" for line in lines))
    return outdir
