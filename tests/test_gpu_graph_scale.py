"""K7 at benchmark-like sizes: properties that do not need the (slow, pure-Python) oracle on the whole graph, plus the
oracle on sampled regions."""
import numpy as np
import pytest

import graph_util as gr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from grafimo_b200.engine import Context
    return Context(0)


def test_snp_graph_equals_sort_based_haplotype_tally(ctx):
    """SNP-only graph, 100 kb x 2,504 haplotypes: the rows K7 reports with frequency > 0 are exactly the distinct
    (position, k-mer) pairs of the 2.5e8 per-haplotype windows with their multiplicities (gb2_tally_haplotypes, an
    independent sort + run-length route), and the ref flags agree."""
    import torch
    from grafimo_b200 import synth
    from grafimo_b200.vgraph import VariationGraph
    L, H, w = 100_000, 2504, 19
    ref, variants, gt = synth.variant_set(L, H, 77, indel_frac=0.0)
    dg = VariationGraph.build("1", ref, variants, gt).to_device(ctx)
    rows = dg.extract([(0, L)], w)
    with torch.cuda.stream(ctx.stream):
        keep = rows.freq[:rows.n] > 0
        k_start, k_packed = rows.start[:rows.n][keep], rows.packed[:rows.n][keep]
        k_freq, k_ref = rows.freq[:rows.n][keep].to(torch.int64), rows.isref[:rows.n][keep]
        ref_codes = torch.from_numpy(np.frombuffer(ref.encode(), dtype=np.uint8).copy()).to(ctx.device)
        lut = torch.zeros(256, dtype=torch.int64, device=ctx.device)
        lut[torch.tensor([65, 67, 71, 84], device=ctx.device)] = torch.arange(4, device=ctx.device)
        ref_codes = lut[ref_codes.long()]
        vpos = torch.tensor([v[0] for v in variants], device=ctx.device)
        valt = lut[torch.tensor([ord(v[2]) for v in variants], device=ctx.device)]
        gtd = torch.from_numpy(gt).to(ctx.device)
        per = L - w + 1
        packed = torch.empty(H * per, dtype=torch.int64, device=ctx.device)
        pos = torch.arange(per, device=ctx.device).repeat(H)
        for lo in range(0, H, 64):
            hi = min(H, lo + 64)
            codes = ref_codes[None, :].repeat(hi - lo, 1)
            codes[:, vpos] = torch.where(gtd[:, lo:hi].T.bool(), valt[None, :], ref_codes[vpos][None, :])
            packed[lo * per:hi * per] = synth.pack_windows(codes, w).reshape(-1)
        refw = synth.pack_windows(ref_codes[None, :], w).reshape(-1)
    ctx.sync()
    u_pos, u_packed, u_freq, u_isref = ctx.tally_haplotypes(pos, packed, refw, 0)
    with torch.cuda.stream(ctx.stream):
        a = torch.stack([k_start, k_packed, k_freq, k_ref.to(torch.int64)], 1)
        b = torch.stack([u_pos, u_packed, u_freq.to(torch.int64), u_isref.to(torch.int64)], 1)
        a = a[torch.argsort(a[:, 1], stable=True)]
        a = a[torch.argsort(a[:, 0], stable=True)]
        b = b[torch.argsort(b[:, 1], stable=True)]
        b = b[torch.argsort(b[:, 0], stable=True)]
        same = a.shape == b.shape and bool(torch.equal(a, b))
        total = int(k_freq.sum().item())
    assert same
    assert total == H * per


def test_indel_graph_sampled_regions_equal_oracle(ctx):
    """300 kb x 500 haplotypes, 30 % indels: on sampled 200-bp regions every row (start, stop, sequence, frequency,
    ref flag) equals the oracle's, which spells out every haplotype of a slice around the region."""
    from grafimo_b200 import synth
    from grafimo_b200.extract_regions import decode_kmers
    from grafimo_b200.vgraph import VariationGraph
    L, H, w = 300_000, 500, 19
    from grafimo_b200.extract_regions import DeviceGraph
    ref, variants, gt = synth.variant_set(L, H, 99, indel_frac=0.3, density=1 / 25)
    dg = DeviceGraph.build(ctx, "1", ref, variants, gt=gt)  # the library's builder
    rng = np.random.default_rng(5)
    regions = [(int(s), int(s) + 200) for s in rng.integers(1000, L - 1000, size=6)]
    rows = dg.extract(regions, w)
    h = rows.host()
    asc = decode_kmers(h["packed"], w)
    vpos = np.array([v[0] for v in variants])
    for r, (rs, re) in enumerate(regions):
        lo, hi = rs - 100, re + 100
        m = np.nonzero((vpos >= lo + 10) & (vpos < hi - 10))[0]
        sub = [(variants[i][0] - lo, variants[i][1], variants[i][2]) for i in m]
        exp = sorted((s + lo, e + lo, q, f, isref) for (s, e, q, f, isref, _) in
                     gr.oracle_rows(ref[lo:hi], sub, gt[m], (rs - lo, re - lo), w))
        sel = np.nonzero(h["region"] == r)[0]
        got = sorted((int(h["start"][i]), int(h["stop"][i]), asc[i].tobytes().decode(), int(h["freq"][i]), bool(h["isref"][i]))
                     for i in sel)
        assert got == exp and len(got) > 150


def test_frequency_sums_per_first_base(ctx):
    """1 Mb x 2,504 haplotypes (C2 parity form): the frequencies of the walks that share a first base add up to the
    number of haplotypes through that base -- every haplotype spells exactly one k-mer from there."""
    import torch
    from grafimo_b200 import synth
    from grafimo_b200.vgraph import VariationGraph
    L, H, w = 1_000_000, 2504, 19
    ref, variants, gt = synth.variant_set(L, H, 20240)
    g = VariationGraph.build("1", ref, variants, gt)
    dg = g.to_device(ctx)
    rows = dg.extract([(0, L)], w, want_walks=True)
    assert 1_400_000 < rows.n < 1_800_000
    # the library's builder gives the very same rows (k-mers, coordinates, frequencies, flags, walks)
    from grafimo_b200.extract_regions import DeviceGraph
    dn = DeviceGraph.build(ctx, "1", ref, variants, gt=gt)
    assert (dn.info.n_nodes, dn.info.n_edges, dn.info.n_sets) == (g.n_nodes, g.n_edges, g.n_cons)
    rn = dn.extract([(0, L)], w, want_walks=True)
    assert rn.n == rows.n
    for col in ("packed", "start", "stop", "freq", "isref", "walk_len", "walk_off"):
        assert bool(torch.equal(getattr(rn, col)[:rn.n], getattr(rows, col)[:rows.n])), col
    wl = rows.walk_len[:rows.n].to(torch.int64)
    valid = torch.arange(32, device=ctx.device)[None, :] < wl[:, None]
    assert bool(torch.equal(rn.walk.view(-1, 32)[:rn.n][valid], rows.walk.view(-1, 32)[:rows.n][valid]))
    with torch.cuda.stream(ctx.stream):
        first = rows.walk.view(-1, 32)[:rows.n, 0].to(torch.int64) * 64 + rows.walk_off[:rows.n].to(torch.int64)
        uniq, inv = torch.unique_consecutive(first, return_inverse=True)
        sums = torch.zeros(uniq.shape[0], dtype=torch.int64, device=ctx.device).index_add_(0, inv, rows.freq[:rows.n].to(torch.int64))
        node, sums = (uniq // 64).cpu().numpy(), sums.cpu().numpy()
    ctx.sync()
    through = np.array([H if c == 0xFFFFFFFF else int(np.unpackbits(g.cons_bits[c].view(np.uint8)).sum()) for c in g.node_cons[node]])
    inner = g.node_a0[node] < L - 64
    assert np.array_equal(sums[inner], through[inner])
    assert len(np.unique(first.cpu().numpy())) == len(uniq)  # rows of one first base are contiguous
