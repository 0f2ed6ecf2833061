"""Seeded synthetic inputs of the shapes named in BASELINE.json / SURVEY.md 8(d).

`haplotype_windows` builds the throughput form of config C2: a random reference region, 1000-Genomes-like
variant sites (density 1/40 bp, 90 % SNPs / 10 % indels of 1-5 bp, allele frequency ~ 1/x on [1/H, 0.5]), H
haplotypes that carry each variant with probability af, and every motif-width window of every haplotype as a
2-bit packed uint64 (the layout of include/grafimo_b200.h).  Works on CPU tensors (tests, reference arm) and on
CUDA tensors (bench) with the same code.  These are generators of test/bench inputs, not part of the scanning path.
"""
import numpy as np
import torch

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)


def variant_model(region_len, n_hap, seed, density=1.0 / 40.0, indel_frac=0.1, device="cpu"):
    """Reference codes + variant sites.  Returns a dict of tensors on `device`."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    pad = 64
    ref = torch.randint(0, 4, (region_len + pad,), generator=g, dtype=torch.int64)
    n_var = max(1, int(region_len * density))
    site = torch.randperm(region_len - 1, generator=g)[:n_var].sort().values + 1
    is_indel = torch.rand(n_var, generator=g) < indel_frac
    length = torch.randint(1, 6, (n_var,), generator=g)
    sign = torch.where(torch.rand(n_var, generator=g) < 0.5, -1, 1)
    lo, hi = 1.0 / n_hap, 0.5
    af = lo * (hi / lo) ** torch.rand(n_var, generator=g, dtype=torch.float64)  # density ~ 1/x
    alt = (ref[site] + torch.randint(1, 4, (n_var,), generator=g)) % 4
    m = dict(ref=ref, site=site, is_indel=is_indel, shift=(length * sign) * is_indel, af=af.float(), alt=alt,
             region_len=region_len, pad=pad, seed=int(seed), n_hap=n_hap)
    return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in m.items()}


def haplotype_codes(model, hap_lo, hap_hi, return_index=False):
    """int64 [hap_hi-hap_lo, region_len] base codes of the haplotypes (deterministic per haplotype index).
    With return_index also the reference coordinate of every haplotype base (indels shift it)."""
    dev = model["ref"].device
    L, nh = model["region_len"], hap_hi - hap_lo
    n_var = model["site"].shape[0]
    g = torch.Generator(device=dev)
    g.manual_seed(model["seed"] * 1000003 + hap_lo)
    carry = torch.rand((nh, n_var), generator=g, device=dev) < model["af"][None, :]
    # indels shift the reference coordinate of everything downstream
    delta = torch.zeros((nh, L), dtype=torch.int32, device=dev)
    delta[:, model["site"]] = (carry * model["shift"][None, :]).to(torch.int32)
    idx = torch.arange(L, device=dev)[None, :] + torch.cumsum(delta, dim=1)
    idx.clamp_(0, L + model["pad"] - 1)
    hap = model["ref"][idx]
    # SNP alleles, placed in reference coordinates and carried through the same index map
    snp = torch.full((nh, L + model["pad"]), -1, dtype=torch.int64, device=dev)
    snp_sites = model["site"][~model["is_indel"]]
    snp[:, snp_sites] = torch.where(carry[:, ~model["is_indel"]], model["alt"][~model["is_indel"]][None, :], -1)
    alt = torch.gather(snp, 1, idx)
    codes = torch.where(alt >= 0, alt, hap)
    return (codes, idx) if return_index else codes


def pack_windows(codes, w):
    """int64 [H, L] codes -> int64 [H, L-w+1] packed windows (base i in bits 2i, 2i+1)."""
    n = codes.shape[1] - w + 1
    out = torch.zeros((codes.shape[0], n), dtype=torch.int64, device=codes.device)
    for j in range(w):
        out |= codes[:, j:j + n] << (2 * j)
    return out


def haplotype_windows(region_len, n_hap, w, seed, device="cpu", hap_batch=32, out=None):
    """Packed windows of every haplotype, flattened haplotype-major: int64 [n_hap * (region_len - w + 1)]."""
    model = variant_model(region_len, n_hap, seed, device=device)
    per = region_len - w + 1
    if out is None:
        out = torch.empty(n_hap * per, dtype=torch.int64, device=device)
    for lo in range(0, n_hap, hap_batch):
        hi = min(lo + hap_batch, n_hap)
        out[lo * per:hi * per] = pack_windows(haplotype_codes(model, lo, hi), w).reshape(-1)
    return out, model


def pack_codes_2bit(codes):
    """int64 [H, L] base codes -> int64 [H, ceil(L/32)] words of the 2-bit sequence layout (base i of a row in bits
    2(i & 31) of word i >> 5; include/grafimo_b200.h, "K2 over sequences")."""
    H, L = codes.shape
    nw = (L + 31) // 32
    pad = nw * 32 - L
    if pad:
        codes = torch.cat([codes, torch.zeros((H, pad), dtype=codes.dtype, device=codes.device)], dim=1)
    sh = 2 * torch.arange(32, device=codes.device, dtype=torch.int64)
    return ((codes.view(H, nw, 32) & 3) << sh[None, None, :]).sum(dim=2)  # disjoint bit fields: the sum is an OR


def haplotype_sequences(region_len, n_hap, seed, device="cpu", hap_batch=32, ascii_out=None):
    """The same haplotypes as `haplotype_windows` (same seed -> same bases) as whole sequences: int64 [n_hap, ceil(L/32)]
    2-bit words on `device`; when `ascii_out` (uint8 [n_hap, region_len], e.g. pinned host memory) is given it receives
    the ASCII letters as well."""
    model = variant_model(region_len, n_hap, seed, device=device)
    nw = (region_len + 31) // 32
    words = torch.empty((n_hap, nw), dtype=torch.int64, device=device)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=device)
    for lo in range(0, n_hap, hap_batch):
        hi = min(lo + hap_batch, n_hap)
        codes = haplotype_codes(model, lo, hi)
        words[lo:hi] = pack_codes_2bit(codes)
        if ascii_out is not None:
            ascii_out[lo:hi].copy_(lut[codes], non_blocking=True)
    return words, model


def windows_to_ascii(packed, w, out=None):
    """int64 [n] packed windows -> uint8 [n, w] ASCII k-mers (on the tensor's device)."""
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=packed.device)
    sh = 2 * torch.arange(w, device=packed.device, dtype=torch.int64)
    codes = (packed[:, None] >> sh[None, :]) & 3
    a = lut[codes]
    if out is not None:
        out.copy_(a)
        return out
    return a


def revcomp_ascii(a):
    """uint8 [n, w] ASCII -> reverse complement (numpy)."""
    comp = np.zeros(256, dtype=np.uint8)
    comp[:] = ord("N")
    for x, y in zip(b"ACGTacgt", b"TGCAtgca"):
        comp[x] = y
    return np.ascontiguousarray(comp[a][:, ::-1])


def reference_windows(model, w):
    """Packed windows of the unmodified reference (for the ref / non.ref flag of the tally)."""
    return pack_windows(model["ref"][None, :model["region_len"]], w).reshape(-1)


def synthetic_meme_collection(n_motifs=800, seed=20242, wmin=6, wmax=30, alpha=0.5):
    """A JASPAR-CORE-sized MEME file (text) with seeded random motifs: widths in [wmin, wmax] with mode ~11,
    Dirichlet(alpha) columns, nsites 100-5000 (JASPAR itself is not available offline)."""
    rng = np.random.default_rng(seed)
    widths = np.clip(np.round(rng.gamma(shape=6.0, scale=2.0, size=n_motifs)).astype(int) + 1, wmin, wmax)
    lines = ["MEME version 4", "", "ALPHABET= ACGT", "", "strands: + -", "", "Background letter frequencies",
             "A 0.25 C 0.25 G 0.25 T 0.25", ""]
    for k, w in enumerate(widths):
        nsites = int(rng.integers(100, 5001))
        lines.append(f"MOTIF SYN{k:04d}.1 SYN{k:04d}")
        lines.append(f"letter-probability matrix: alength= 4 w= {w} nsites= {nsites} E= 0")
        for row in rng.dirichlet([alpha] * 4, size=int(w)):
            lines.append(" " + "  ".join(f"{v:.6f}" for v in row))
        lines.append("URL none")
        lines.append("")
    return "\n".join(lines), widths


def variant_set(region_len, n_hap, seed, density=1.0 / 40.0, indel_frac=0.1, max_indel=5):
    """Seeded phased variant set of the C2 shape for the graph path (vgraph.VariationGraph.build): a random reference,
    variant sites at `density`, `indel_frac` of them insertions/deletions of 1..max_indel bp, allele frequency ~ 1/x
    on [1/n_hap, 0.5], every haplotype carrying each variant with probability af.
    -> (reference str, [(pos0, ref_allele, alt_allele)], uint8 [n_variants, n_hap])."""
    rng = np.random.default_rng(int(seed))
    codes = rng.integers(0, 4, size=region_len, dtype=np.uint8)
    ref = _ASCII[codes].tobytes().decode("ascii")
    n_var = max(1, int(region_len * density))
    pos = np.sort(rng.choice(np.arange(1, region_len - max_indel - 1), size=n_var, replace=False))
    kind = rng.random(n_var)
    length = rng.integers(1, max_indel + 1, size=n_var)
    alt_shift = rng.integers(1, 4, size=n_var)
    ins = rng.integers(0, 4, size=(n_var, max_indel), dtype=np.uint8)
    variants, last_end = [], 0
    keep = np.zeros(n_var, dtype=bool)
    for i in range(n_var):
        p = int(pos[i])
        if p < last_end:  # keep the set free of overlapping alleles
            continue
        if kind[i] >= indel_frac:
            r, a = ref[p], "ACGT"[(codes[p] + alt_shift[i]) % 4]
        elif kind[i] < indel_frac / 2:
            r, a = "", _ASCII[ins[i, :length[i]]].tobytes().decode("ascii")
        else:
            r, a = ref[p:p + int(length[i])], ""
        variants.append((p, r, a))
        keep[i] = True
        last_end = p + max(len(r), 1)
    lo, hi = 1.0 / n_hap, 0.5
    af = lo * (hi / lo) ** rng.random(len(variants))
    gt = (rng.random((len(variants), n_hap), dtype=np.float32) < af[:, None].astype(np.float32)).astype(np.uint8)
    return ref, variants, gt


def variant_arrays(region_len, n_hap, seed, density=1.0 / 40.0, indel_frac=0.1, max_indel=5, device="cpu"):
    """Array form of `variant_set` for chromosome-sized inputs (no per-variant Python): the layout gb2_graph_build takes.
    -> (reference uint8 ASCII numpy, {pos int64, ref_len int32, alt_off int64[n+1], alt uint8 ASCII}, (gt_bits uint32
    [n, words], n_hap)).  Genotype bits are drawn on `device` (a CUDA device makes 10^6 variants x 5,008 haplotypes a
    matter of seconds) with allele frequency ~ 1/x on [1/n_hap, 0.5]; alleles never overlap."""
    rng = np.random.default_rng(int(seed))
    codes = rng.integers(0, 4, size=region_len, dtype=np.uint8)
    ref = _ASCII[codes]
    n_var = max(1, int(region_len * density))
    pos = np.unique(rng.integers(1, region_len - max_indel - 1, size=n_var))
    pos = pos[np.concatenate([[True], np.diff(pos) > max_indel + 1])]
    n = len(pos)
    kind = rng.random(n)
    length = rng.integers(1, max_indel + 1, size=n)
    is_snp, is_ins = kind >= indel_frac, kind < indel_frac / 2
    is_del = ~is_snp & ~is_ins
    ref_len = np.where(is_snp, 1, np.where(is_del, length, 0)).astype(np.int32)
    alt_len = np.where(is_snp, 1, np.where(is_ins, length, 0)).astype(np.int64)
    alt_off = np.concatenate([[0], np.cumsum(alt_len)]).astype(np.int64)
    alt_codes = rng.integers(0, 4, size=int(alt_off[-1]), dtype=np.uint8)
    snp_alt = (codes[pos[is_snp]] + rng.integers(1, 4, size=int(is_snp.sum()), dtype=np.uint8)) % 4
    alt_codes[alt_off[:-1][is_snp]] = snp_alt
    alt = _ASCII[alt_codes]
    lo, hi = 1.0 / n_hap, 0.5
    af = torch.from_numpy(lo * (hi / lo) ** rng.random(n)).to(device=device, dtype=torch.float32)
    words = max(4, ((n_hap + 31) // 32 + 3) // 4 * 4)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed) + 12345)
    weights = (2 ** torch.arange(32, device=device, dtype=torch.int64))[None, None, :]
    bits = torch.empty((n, words), dtype=torch.int32, device="cpu")
    chunk = max(1, (1 << 28) // (words * 32))
    valid = (torch.arange(words * 32, device=device) < n_hap)[None, :]
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        carry = (torch.rand((b - a, words * 32), generator=g, device=device) < af[a:b, None]) & valid
        packed = (carry.view(b - a, words, 32).to(torch.int64) * weights).sum(dim=2)
        bits[a:b] = packed.to(torch.int32).cpu()  # low 32 bits
    variants = {"pos": pos.astype(np.int64), "ref_len": ref_len, "alt_off": alt_off, "alt": alt}
    return ref, variants, (bits.numpy().view(np.uint32), n_hap)
