"""GPU parity tests: every CUDA kernel against the CPU oracle / the reference-generated goldens.
All calls go through the C ABI (grafimo_b200._lib via grafimo_b200.engine).  Integer / byte / index
results and fp64 p-values, scores and q-values are compared BIT-EXACT (q-values are additionally
allowed 1e-12 relative by the spec; we assert equality)."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    from grafimo_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _orc():
    from oracle import oracle as orc
    return orc


def _ascii_dev(seqs, w):
    a = _orc().kmers_to_matrix(seqs, w)
    return a, torch.from_numpy(a).cuda()


# ---------------------------------------------------------------------------------------------------
def test_k3_batched_dp_bit_exact(ctx):
    motifs = [gu.load_motif(t) for t in gu.motif_tags()]
    outs = ctx.pval_dp_batched([m["score_matrix"] for m in motifs], [m["bg_acgt"] for m in motifs])
    for m, o in zip(motifs, outs):
        assert o.shape == m["pval_mat"].shape
        assert np.array_equal(o, m["pval_mat"]), m["tag"]


@pytest.mark.parametrize("tag", ["ctcf_meme__unif", "ctcf_meme__bgnt", "synth_w6_meme__bgnt", "synth_w30_meme__bgnt",
                                 "synth_w32_meme__bgnt", "synth_w33_meme__bgnt", "synth_w48_meme__bgnt", "synth_w64_meme__bgnt"])
def test_k4_ptable_bit_exact(ctx, tag):
    m = gu.load_motif(tag)
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    tab = _orc().pvalue_table(m["pval_mat"])
    assert np.array_equal(dm.ptable, tab[dm.lo:dm.hi + 1])
    assert dm.ptable[0] == 1.0
    nz = np.nonzero(m["pval_mat"])[0]
    assert dm.lo == nz[0] and dm.hi == nz[-1]


def test_k1_encoder(ctx):
    rng = np.random.default_rng(5)
    for w in (1, 5, 19, 31, 32, 33, 36, 47, 63, 64):  # above 32: two packed words per k-mer
        n = 1000 + w
        letters = np.array(list("ACGTacgt"))
        seqs = ["".join(rng.choice(letters, size=w)) for _ in range(n)]
        seqs[3] = "N" * w
        seqs[40] = seqs[40][:-1] + "N"
        seqs[77] = "x" + seqs[77][1:]
        seqs[n - 1] = seqs[n - 1][: w // 2] + "n" + seqs[n - 1][w // 2 + 1:]
        a, d = _ascii_dev(seqs, w)
        packed, nmask, counts = ctx.encode(d)
        ctx.sync()
        assert tuple(packed.shape) == ((n, 2) if w > 32 else (n,))
        packed = packed.cpu().numpy().view(np.uint64)
        nm = nmask.cpu().numpy().view(np.uint32)
        code = {"A": 0, "C": 1, "G": 2, "T": 3}
        for r, s in enumerate(seqs):
            bad = any(ch.upper() not in code for ch in s)
            assert ((nm[r >> 5] >> (r & 31)) & 1) == int(bad), (w, r)
            if not bad:
                x = 0
                for i, ch in enumerate(s):
                    x |= code[ch.upper()] << (2 * i)
                if w > 32:
                    assert int(packed[r, 0]) == x & ((1 << 64) - 1) and int(packed[r, 1]) == x >> 64
                else:
                    assert int(packed[r]) == x
        assert counts.cpu().numpy().tolist() == [4, 2]


def _score_case(ctx, m, seqs, strands, threshold, want_q=True, q_filter=False):
    from grafimo_b200.engine import Scan
    w = m["width"]
    a, d = _ascii_dev(seqs, w)
    packed, nmask, counts = ctx.encode(d)
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    sc = Scan(ctx, dm, strands=strands, threshold=threshold, want_q=want_q, hit_capacity=2 * len(seqs) + 8)
    sc.score(packed, nmask)
    out = sc.finalize(q_filter=q_filter)
    hist = sc.histogram().cpu().numpy() if want_q else None
    return dm, out, hist


@pytest.mark.parametrize("tag", gu.scoring_tags())
def test_k2_k5_k6_against_reference_goldens(ctx, tag):
    """Rows of the golden case scored as given (strands=1: vg already emits the '-' rows)."""
    orc = _orc()
    c = gu.load_scoring(tag)
    m = gu.load_motif(c["motif_tag"])
    o = c["options"]
    lines = [ln for f in c["files"] for ln in f]
    r = orc.parse_rows(lines, o["noreverse"])
    dm, out, hist = _score_case(ctx, m, r["seq"], 1, o["threshold"], want_q=not o["noqvalue"], q_filter=o["qvalueT"])
    # expected from the reference-generated table (recomb filter applied on the host in the product, so undo it here)
    exp = orc.compute_results(m, lines, threshold=o["threshold"], noqvalue=o["noqvalue"], qvalueT=o["qvalueT"],
                              noreverse=o["noreverse"], recomb=True)
    rows = out["row"].astype(np.int64)
    got = {
        "start": r["start"][rows], "stop": r["stop"][rows], "strand": np.array(r["strand"], dtype=object)[rows],
        "score": out["score"], "p-value": out["p-value"],
        "matched_sequence": np.array(r["seq"], dtype=object)[rows], "haplotype_frequency": r["freq"][rows],
        "_int_score": out["int_score"].astype(np.int64),
    }
    cols = ["start", "stop", "strand", "score", "p-value", "matched_sequence", "haplotype_frequency", "_int_score"]
    if not o["noqvalue"]:
        got["q-value"] = out["q-value"]
        cols.append("q-value")
        assert out["total"] == len(r["seq"])
        assert int(hist.sum()) == len(r["seq"])
    gu.assert_tables_equal(got, exp, cols)
    assert np.all(np.diff(out["p-value"]) >= 0)  # sorted by p ascending
    # and the golden itself when recomb was on
    if o["recomb"]:
        assert len(c["table"]["start"]) == len(rows)


def test_k2_both_strands_equals_reverse_complement_rows(ctx):
    """strands=2 on forward k-mers == the reference scoring the k-mer and its reverse complement row."""
    orc = _orc()
    rng = np.random.default_rng(11)
    for tag, n in (("ctcf_meme__bgnt", 20001), ("synth_w30_meme__bgnt", 4097), ("synth_w6_meme__bgnt", 3000),
                   ("synth_w32_meme__bgnt", 2049), ("synth_w33_meme__bgnt", 4100), ("synth_w35_meme__bgnt", 3001),
                   ("synth_w48_meme__bgnt", 1025), ("synth_w64_meme__bgnt", 1500)):
        m = gu.load_motif(tag)
        w = m["width"]
        seqs = ["".join(rng.choice(list("ACGT"), size=w)) for _ in range(n)]
        for k in (0, 5, n - 1):
            seqs[k] = seqs[k][:w // 2] + "N" + seqs[k][w // 2 + 1:]
        comp = str.maketrans("ACGTN", "TGCAN")
        rc = [s.translate(comp)[::-1] for s in seqs]
        thr = 0.02
        dm, out, hist = _score_case(ctx, m, seqs, 2, thr)
        a_f = orc.kmers_to_matrix(seqs, w)
        a_r = orc.kmers_to_matrix(rc, w)
        isf, lof, pf = orc.score_rows(a_f, m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"], 4)
        isr, lor, pr = orc.score_rows(a_r, m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"], 4)
        p_all = np.concatenate([pf, pr])
        q_all = orc.bh(p_all)
        exp_keep = np.nonzero(p_all < thr)[0]
        got_idx = out["row"].astype(np.int64) + n * out["strand"].astype(np.int64)
        assert sorted(got_idx.tolist()) == sorted(exp_keep.tolist())
        is_all = np.concatenate([isf, isr]); lo_all = np.concatenate([lof, lor])
        assert np.array_equal(out["int_score"], is_all[got_idx])
        assert np.array_equal(out["score"], lo_all[got_idx])
        assert np.array_equal(out["p-value"], p_all[got_idx])
        assert np.array_equal(out["q-value"], q_all[got_idx])
        # histogram == bincount of the oracle's integer scores (N rows in the last bin)
        isn = np.array(["N" in s for s in seqs])
        exp_hist = np.bincount(np.concatenate([isf[~isn], isr[~isn]]) - dm.lo, minlength=dm.span + 1)
        exp_hist[dm.span] = 2 * isn.sum()
        assert np.array_equal(hist, exp_hist)


@pytest.mark.parametrize("tag", ["synth_w25_meme__bgnt", "synth_w35_meme__bgnt"])
def test_dense_output_and_no_hist(ctx, tag):
    from grafimo_b200.engine import Scan
    orc = _orc()
    m = gu.load_motif(tag)
    w = m["width"]
    rng = np.random.default_rng(3)
    n = 5003
    seqs = ["".join(rng.choice(list("ACGT"), size=w)) for _ in range(n)]
    seqs[17] = "N" + seqs[17][1:]
    a, d = _ascii_dev(seqs, w)
    packed, nmask, _ = ctx.encode(d)
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    sc = Scan(ctx, dm, strands=2, threshold=1e-3, want_q=False, hit_capacity=64)
    dense = ctx.empty(n + 1, torch.int32)
    sc.score(packed, nmask, dense_out=dense)
    ctx.sync()
    dn = dense.cpu().numpy().view(np.uint32)[:n]
    isf, _, _ = orc.score_rows(a, m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"], 2, want_p=False)
    comp = str.maketrans("ACGTN", "TGCAN")
    a_r = orc.kmers_to_matrix([s.translate(comp)[::-1] for s in seqs], w)
    isr, _, _ = orc.score_rows(a_r, m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"], 2, want_p=False)
    ok = np.ones(n, bool); ok[17] = False
    assert dn[17] == 0xFFFFFFFF
    assert np.array_equal((dn[ok] & 0xFFFF).astype(np.int64) + dm.lo, isf[ok])
    assert np.array_equal((dn[ok] >> 16).astype(np.int64) + dm.lo, isr[ok])


@pytest.mark.parametrize("tag", ["ctcf_meme__bgnt", "synth_w35_meme__bgnt"])
def test_dense_finalize_equals_hit_path(ctx, tag):
    """Unselective thresholds: dense scores + gb2_finalize_dense == hit records + gb2_finalize_hits, row for row (same
    (p, row, strand) order), for one and two strands, a q-value filter, N rows and batches that begin at odd rows."""
    from grafimo_b200.engine import Scan
    orc = _orc()
    m = gu.load_motif(tag)
    w = m["width"]
    rng = np.random.default_rng(9)
    n = 20001
    seqs = ["".join(rng.choice(list("ACGT"), size=w)) for _ in range(n)]
    for k in (0, 7000, 7001, n - 1):
        seqs[k] = seqs[k][:3] + "N" + seqs[k][4:]
    a = orc.kmers_to_matrix(seqs, w)
    packed, nmask, _ = ctx.encode(torch.from_numpy(a).cuda())
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    for strands, thr, qf in ((2, 1.0, False), (1, 0.3, False), (2, 0.9, True)):
        ref = Scan(ctx, dm, strands=strands, threshold=thr, hit_capacity=2 * n)
        ref.score(packed, nmask, row_base=100)
        exp = ref.finalize(q_filter=qf)
        den = Scan(ctx, dm, strands=strands, threshold=thr, dense_rows=n)
        cut = 7008  # the N mask of a batch starts at its own bit 0: split on a multiple of 32
        den.score(packed[:cut], nmask[:cut // 32], row_base=100)
        den.score(packed[cut:], nmask[cut // 32:], row_base=100 + cut)
        got = den.finalize(q_filter=qf)
        assert len(exp["row"]) > 100
        for k in exp:
            assert np.array_equal(np.asarray(got[k]), np.asarray(exp[k])), (strands, thr, qf, k)
    # ... and an odd first batch (the dense pairs of the second batch are then 4-byte aligned only)
    a2 = orc.kmers_to_matrix([s.replace("N", "A") for s in seqs], w)
    packed2, _, _ = ctx.encode(torch.from_numpy(a2).cuda())
    ref = Scan(ctx, dm, strands=2, threshold=1.0, hit_capacity=2 * n)
    ref.score(packed2)
    exp = ref.finalize()
    den = Scan(ctx, dm, strands=2, threshold=1.0, dense_rows=n)
    den.score(packed2[:7001].clone())
    den.score(packed2[7001:].clone(), row_base=7001)  # clone: a 16-byte aligned k-mer buffer of its own
    got = den.finalize()
    for k in exp:
        assert np.array_equal(np.asarray(got[k]), np.asarray(exp[k])), k


@pytest.mark.parametrize("tag,n", [("synth_w25_meme__bgnt", 2_000_003), ("synth_w8_meme__bgnt", 300_001)])
def test_dense_partition_passes_equal_library_sort(ctx, tag, n, monkeypatch):
    """gb2_finalize_dense (two hand-written stable partition passes, csrc/dense_sort.cu) == the library-sort form it replaced
    (GB2_DENSE_CUB=1), column for column, over many tiles: threshold 1, a threshold that drops windows, a q-value filter,
    one strand, N rows."""
    from grafimo_b200.engine import Scan
    m = gu.load_motif(tag)
    w = m["width"]
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    g = torch.Generator(device="cuda"); g.manual_seed(31)
    packed = torch.randint(0, 1 << 62, (n,), dtype=torch.int64, device="cuda", generator=g) & ((1 << (2 * w)) - 1)
    nmask = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    nmask[::97] = 0x00010400  # a few N rows
    for strands, thr, qf in ((2, 1.0, False), (2, 0.4, False), (1, 1.0, False), (2, 0.97, True)):
        out = {}
        for form in ("cub", "passes", "passes+table"):  # "+table": bin -> rank through the shared-memory table even if p is monotone
            monkeypatch.setenv("GB2_DENSE_CUB", "1" if form == "cub" else "0")
            monkeypatch.setenv("GB2_DENSE_RANK_TABLE", "1" if form.endswith("table") else "0")
            sc = Scan(ctx, dm, strands=strands, threshold=thr, dense_rows=n)
            sc.score(packed, nmask, row_base=5)
            out[form] = sc.finalize(q_filter=qf)
        assert qf or len(out["cub"]["row"]) > n // 4
        for form in ("passes", "passes+table"):
            for k in out["cub"]:
                assert np.array_equal(np.asarray(out[form][k]), np.asarray(out["cub"][k])), (tag, form, strands, thr, qf, k)


def test_scan_host_wide_motif(ctx):
    """gb2_scan_host with a 48-bp motif: two packed words per k-mer through the chunked encode + score loop."""
    from grafimo_b200.engine import scan_host
    orc = _orc()
    m = gu.load_motif("synth_w48_meme__bgnt")
    rng = np.random.default_rng(48)
    n, w = 3001, 48
    seqs = ["".join(rng.choice(list("ACGT"), size=w)) for _ in range(n)]
    seqs[5] = seqs[5][:40] + "N" + seqs[5][41:]
    a = orc.kmers_to_matrix(seqs, w)
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    out = scan_host(ctx, dm, a, strands=1, threshold=0.2)
    isc, lo, pv = orc.score_rows(a, m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    q = orc.bh(pv)
    rows = out["row"].astype(np.int64)
    assert sorted(rows.tolist()) == np.nonzero(pv < 0.2)[0].tolist()
    assert np.array_equal(out["int_score"], isc[rows]) and np.array_equal(out["score"], lo[rows])
    assert np.array_equal(out["p-value"], pv[rows]) and np.array_equal(out["q-value"], q[rows])
    assert out["stats"]["n_rows"] == 1


def test_scan_host_matches_device_path(ctx):
    from grafimo_b200.engine import scan_host
    orc = _orc()
    c = gu.load_scoring("fixture_plus_N_2files")
    m = gu.load_motif(c["motif_tag"])
    lines = [ln for f in c["files"] for ln in f]
    r = orc.parse_rows(lines)
    a = orc.kmers_to_matrix(r["seq"], 19)
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    out = scan_host(ctx, dm, a, strands=1, threshold=1.0)
    isc, lo, pv = orc.score_rows(a, m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    q = orc.bh(pv)
    rows = out["row"].astype(np.int64)
    assert sorted(rows.tolist()) == np.nonzero(pv < 1.0)[0].tolist()
    assert np.array_equal(out["p-value"], pv[rows]) and np.array_equal(out["q-value"], q[rows])
    assert np.array_equal(out["score"], lo[rows])
    assert out["stats"]["windows"] == len(r["seq"]) and out["stats"]["n_rows"] == 24


def test_hit_capacity_is_reported(ctx):
    from grafimo_b200.engine import scan_host
    from grafimo_b200._lib import GrafimoB200Error
    m = gu.load_motif("ctcf_meme__unif")
    a = _orc().kmers_to_matrix(["ACGTACGTACGTACGTACG"] * 100, 19)
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    with pytest.raises(GrafimoB200Error) as e:
        scan_host(ctx, dm, a, strands=2, threshold=1.0, hit_capacity=10)
    assert e.value.code == 4


def test_tally_haplotypes(ctx):
    rng = np.random.default_rng(9)
    n_pos, n_hap = 300, 64
    ref = rng.integers(0, 1 << 38, size=n_pos, dtype=np.int64)
    pos = np.repeat(np.arange(n_pos, dtype=np.int64) + 1000, n_hap)
    packed = np.repeat(ref, n_hap)
    alt = rng.random(pos.shape[0]) < 0.2
    packed[alt] = packed[alt] ^ rng.integers(1, 4, size=alt.sum())
    perm = rng.permutation(pos.shape[0])
    pos, packed = pos[perm], packed[perm]
    u_pos, u_packed, u_freq, u_isref = ctx.tally_haplotypes(torch.from_numpy(pos).cuda(), torch.from_numpy(packed).cuda(),
                                                            torch.from_numpy(ref).cuda(), pos_base=1000)
    got = sorted(zip(u_pos.cpu().tolist(), u_packed.cpu().tolist(), u_freq.cpu().tolist(), u_isref.cpu().tolist()))
    import collections
    cnt = collections.Counter(zip(pos.tolist(), packed.tolist()))
    exp = sorted((p, k, c, int(ref[p - 1000] == k)) for (p, k), c in cnt.items())
    assert got == exp


def test_large_batch_properties(ctx):
    """Size-independent properties at a size the oracle cannot score row by row: histogram mass,
    hit count == histogram tail, strand symmetry of a reverse-complemented batch."""
    from grafimo_b200.engine import Scan
    m = gu.load_motif("ctcf_meme__unif")
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    n = (1 << 24) + 3
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    packed = torch.randint(0, 1 << 38, (n,), dtype=torch.int64, device="cuda", generator=g)
    thr = 1e-4
    sc = Scan(ctx, dm, strands=2, threshold=thr, hit_capacity=1 << 16)
    sc.score(packed)
    out = sc.finalize()
    hist = sc.histogram().cpu().numpy()
    assert int(hist.sum()) == 2 * n and out["total"] == 2 * n
    cut = np.nonzero(dm.ptable < thr)[0][0]
    assert int(hist[cut:dm.span].sum()) == len(out["row"])
    # reverse-complement the batch on the device: fwd/rc scores swap, so the histogram is identical
    x = packed
    rcx = torch.zeros_like(x)
    for i in range(19):
        rcx |= ((3 - ((x >> (2 * i)) & 3)) << (2 * (18 - i)))
    sc2 = Scan(ctx, dm, strands=2, threshold=thr, hit_capacity=1 << 16)
    sc2.score(rcx)
    out2 = sc2.finalize()
    assert np.array_equal(sc2.histogram().cpu().numpy(), hist)
    a = sorted(zip(out["row"].tolist(), out["strand"].tolist(), out["int_score"].tolist()))
    b = sorted(zip(out2["row"].tolist(), (1 - out2["strand"]).tolist(), out2["int_score"].tolist()))
    assert a == b
    # spot-check 2000 random rows against the oracle
    orc = _orc()
    idx = np.random.default_rng(2).integers(0, n, 2000)
    xs = packed[torch.from_numpy(idx).cuda()].cpu().numpy()
    seqs = ["".join("ACGT"[(int(v) >> (2 * i)) & 3] for i in range(19)) for v in xs]
    isf, _, _ = orc.score_rows(orc.kmers_to_matrix(seqs, 19), m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"],
                               m["offset"], 4, want_p=False)
    dense = ctx.empty(n + 1, torch.int32)
    sc3 = Scan(ctx, dm, strands=2, threshold=thr, want_q=False, hit_capacity=1 << 16)
    sc3.score(packed, dense_out=dense)
    ctx.sync()
    dn = dense[torch.from_numpy(idx).cuda()].cpu().numpy().view(np.uint32)
    assert np.array_equal((dn & 0xFFFF).astype(np.int64) + dm.lo, isf)


def test_k1b_device_tsv_reader(ctx):
    """gb2_tsv_index_lines / gb2_tsv_parse_rows against the reference's field rules (oracle.parse_rows)."""
    orc = _orc()
    c = gu.load_scoring("fixture_plus_N_2files")
    lines = [ln for f in c["files"] for ln in f]
    # the same rows with awkward but legal whitespace: blanks instead of tabs, CRLF, leading blanks, empty lines
    odd = []
    for i, ln in enumerate(lines):
        f = ln.split("\t")
        if i % 5 == 1:
            ln = "  ".join(f)
        elif i % 5 == 2:
            ln = "\t" + ln + "\r"
        elif i % 5 == 3:
            ln = ln + "\n"  # followed by an empty line
        odd.append(ln)
    text = ("\n".join(odd) + "\n").encode()
    for skip in (False, True):
        rows = ctx.parse_kmer_tsv(torch.frombuffer(bytearray(text), dtype=torch.uint8), 19, skip_minus=skip)
        r = orc.parse_rows(lines, skip)
        assert rows.n == len(r["seq"])
        st = rows.stats()
        assert st["malformed"] == 0
        a = orc.kmers_to_matrix(r["seq"], 19)
        packed, nmask, counts = ctx.encode(torch.from_numpy(a).cuda())
        ctx.sync()
        assert torch.equal(rows.packed, packed) and torch.equal(rows.nmask, nmask)
        assert st["n_rows"] == int(counts[0]) and st["bad_rows"] == int(counts[1])
        assert np.array_equal(rows.start.cpu().numpy(), r["start"]) and np.array_equal(rows.stop.cpu().numpy(), r["stop"])
        assert np.array_equal(rows.freq.cpu().numpy(), r["freq"])
        assert [chr(x) for x in rows.strand.cpu().tolist()] == r["strand"]
        assert [("ref" if x == 1 else "non.ref") for x in rows.ref.cpu().tolist()] == r["ref"]
        off, nl, so = rows.line_off.cpu().numpy(), rows.name_len.cpu().numpy(), rows.seq_off.cpu().numpy()
        for k in range(0, rows.n, 37):
            line = text[off[k]:].split(b"\n", 1)[0]
            assert line.split()[0].decode() == r["seqname"][k] and len(r["seqname"][k]) == nl[k]
            assert text[off[k] + so[k]:off[k] + so[k] + 19].decode() == r["seq"][k]
    # k-mers wider than one packed word (reference-generated rows of the w = 35 golden)
    c = gu.load_scoring("synth_w35")
    lines = [ln for f in c["files"] for ln in f]
    text = ("\n".join(lines) + "\n").encode()
    rows = ctx.parse_kmer_tsv(torch.frombuffer(bytearray(text), dtype=torch.uint8), 35)
    r = orc.parse_rows(lines, False)
    assert rows.n == len(r["seq"]) and rows.stats()["malformed"] == 0
    packed, nmask, counts = ctx.encode(torch.from_numpy(orc.kmers_to_matrix(r["seq"], 35)).cuda())
    ctx.sync()
    assert tuple(rows.packed.shape) == (rows.n, 2)
    assert torch.equal(rows.packed, packed) and torch.equal(rows.nmask, nmask)
    assert np.array_equal(rows.start.cpu().numpy(), r["start"]) and np.array_equal(rows.stop.cpu().numpy(), r["stop"])
    # malformed lines are counted, not silently scored
    bad = b"1:1-9\tACGT\t1:1+\t1:5+\t3\tref\t1+,\n1:1-9\tACGTACGTACGTACGTACG\t1:1+\t1:20+\tx\tref\t1+,\n1:1-9\tACGTACGTACGTACGTACG\t1:1+\n"
    rows = ctx.parse_kmer_tsv(torch.frombuffer(bytearray(bad), dtype=torch.uint8), 19)
    assert rows.n == 3 and rows.stats()["malformed"] == 3
    empty = ctx.parse_kmer_tsv(torch.frombuffer(bytearray(b"\n\n  \n"), dtype=torch.uint8), 19)
    assert empty.n == 0
