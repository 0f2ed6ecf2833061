"""GPU parity tests of the sequence form of K2 (csrc/seqscan.cu): windows formed on the device from 2-bit / ASCII
sequences must give exactly what the k-mer form gives on the expanded list of windows -- and that form is pinned on the
oracle and the reference goldens (test_gpu_kernels.py).  One case is also checked against the oracle directly.
Everything goes through the C ABI; integer results and fp64 columns are compared bit for bit."""
import os

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


def _random_seqs(rng, lens, n_rate=0.002, lower=True, bad=False):
    out = []
    for n in lens:
        letters = np.array(list("ACGT"))
        s = rng.choice(letters, size=n)
        if lower and n:
            m = rng.random(n) < 0.1
            s = np.where(m, np.char.lower(s), s)
        if n and n_rate > 0:
            m = rng.random(n) < n_rate
            s = np.where(m, "N", s)
        s = "".join(s)
        if bad and n > 5:
            s = s[:3] + "x" + s[4:]
        out.append(s)
    return out


def _windows(seqs, w):
    """expanded windows (upper-cased, non-ACGT -> N) in (sequence, position) order"""
    out = []
    for s in seqs:
        u = "".join(ch if ch in "ACGT" else "N" for ch in s.upper())
        out.extend(u[i:i + w] for i in range(len(u) - w + 1))
    return out


def _layout_text(rng, seqs):
    """the sequences in one byte buffer at odd offsets with junk between them"""
    parts, offs, pos = [], [], 0
    for s in seqs:
        gap = int(rng.integers(0, 7))
        parts.append(b"#" * gap)
        pos += gap
        offs.append(pos)
        parts.append(s.encode("ascii"))
        pos += len(s)
    parts.append(b"##")
    return np.frombuffer(b"".join(parts), dtype=np.uint8).copy(), np.array(offs, dtype=np.int64)


LENS = [0, 1, 5, 18, 19, 20, 31, 32, 33, 63, 64, 65, 1023, 1024, 1025, 1042, 1043, 2048 + 18, 5000, 40000]


def test_sequence_encoder_matches_numpy_packer(ctx):
    from grafimo_b200.engine import SeqBatch, pack_sequences_2bit
    rng = np.random.default_rng(11)
    seqs = _random_seqs(rng, LENS, n_rate=0.01, bad=True)
    text, offs = _layout_text(rng, seqs)
    words, nbits, word_off, lens = pack_sequences_2bit(seqs)
    b, counts = SeqBatch.from_ascii(ctx, torch.from_numpy(text).cuda(), offs, lens)
    ctx.sync()
    assert np.array_equal(b.word_off, word_off)
    n = b.n_words
    assert np.array_equal(b.seq2.cpu().numpy().view(np.uint64)[:n], words[:n])
    assert np.array_equal(b.nbits.cpu().numpy().view(np.uint32)[:n], nbits[:n])
    n_bad = sum(ch not in "ACGTacgt" for s in seqs for ch in s)
    n_other = sum(ch not in "ACGTacgtNn" for s in seqs for ch in s)
    assert counts.cpu().numpy().tolist() == [n_bad, n_other]


def _kmer_scan(ctx, dm, wins, w, strands, threshold, dense=False):
    """the k-mer form (gb2_encode_kmers + gb2_score) on the expanded windows"""
    from grafimo_b200.engine import Scan
    a = _orc().kmers_to_matrix(wins, w)
    packed, nmask, _ = ctx.encode(torch.from_numpy(a).cuda())
    sc = Scan(ctx, dm, strands=strands, threshold=threshold, hit_capacity=2 * len(wins) + 8, dense_rows=len(wins) if dense else 0)
    sc.score(packed, nmask)
    return sc


@pytest.mark.parametrize("tag,strands,threshold", [("ctcf_meme__unif", 2, 1e-2), ("ctcf_meme__bgnt", 1, 0.05),
                                                    ("synth_w6_meme__bgnt", 2, 0.3), ("synth_w30_meme__bgnt", 2, 1e-2),
                                                    ("synth_w32_meme__bgnt", 2, 1e-2), ("synth_w8_meme__bgnt", 2, 0.5)])
def test_score_sequences_equals_kmer_form(ctx, tag, strands, threshold):
    from grafimo_b200.engine import Scan, SeqBatch
    m = gu.load_motif(tag)
    w = m["width"]
    rng = np.random.default_rng(w * 7 + strands)
    seqs = _random_seqs(rng, LENS)
    text, offs = _layout_text(rng, seqs)
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    b, _ = SeqBatch.from_ascii(ctx, torch.from_numpy(text).cuda(), offs, [len(s) for s in seqs])
    wins = _windows(seqs, w)
    assert b.n_windows(w) == len(wins)
    ref = _kmer_scan(ctx, dm, wins, w, strands, threshold)
    sc = Scan(ctx, dm, strands=strands, threshold=threshold, hit_capacity=2 * len(wins) + 8)
    assert sc.score_sequences(b) == len(wins)
    ctx.sync()  # the kernels run on the context's stream; .cpu() below runs on torch's
    assert np.array_equal(sc.histogram().cpu().numpy(), ref.histogram().cpu().numpy())
    got, exp = sc.finalize(), ref.finalize()
    for k in ("row", "strand", "int_score", "score", "p-value", "q-value"):
        assert np.array_equal(got[k], exp[k]), k
    assert got["total"] == exp["total"] == strands * len(wins)
    # dense scores, indexed by window
    dref = _kmer_scan(ctx, dm, wins, w, strands, 1.0, dense=True)
    dsc = Scan(ctx, dm, strands=strands, threshold=1.0, dense_rows=len(wins))
    dsc.score_sequences(b)
    ctx.sync()
    assert np.array_equal(dsc.dense.cpu().numpy()[:len(wins)], dref.dense.cpu().numpy()[:len(wins)])


def test_score_sequences_against_oracle(ctx):
    """direct check against the CPU oracle (not only against the k-mer kernel): CTCF, both strands, with N windows"""
    from grafimo_b200.engine import Scan, SeqBatch, pack_sequences_2bit
    orc = _orc()
    m = gu.load_motif("ctcf_meme__bgnt")
    w = m["width"]
    rng = np.random.default_rng(3)
    seqs = _random_seqs(rng, [700, 19, 18, 2500, 33], n_rate=0.004, lower=False)
    words, nbits, word_off, lens = pack_sequences_2bit(seqs)
    b = SeqBatch(ctx, lens, seq2=torch.from_numpy(words.view(np.int64)).cuda(), nbits=torch.from_numpy(nbits.view(np.int32)).cuda())
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    sc = Scan(ctx, dm, strands=2, threshold=0.02, hit_capacity=1 << 16)
    n = sc.score_sequences(b)
    out = sc.finalize()
    wins = _windows(seqs, w)
    assert n == len(wins)
    comp = str.maketrans("ACGTN", "TGCAN")
    a = orc.kmers_to_matrix(wins, w)
    rc = orc.kmers_to_matrix([s.translate(comp)[::-1] for s in wins], w)
    args = (m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    isf, lof, pf = orc.score_rows(a, *args)
    isr, lor, pr = orc.score_rows(rc, *args)
    p_all, lo_all, is_all = np.concatenate([pf, pr]), np.concatenate([lof, lor]), np.concatenate([isf, isr])
    q_all = orc.bh(p_all)
    idx = out["row"].astype(np.int64) + n * out["strand"].astype(np.int64)
    assert sorted(idx.tolist()) == np.nonzero(p_all < 0.02)[0].tolist()
    assert np.array_equal(out["p-value"], p_all[idx]) and np.array_equal(out["score"], lo_all[idx])
    assert np.array_equal(out["int_score"], is_all[idx]) and np.array_equal(out["q-value"], q_all[idx])
    assert any("N" in x for x in wins)


@pytest.mark.parametrize("chunk", [None, "4096"])
@pytest.mark.parametrize("fmt", ["ascii", "2bit"])
def test_scan_host_sequences_equals_scan_host(ctx, fmt, chunk, monkeypatch):
    """host entry: sequences (ASCII bytes / 2-bit words) -> the table gb2_scan_host gives on the expanded ASCII k-mers;
    with a tiny chunk size the long sequences are cut into overlapping pieces over many chunks"""
    from grafimo_b200 import engine
    if chunk:
        monkeypatch.setenv("GB2_SEQ_CHUNK_BASES", chunk)
    else:
        monkeypatch.delenv("GB2_SEQ_CHUNK_BASES", raising=False)
    m = gu.load_motif("ctcf_meme__unif")
    w = m["width"]
    rng = np.random.default_rng(29)
    seqs = _random_seqs(rng, LENS + [9000, 4096 + 18, 4097], n_rate=0.001 if fmt == "ascii" else 0.0005)
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    wins = _windows(seqs, w)
    exp = engine.scan_host(ctx, dm, _orc().kmers_to_matrix(wins, w), strands=2, threshold=0.01)
    if fmt == "ascii":
        text, offs = _layout_text(rng, seqs)
        got = engine.scan_host_sequences(ctx, dm, text, offs, [len(s) for s in seqs], fmt="ascii", strands=2, threshold=0.01)
        assert got["stats"]["n_bases"] == sum(ch not in "ACGTacgt" for s in seqs for ch in s)
    else:
        words, nbits, word_off, lens = engine.pack_sequences_2bit(seqs)
        got = engine.scan_host_sequences(ctx, dm, words, word_off, lens, fmt="2bit", nbits=nbits, strands=2, threshold=0.01)
    assert got["stats"]["windows"] == 2 * len(wins) and got["stats"]["hits"] == exp["stats"]["hits"] > 0
    for k in ("row", "strand", "int_score", "score", "p-value", "q-value"):
        assert np.array_equal(got[k], exp[k]), k



@pytest.mark.parametrize("n_rate", [0.002, 0.00002])  # nearly every 8 Ki-base chunk holds an N / hardly any does (its N bits stay home)
@pytest.mark.parametrize("threads", ["0", "1", "5"])
def test_scan_host_sequences_host_packers(ctx, threads, n_rate, monkeypatch):
    """ASCII sequences with part of the chunks re-coded to 2-bit words by host threads (csrc/host_pack.cpp, transfer
    compression) and the rest encoded on the device: table and counters are those of the device-only route (threads = 0),
    which the test above ties to gb2_scan_host -- also into a reusable pinned HostTable, with N, lower case and other symbols."""
    from grafimo_b200 import engine
    monkeypatch.setenv("GB2_SEQ_CHUNK_BASES", "8192")
    m = gu.load_motif("ctcf_meme__unif")
    w = m["width"]
    rng = np.random.default_rng(61)
    seqs = _random_seqs(rng, [70001, 33, 18, 19, 50000, 4096, 12345, 99999], n_rate=n_rate)
    seqs[3] = "acgtnACGTRYKMacgtacgTT"[:19]
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    text, offs = _layout_text(rng, seqs)
    lens = [len(s) for s in seqs]
    monkeypatch.setenv("GB2_HOST_PACK_THREADS", "0")
    exp = engine.scan_host_sequences(ctx, dm, text, offs, lens, fmt="ascii", strands=2, threshold=0.01)
    monkeypatch.setenv("GB2_HOST_PACK_THREADS", threads)
    table = engine.HostTable(1 << 16)
    for rep in range(3):  # the schedule (which chunks the packers take) differs from call to call
        got = engine.scan_host_sequences(ctx, dm, text, offs, lens, fmt="ascii", strands=2, threshold=0.01, hit_capacity=1 << 16, out=table)
        assert got["stats"] == exp["stats"] and got["stats"]["hits"] > 0
        for k in ("row", "strand", "int_score", "score", "p-value", "q-value"):
            assert np.array_equal(got[k], exp[k]), (threads, rep, k)



def test_scan_host_sequences_random_schedules(ctx, monkeypatch):
    """Random batches x random chunk sizes x random packer thread counts: the table and the counters never depend on how the
    batch was cut into chunks / pieces or on which chunks the host packers took."""
    from grafimo_b200 import engine
    m = gu.load_motif("ctcf_meme__unif")
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    rng = np.random.default_rng(2024)
    for trial in range(6):
        lens = [int(x) for x in rng.integers(0, 60000, size=int(rng.integers(3, 12)))] + [int(rng.integers(100000, 250000))]
        seqs = _random_seqs(rng, lens, n_rate=float(rng.choice([0.0, 0.00005, 0.003])), bad=bool(trial & 1))
        text, offs = _layout_text(rng, seqs)
        ln = [len(x) for x in seqs]
        monkeypatch.setenv("GB2_HOST_PACK_THREADS", "0")
        monkeypatch.delenv("GB2_SEQ_CHUNK_BASES", raising=False)
        exp = engine.scan_host_sequences(ctx, dm, text, offs, ln, fmt="ascii", strands=2, threshold=0.004)
        for rep in range(3):
            monkeypatch.setenv("GB2_HOST_PACK_THREADS", str(int(rng.integers(0, 8))))
            monkeypatch.setenv("GB2_SEQ_CHUNK_BASES", str(int(rng.choice([1024, 3000, 8192, 40000, 1 << 17]))))
            got = engine.scan_host_sequences(ctx, dm, text, offs, ln, fmt="ascii", strands=2, threshold=0.004)
            assert got["stats"] == exp["stats"], (trial, rep)
            for k in ("row", "strand", "int_score", "score", "p-value", "q-value"):
                assert np.array_equal(got[k], exp[k]), (trial, rep, k)


def test_scan_host_packed_equals_scan_host(ctx):
    from grafimo_b200 import engine
    for tag in ("ctcf_meme__unif", "synth_w35_meme__bgnt"):
        m = gu.load_motif(tag)
        w = m["width"]
        rng = np.random.default_rng(w)
        seqs = ["".join(rng.choice(list("ACGT"), size=w)) for _ in range(5003)]
        seqs[17] = "N" * w
        seqs[4000] = seqs[4000][:-1] + "n"
        a = _orc().kmers_to_matrix(seqs, w)
        dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
        exp = engine.scan_host(ctx, dm, a, strands=2, threshold=0.05)
        packed, nmask, _ = ctx.encode(torch.from_numpy(a).cuda())
        ctx.sync()
        got = engine.scan_host_packed(ctx, dm, packed.cpu().numpy(), nmask.cpu().numpy(), strands=2, threshold=0.05)
        assert len(got["row"]) == len(exp["row"]) > 0
        for k in ("row", "strand", "int_score", "score", "p-value", "q-value"):
            assert np.array_equal(got[k], exp[k]), (tag, k)


def test_sequence_entry_argument_errors(ctx):
    from grafimo_b200 import engine
    from grafimo_b200._lib import GrafimoB200Error
    m = gu.load_motif("synth_w35_meme__bgnt")
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    with pytest.raises(GrafimoB200Error):  # wider than one packed word: the k-mer form is the route
        engine.scan_host_sequences(ctx, dm, np.frombuffer(b"ACGT" * 20, dtype=np.uint8).copy(), [0], [80])
    m = gu.load_motif("ctcf_meme__unif")
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    out = engine.scan_host_sequences(ctx, dm, np.frombuffer(b"ACGT", dtype=np.uint8).copy(), [0], [4], strands=2)
    assert len(out["row"]) == 0 and out["stats"]["windows"] == 0  # shorter than the motif: nothing to score
