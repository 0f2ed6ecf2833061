"""Thin Python objects over the C ABI (include/grafimo_b200.h).

torch is used for what it is good at here -- device memory, streams and torch.distributed -- and
nothing else: every computation is a call into libgrafimo_b200.so with raw pointers.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import GrafimoB200Error, Hit, MotifInfo, check

_HIT_BYTES = ctypes.sizeof(Hit)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def packed_rows(ctx, n, width):
    """Device buffer for n packed k-mers: int64[n] (width <= 32) or int64[n, 2] = {bases 0..31, bases 32..} (wider),
    16-byte aligned either way (include/grafimo_b200.h, "Data layout")."""
    m = max(int(n), 1)
    if width > _lib.NARROW_WIDTH:
        return ctx.empty(2 * m, torch.int64).view(m, 2)
    return ctx.empty(m + (m & 1), torch.int64)[:m]


class Context:
    """One per (process, GPU): owns a CUDA stream and the library context bound to it."""

    def __init__(self, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available() or self.lib.gb2_device_count() == 0:
            raise GrafimoB200Error(2, "grafimo_b200.Context", "no CUDA device: grafimo_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.Stream(device=self.device)
        h = ctypes.c_void_p()
        check(self.lib.gb2_ctx_create(self.device.index, ctypes.c_void_p(self.stream.cuda_stream), ctypes.byref(h)),
              "gb2_ctx_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.gb2_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ------------------------------------------------------------------------------
    def enter(self):
        """Order this context's stream after the caller's current stream."""
        self.stream.wait_stream(torch.cuda.current_stream(self.device))

    def leave(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def sync(self):
        check(self.lib.gb2_ctx_sync(self.h), "gb2_ctx_sync", self.h)

    @property
    def launches(self):
        return int(self.lib.gb2_ctx_launch_count(self.h))

    @property
    def sm_count(self):
        return int(self.lib.gb2_ctx_sm_count(self.h))

    def empty(self, n, dtype):
        with torch.cuda.stream(self.stream):
            return torch.empty(int(n), dtype=dtype, device=self.device)

    def zeros(self, n, dtype):
        with torch.cuda.stream(self.stream):
            return torch.zeros(int(n), dtype=dtype, device=self.device)

    def last_transfer(self):
        """Bytes the last scan_host* call on this context moved over PCIe and how its chunks travelled (gb2_scan_last_transfer)."""
        v = [ctypes.c_uint64(0) for _ in range(4)]
        check(self.lib.gb2_scan_last_transfer(self.h, *[ctypes.byref(x) for x in v]), "gb2_scan_last_transfer", self.h)
        return dict(h2d_bytes=int(v[0].value), d2h_bytes=int(v[1].value), chunks_as_given=int(v[2].value), chunks_host_packed=int(v[3].value))

    # -- multi-GPU: NCCL communicator owned by the library context (csrc/comm.cu) -------------------
    world = 1
    rank = 0

    def comm_unique_id(self):
        buf = (ctypes.c_uint8 * _lib.NCCL_ID_BYTES)()
        check(self.lib.gb2_comm_unique_id(buf), "gb2_comm_unique_id", self.h)
        return bytes(buf)

    def comm_init(self, unique_id, rank, world):
        """Collective over the `world` contexts of a run (one per GPU); `unique_id` = rank 0's comm_unique_id()."""
        buf = (ctypes.c_uint8 * _lib.NCCL_ID_BYTES).from_buffer_copy(unique_id) if unique_id is not None else None
        check(self.lib.gb2_comm_init(self.h, buf, int(rank), int(world)), "gb2_comm_init", self.h)
        self.rank, self.world = int(rank), int(world)

    def allreduce_hist(self, hist):
        """In-place sum over the ranks of an int64 device tensor of score-histogram counters (stream-ordered)."""
        assert hist.is_cuda and hist.dtype == torch.int64 and hist.is_contiguous()
        self.enter()
        hist.record_stream(self.stream)
        check(self.lib.gb2_allreduce_hist(self.h, _ptr(hist), hist.numel()), "gb2_allreduce_hist", self.h)
        return hist

    def allreduce_max(self, values):
        """Element-wise max over the ranks of a few host floats (device-timed durations) -> list of floats."""
        t = torch.tensor(list(values), dtype=torch.float64, device=self.device)
        self.enter()
        t.record_stream(self.stream)
        check(self.lib.gb2_allreduce_max_f64(self.h, _ptr(t), t.numel()), "gb2_allreduce_max_f64", self.h)
        self.sync()
        return t.cpu().tolist()

    def allgather(self, t):
        """Concatenation over the ranks of equally shaped contiguous device tensors -> tensor [world, *t.shape]."""
        assert t.is_cuda and t.is_contiguous()
        self.enter()
        t.record_stream(self.stream)
        with torch.cuda.stream(self.stream):
            out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=self.device)
        check(self.lib.gb2_allgather_bytes(self.h, _ptr(t), _ptr(out), t.numel() * t.element_size()), "gb2_allgather_bytes", self.h)
        return out

    # -- K1 -----------------------------------------------------------------------------------
    def encode(self, ascii_rows, width=None):
        """uint8 device tensor [n, stride] (ASCII k-mers) -> (packed int64[n] (uint64 bits; int64[n, 2] when the width is
        above 32), nmask int32[ceil(n/32)], counts int64[2] = rows masked, rows with a non-ACGTN symbol)."""
        assert ascii_rows.is_cuda and ascii_rows.dtype == torch.uint8 and ascii_rows.dim() == 2
        assert ascii_rows.stride(1) == 1
        n, stride = ascii_rows.shape[0], ascii_rows.stride(0) if ascii_rows.shape[0] > 1 else ascii_rows.shape[1]
        w = ascii_rows.shape[1] if width is None else int(width)
        self.enter()
        ascii_rows.record_stream(self.stream)
        packed = packed_rows(self, n, w)[:n]
        nmask = self.zeros((n + 31) // 32, torch.int32)
        counts = self.zeros(2, torch.int64)
        check(self.lib.gb2_encode_kmers(self.h, _ptr(ascii_rows), n, w, stride, _ptr(packed), _ptr(nmask), _ptr(counts)),
              "gb2_encode_kmers", self.h)
        self.leave()
        return packed, nmask, counts

    # -- line index shared by the text readers (K1b, K9) -------------------------------------------
    def index_lines(self, d_text, skip_minus=False, min_line_bytes=12):
        """Byte offsets of the non-blank lines of a device text buffer -> (line_off int64[n] device tensor, n)."""
        n_bytes = d_text.shape[0]
        cap = n_bytes // max(1, int(min_line_bytes)) + 2
        n_rows_d = self.zeros(1, torch.int64)
        while True:
            line_off = self.empty(cap, torch.int64)
            check(self.lib.gb2_tsv_index_lines(self.h, _ptr(d_text), n_bytes, int(bool(skip_minus)), _ptr(line_off), cap,
                                               _ptr(n_rows_d)), "gb2_tsv_index_lines", self.h)
            self.sync()
            n = int(n_rows_d.item())
            if n <= cap:
                break
            cap = n
        return line_off[:n], n

    # -- K1b: k-mer TSV on the device ------------------------------------------------------------
    def parse_kmer_tsv(self, text, width, skip_minus=False):
        """uint8 host tensor (pinned for speed) or device tensor with the bytes of `vg find -K w -E` TSV files ->
        DeviceRows (packed k-mers, N mask and the numeric side arrays on the device)."""
        assert text.dtype == torch.uint8 and text.dim() == 1
        n_bytes = text.shape[0]
        self.enter()
        with torch.cuda.stream(self.stream):
            d_text = text if text.is_cuda else text.to(self.device, non_blocking=True)
        # a well-formed line is at least 20 bytes ("a c a:1+ a:2+ 1 ref"); size the offset array for 12 and grow if needed
        line_off, n = self.index_lines(d_text, skip_minus, 12)
        rows = DeviceRows(self, d_text, line_off, n, width)
        if n:
            check(self.lib.gb2_tsv_parse_rows(self.h, _ptr(d_text), n_bytes, _ptr(rows.line_off), n, int(width), _ptr(rows.packed),
                                              _ptr(rows.nmask), _ptr(rows.start), _ptr(rows.stop), _ptr(rows.strand),
                                              _ptr(rows.freq), _ptr(rows.ref), _ptr(rows.name_len), _ptr(rows.seq_off),
                                              _ptr(rows.counts)), "gb2_tsv_parse_rows", self.h)
        self.leave()
        return rows

    # -- K3 -----------------------------------------------------------------------------------
    def pval_dp_batched(self, score_matrices, backgrounds):
        """list of int[4,w] (rows A,C,G,T) + list of [A,C,G,T] backgrounds -> list of float64[1000*w+1]."""
        m = len(score_matrices)
        if m == 0:
            return []
        widths = np.array([np.asarray(s).shape[1] for s in score_matrices], dtype=np.int32)
        sm = np.concatenate([np.ascontiguousarray(s, dtype=np.int64).reshape(-1) for s in score_matrices])
        bgs = np.ascontiguousarray(np.asarray(backgrounds, dtype=np.float64).reshape(m, 4))
        lens = _lib.RANGE * widths.astype(np.int64) + 1
        out = np.empty(int(lens.sum()), dtype=np.float64)
        check(self.lib.gb2_pval_dp_batched(self.h, m, _np_ptr(widths), _np_ptr(sm), _np_ptr(bgs), _np_ptr(out)),
              "gb2_pval_dp_batched", self.h)
        offs = np.concatenate([[0], np.cumsum(lens)])
        return [out[offs[i]:offs[i + 1]] for i in range(m)]  # views of one buffer: no second pass over ~80 MB for a collection

    def motif(self, score_matrix, pval_mat, min_val, scale, offset):
        return DeviceMotif(self, score_matrix, pval_mat, min_val, scale, offset)

    # -- haplotype tally -------------------------------------------------------------------------
    def tally_haplotypes(self, pos, packed, ref_packed=None, pos_base=0):
        """Per-haplotype windows (pos int64[n], packed int64[n], device; sorted in place) -> deduplicated rows:
        (pos, packed, freq, isref) device tensors trimmed to the number of distinct rows."""
        n = pos.shape[0]
        self.enter()
        u_pos = self.empty(n, torch.int64)
        u_packed = self.empty(n, torch.int64)
        u_freq = self.empty(n, torch.int32)
        u_isref = self.empty(n, torch.uint8)
        n_unique = self.zeros(1, torch.int64)
        n_ref = 0 if ref_packed is None else ref_packed.shape[0]
        check(self.lib.gb2_tally_haplotypes(self.h, _ptr(pos), _ptr(packed), n, _ptr(ref_packed), pos_base, n_ref,
                                            _ptr(u_pos), _ptr(u_packed), _ptr(u_freq), _ptr(u_isref), _ptr(n_unique)),
              "gb2_tally_haplotypes", self.h)
        self.sync()
        k = int(n_unique.item())
        self.leave()
        return u_pos[:k], u_packed[:k], u_freq[:k], u_isref[:k]


class DeviceRows:
    """Device-resident rows of a k-mer TSV: what score_seqs keeps per line (score_sequences.py:285-293), as arrays."""

    def __init__(self, ctx, d_text, line_off, n, width):
        self.ctx, self.d_text, self.line_off, self.n, self.width = ctx, d_text, line_off, n, width
        m = max(n, 1)
        self.packed = packed_rows(ctx, m, width)
        self.nmask = ctx.zeros((m + 31) // 32, torch.int32)
        self.start = ctx.empty(m, torch.int64)
        self.stop = ctx.empty(m, torch.int64)
        self.strand = ctx.empty(m, torch.uint8)
        self.freq = ctx.empty(m, torch.int64)
        self.ref = ctx.empty(m, torch.uint8)
        self.name_len = ctx.empty(m, torch.int32)
        self.seq_off = ctx.empty(m, torch.int32)
        self.counts = ctx.zeros(4, torch.int64)  # masked rows, bad-symbol rows, malformed lines

    def stats(self):
        self.ctx.sync()
        c = self.counts.cpu().numpy()
        return dict(n_rows=int(c[0]), bad_rows=int(c[1]), malformed=int(c[2]))

    def gather(self, rows):
        """Side arrays of the selected rows (device int64 indices) as numpy arrays."""
        with torch.cuda.stream(self.ctx.stream):
            out = {k: getattr(self, k)[rows].cpu().numpy() for k in
                   ("start", "stop", "strand", "freq", "ref", "name_len", "seq_off", "line_off")}
        return out


class SeqBatch:
    """A batch of sequences on the device in the 2-bit layout of include/grafimo_b200.h ("K2 over sequences"):
    seq2 int64[words] (32 bases per word), nbits int32[words] or None (1 bit per base: not A/C/G/T), and the host-side
    layout arrays lens / word_off (int64[n_seqs])."""

    def __init__(self, ctx, lens, seq2=None, nbits=None, word_off=None):
        self.ctx = ctx
        self.lens = np.ascontiguousarray(lens, dtype=np.int64)
        self.n_seqs = int(self.lens.shape[0])
        words = (self.lens + 31) // 32
        if word_off is None:
            word_off = np.concatenate([[0], np.cumsum(words)[:-1]]) if self.n_seqs else np.zeros(0)
        self.word_off = np.ascontiguousarray(word_off, dtype=np.int64)
        self.n_words = int((self.word_off + words).max()) if self.n_seqs else 0
        self.seq2 = seq2 if seq2 is not None else ctx.empty(max(self.n_words, 1), torch.int64)
        self.nbits = nbits

    def n_windows(self, w):
        return int(np.maximum(self.lens - int(w) + 1, 0).sum())

    def window_base(self, w):
        """int64[n_seqs + 1]: row index of the first window of every sequence (exclusive prefix of the window counts)."""
        return np.concatenate([[0], np.cumsum(np.maximum(self.lens - int(w) + 1, 0))]).astype(np.int64)

    @classmethod
    def from_ascii(cls, ctx, text, text_off, lens, want_nbits=True):
        """text: uint8 device tensor; sequence s = text[text_off[s] : text_off[s] + lens[s]] (ASCII, any case).
        -> (SeqBatch, counts int64[2] device tensor = bases not ACGT, bases neither ACGT nor N)."""
        assert text.is_cuda and text.dtype == torch.uint8 and text.dim() == 1 and text.is_contiguous()
        b = cls(ctx, lens)
        toff = np.ascontiguousarray(text_off, dtype=np.int64)
        ctx.enter()
        text.record_stream(ctx.stream)
        b.nbits = ctx.empty(max(b.n_words, 1), torch.int32) if want_nbits else None
        counts = ctx.zeros(2, torch.int64)
        check(ctx.lib.gb2_encode_sequences(ctx.h, _ptr(text), text.shape[0], b.n_seqs, _np_ptr(toff), _np_ptr(b.lens),
                                           _np_ptr(b.word_off), _ptr(b.seq2), _ptr(b.nbits), _ptr(counts)),
              "gb2_encode_sequences", ctx.h)
        ctx.leave()
        return b, counts


def pack_sequences_2bit(seqs):
    """Host-side (numpy) packer of a list of ACGTN strings into the 2-bit layout: (words uint64, nbits uint32, word_off
    int64, lens int64).  For callers / tests that hold sequences on the host; the device encoder is gb2_encode_sequences."""
    lens = np.array([len(x) for x in seqs], dtype=np.int64)
    nwords = (lens + 31) // 32
    word_off = np.concatenate([[0], np.cumsum(nwords)[:-1]]).astype(np.int64) if len(seqs) else np.zeros(0, np.int64)
    total = int(nwords.sum())
    words = np.zeros(max(total, 1), dtype=np.uint64)
    nbits = np.zeros(max(total, 1), dtype=np.uint32)
    code = np.full(256, 4, dtype=np.uint8)
    for k, ch in enumerate(b"ACGT"):
        code[ch] = k
        code[ch + 32] = k
    for x, off, n in zip(seqs, word_off, lens):
        if n == 0:
            continue
        c = code[np.frombuffer(x.encode("ascii"), dtype=np.uint8)]
        pad = (-n) % 32
        bad = np.concatenate([c > 3, np.zeros(pad, dtype=bool)]).reshape(-1, 32)
        c2 = np.concatenate([np.where(c > 3, 0, c), np.zeros(pad, dtype=np.uint8)]).astype(np.uint64).reshape(-1, 32)
        sh = (2 * np.arange(32, dtype=np.uint64))[None, :]
        words[off:off + c2.shape[0]] = np.bitwise_or.reduce(c2 << sh, axis=1)
        nbits[off:off + c2.shape[0]] = np.bitwise_or.reduce(bad.astype(np.uint32) << np.arange(32, dtype=np.uint32)[None, :], axis=1)
    return words, nbits, word_off, lens


def pack_sequences_host(seqs):
    """The library's host packer (gb2_pack_sequence_host: AVX-512 / AVX2 / scalar, no GPU involved) on a list of str / bytes /
    uint8 arrays -> (words uint64, nbits uint32, word_off int64, lens int64, counts uint64[2] = bases that are not ACGT / that
    are not N either): the layout of pack_sequences_2bit, which is its independent numpy check."""
    lib = _lib.load()
    bufs = [np.frombuffer(x.encode("ascii") if isinstance(x, str) else bytes(x) if not isinstance(x, np.ndarray) else x, dtype=np.uint8)
            for x in seqs]
    lens = np.array([len(b) for b in bufs], dtype=np.int64)
    nwords = (lens + 31) // 32
    word_off = np.concatenate([[0], np.cumsum(nwords)[:-1]]).astype(np.int64) if len(bufs) else np.zeros(0, np.int64)
    total = int(nwords.sum())
    words = np.zeros(max(total, 1), dtype=np.uint64)
    nbits = np.zeros(max(total, 1), dtype=np.uint32)
    counts = np.zeros(2, dtype=np.uint64)
    for b, off, n in zip(bufs, word_off, lens):
        if n:
            b = np.ascontiguousarray(b)
            rc = lib.gb2_pack_sequence_host(b.ctypes.data, int(n), words[off:].ctypes.data, nbits[off:].ctypes.data, counts.ctypes.data)
            if rc != 0:
                raise GrafimoB200Error(rc, "gb2_pack_sequence_host", "bad argument")
    return words, nbits, word_off, lens, counts


class DeviceMotif:
    """Device-resident motif: chunk LUTs + the score -> p-value table (K4)."""

    def __init__(self, ctx, score_matrix, pval_mat, min_val, scale, offset, _handle=None):
        self.ctx = ctx
        if _handle is None:
            _handle = DeviceMotif._create(ctx, [(score_matrix, pval_mat, min_val, scale, offset)])[0]
        self.h = _handle
        info = MotifInfo()
        check(ctx.lib.gb2_motif_get_info(self.h, ctypes.byref(info)), "gb2_motif_get_info")
        self.info = info
        self.width, self.lo, self.hi, self.span = info.width, info.lo, info.hi, info.span
        self._ptable = None

    @staticmethod
    def _create(ctx, items):
        """gb2_motif_create_batched: one allocation, one upload, two K4 launches and one synchronisation for all items."""
        sms, pms = [], []
        for sm, pm, _, _, _ in items:
            sm = np.ascontiguousarray(sm, dtype=np.int64)
            if sm.ndim != 2 or sm.shape[0] != 4:
                raise ValueError("score_matrix must be int[4, w] with rows A,C,G,T")
            pm = np.ascontiguousarray(pm, dtype=np.float64)
            if pm.shape != (_lib.RANGE * sm.shape[1] + 1,):
                raise ValueError(f"pval_mat must have {_lib.RANGE * sm.shape[1] + 1} entries for width {sm.shape[1]}")
            sms.append(sm)
            pms.append(pm)
        n = len(items)
        widths = np.array([sm.shape[1] for sm in sms], dtype=np.int32)
        sm_ptrs = (ctypes.c_void_p * n)(*[sm.ctypes.data for sm in sms])  # every motif keeps its own arrays: no concatenation
        pm_ptrs = (ctypes.c_void_p * n)(*[pm.ctypes.data for pm in pms])
        mins = np.array([int(it[2]) for it in items], dtype=np.int64)
        scales = np.array([int(it[3]) for it in items], dtype=np.int64)
        offs = np.array([float(it[4]) for it in items], dtype=np.float64)
        handles = (ctypes.c_void_p * n)()
        check(ctx.lib.gb2_motif_create_batched(ctx.h, n, _np_ptr(widths), sm_ptrs, pm_ptrs, _np_ptr(mins),
                                               _np_ptr(scales), _np_ptr(offs), handles), "gb2_motif_create_batched", ctx.h)
        return [ctypes.c_void_p(h) for h in handles]

    @classmethod
    def create_many(cls, ctx, items):
        """items: list of (score_matrix int[4, w], pval_mat, min_val, scale, offset) -> list of DeviceMotif, created together
        (a motif collection: BASELINE config 3)."""
        if not items:
            return []
        return [cls(ctx, None, None, None, None, None, _handle=h) for h in cls._create(ctx, items)]

    @property
    def ptable(self):
        """float64[span]: p-value of integer score lo+k (bit-exact to score_sequences.py:390-391)."""
        if self._ptable is None:
            out = np.empty(self.span, dtype=np.float64)
            check(self.ctx.lib.gb2_motif_get_ptable(self.ctx.h, self.h, _np_ptr(out)), "gb2_motif_get_ptable", self.ctx.h)
            self._ptable = out
        return self._ptable

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.gb2_motif_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scan:
    """State of one scan of one motif on one GPU: histogram, hit buffer, counters.

    score() may be called for any number of batches; histogram() exposes the per-score counts as an
    int64 tensor so that multi-GPU runs can all-reduce it (the only cross-GPU traffic); finalize()
    turns the hits into the numeric columns of the report."""

    def __init__(self, ctx, motif, strands=2, threshold=1e-4, want_q=True, hit_capacity=1 << 20, dense_rows=0):
        """dense_rows > 0: unselective scan (`-t 1`-like thresholds) of that many k-mers in all -- K2 writes dense scores
        instead of hit records and the report rows come from gb2_finalize_dense (batches must be scored in row order,
        each starting where the previous one ended; dense_rows * strands < 2^31)."""
        self.ctx, self.motif = ctx, motif
        self.strands, self.threshold, self.want_q = int(strands), float(threshold), bool(want_q)
        self.dense_rows = int(dense_rows)
        if self.dense_rows * self.strands >= (1 << 31):
            raise ValueError("dense scans are limited to 2^31-1 windows")
        self.capacity = self.dense_rows * self.strands if self.dense_rows else int(hit_capacity)
        self.hist = ctx.zeros(motif.span + 1, torch.int64) if want_q else None
        self.hits = None if self.dense_rows else ctx.empty(max(self.capacity, 1) * _HIT_BYTES, torch.uint8)
        self.dense = ctx.empty(self.dense_rows + 1, torch.int32) if self.dense_rows else None
        self.counters = ctx.zeros(4, torch.int64)  # [0] hits found, [1] kept, [2] total N
        self.rows_scored = 0
        self.row_limit = 0
        self.first_row = None

    def reset(self):
        with torch.cuda.stream(self.ctx.stream):
            if self.hist is not None:
                self.hist.zero_()
            self.counters.zero_()
        self.rows_scored = 0
        self.row_limit = 0
        self.first_row = None

    def score(self, packed, nmask=None, row_base=0, dense_out=None):
        """packed: int64[n] device tensor, or int64[n, 2] for a motif wider than 32 (two words per k-mer)."""
        n = packed.shape[0]
        lib, ctx = self.ctx.lib, self.ctx
        wide = self.motif.width > _lib.NARROW_WIDTH
        if (packed.dim() == 2) != wide or (wide and (packed.shape[1] != 2 or not packed.is_contiguous())):
            raise ValueError(f"k-mers of a width-{self.motif.width} motif must be packed as "
                             f"{'int64[n, 2]' if wide else 'int64[n]'}")
        ctx.enter()
        packed.record_stream(ctx.stream)  # the caller may free / reuse its tensors right after this asynchronous call
        if nmask is not None:
            nmask.record_stream(ctx.stream)
        if self.dense_rows:
            if self.first_row is None:
                self.first_row = int(row_base)
            if int(row_base) != self.first_row + self.rows_scored or self.rows_scored + n > self.dense_rows or dense_out is not None:
                raise ValueError("dense scan: batches must be consecutive rows within dense_rows")
            dense_out = self.dense[self.rows_scored:]
        check(lib.gb2_score(ctx.h, self.motif.h, _ptr(packed), _ptr(nmask), n, int(row_base), self.strands, self.threshold,
                            _ptr(self.hist), _ptr(self.hits), self.capacity if self.hits is not None else 0,
                            _ptr(self.counters), _ptr(dense_out)),
              "gb2_score", ctx.h)
        self.rows_scored += n
        self.row_limit = max(self.row_limit, int(row_base) + n)

    def score_sequences(self, batch, row_base=0):
        """K2 over a SeqBatch (2-bit sequences on the device): every window of every sequence, formed in registers.
        Window i of sequence s gets row index row_base + (windows of the sequences before s) + i."""
        lib, ctx = self.ctx.lib, self.ctx
        if self.motif.width > _lib.NARROW_WIDTH:
            raise ValueError("sequence scoring takes motifs of at most 32 bp; wider motifs use packed k-mers")
        n_win = batch.n_windows(self.motif.width)
        ctx.enter()
        dense_out = None
        if self.dense_rows:
            if self.first_row is None:
                self.first_row = int(row_base)
            if int(row_base) != self.first_row + self.rows_scored or self.rows_scored + n_win > self.dense_rows:
                raise ValueError("dense scan: batches must be consecutive rows within dense_rows")
            dense_out = self.dense[self.rows_scored:]
        nw = ctypes.c_uint64(0)
        check(lib.gb2_score_sequences(ctx.h, self.motif.h, _ptr(batch.seq2), _ptr(batch.nbits), batch.n_seqs, _np_ptr(batch.lens),
                                      _np_ptr(batch.word_off), None, int(row_base), self.strands, self.threshold, _ptr(self.hist),
                                      _ptr(self.hits), self.capacity if self.hits is not None else 0, _ptr(self.counters),
                                      _ptr(dense_out), ctypes.byref(nw)), "gb2_score_sequences", ctx.h)
        assert int(nw.value) == n_win
        self.rows_scored += n_win
        self.row_limit = max(self.row_limit, int(row_base) + n_win)
        return n_win

    def histogram(self):
        return self.hist

    def qvalues(self):
        """K5 on the (possibly all-reduced) histogram -> (qtab float64[span+1], rank int32[span+1]) device tensors."""
        ctx = self.ctx
        ctx.enter()
        nb = self.motif.span + 1
        self.qtab = ctx.empty(nb, torch.float64) if self.want_q else None
        self.rank = ctx.empty(nb, torch.int32)
        check(ctx.lib.gb2_qvalues_from_hist(ctx.h, self.motif.h, _ptr(self.hist), _ptr(self.qtab), _ptr(self.rank),
                                            ctypes.c_void_p(self.counters.data_ptr() + 16)), "gb2_qvalues_from_hist", ctx.h)
        return self.qtab, self.rank

    def n_hits(self):
        if self.dense_rows:  # every window is a candidate; the kept count comes from finalize
            return self.rows_scored * self.strands
        self.ctx.sync()
        return int(self.counters[0].item())

    def finalize_device(self, q_filter=False, index_only=False):
        """K6 with the columns left on the device (self.out: dict of device tensors); returns the rows kept.
        index_only (dense scans): only row, strand and integer score are written -- what the device report writer (K8)
        reads; score, p-value and q-value are functions of the integer score (13 instead of 37 bytes per row)."""
        ctx = self.ctx
        if not hasattr(self, "rank"):
            self.qvalues()
        n = self.n_hits()
        if n > self.capacity:
            raise GrafimoB200Error(_lib.GB2_ERR_CAPACITY, "Scan.finalize", f"{n} hits exceed the capacity {self.capacity}")
        cap = max(n, 1)
        if self.dense_rows and self.want_q and self.threshold < 0.5:
            # dense scan under a selective threshold: the number of report rows is the histogram mass of the kept bins --
            # size the output columns for that, not for every window (37 bytes each)
            with torch.cuda.stream(ctx.stream):
                pt = getattr(self.motif, "_ptab1_dev", None)
                if pt is None:
                    pt = torch.from_numpy(np.concatenate([np.asarray(self.motif.ptable, dtype=np.float64), [1.0]])).to(ctx.device)
                    self.motif._ptab1_dev = pt
                keep_bins = pt < self.threshold
                if q_filter:
                    keep_bins &= self.qtab < self.threshold
                cap = max(min(int(self.hist[keep_bins].sum().item()), n), 1)  # an all-reduced histogram counts the other ranks' rows too
        index_only = bool(index_only) and bool(self.dense_rows)
        if getattr(self, "_out_cap", 0) < cap or getattr(self, "_out_index_only", False) != index_only:
            f64 = (lambda: None) if index_only else (lambda: ctx.empty(cap, torch.float64))
            self._o = dict(row=ctx.empty(cap, torch.int64), strand=ctx.empty(cap, torch.uint8),
                           iscore=ctx.empty(cap, torch.int32), score=f64(), p=f64(), q=f64() if self.want_q else None)
            self._out_cap, self._out_index_only = cap, index_only
        o = self._o
        if self.dense_rows:
            check(ctx.lib.gb2_finalize_dense(ctx.h, self.motif.h, _ptr(self.dense), self.rows_scored, self.strands,
                                             int(self.first_row or 0), _ptr(self.qtab) if self.want_q else None,
                                             _ptr(self.rank), self.threshold, int(bool(q_filter)), self.threshold,
                                             _ptr(o["row"]), _ptr(o["strand"]), _ptr(o["iscore"]), _ptr(o["score"]),
                                             _ptr(o["p"]), _ptr(o["q"]), ctypes.c_void_p(self.counters.data_ptr() + 8)),
                  "gb2_finalize_dense", ctx.h)
        else:
            check(ctx.lib.gb2_finalize_hits(ctx.h, self.motif.h, _ptr(self.hits), n, int(self.row_limit),
                                            _ptr(self.qtab) if self.want_q else None, _ptr(self.rank), self.threshold,
                                            int(bool(q_filter)), self.threshold, _ptr(o["row"]), _ptr(o["strand"]),
                                            _ptr(o["iscore"]), _ptr(o["score"]), _ptr(o["p"]), _ptr(o["q"]),
                                            ctypes.c_void_p(self.counters.data_ptr() + 8)), "gb2_finalize_hits", ctx.h)
        ctx.sync()
        with torch.cuda.stream(ctx.stream):
            kept = int(self.counters[1].item())
        self.out = o
        return kept

    def finalize(self, q_filter=False):
        """K6 -> dict of numpy columns sorted by (p ascending, row, strand)."""
        ctx = self.ctx
        kept = self.finalize_device(q_filter)
        o_row, o_strand, o_iscore, o_score, o_p, o_q = (self.out[k] for k in ("row", "strand", "iscore", "score", "p", "q"))
        with torch.cuda.stream(ctx.stream):
            out = {
                "row": o_row[:kept].cpu().numpy(), "strand": o_strand[:kept].cpu().numpy(),
                "int_score": o_iscore[:kept].cpu().numpy(), "score": o_score[:kept].cpu().numpy(),
                "p-value": o_p[:kept].cpu().numpy(),
            }
            if self.want_q:
                out["q-value"] = o_q[:kept].cpu().numpy()
            out["total"] = int(self.counters[2].item()) if self.want_q else None
        return out


class ManyScan:
    """One scan of MANY motifs over the same k-mers on one GPU (a JASPAR-sized collection, BASELINE config 3): every
    motif has its own histogram / q-table, the hits of all motifs share one buffer (row_base = motif index << 40) and are
    finalized by ONE sort (gb2_finalize_hits_many) -- no sort and no host round trip per motif.
    score() and qvalues() only queue work; finalize_device() is the single synchronisation point."""

    ROW_BITS = 40

    def __init__(self, ctx, motifs, strands=2, threshold=1e-4, want_q=True, hit_capacity=1 << 22):
        self.ctx, self.motifs = ctx, list(motifs)
        self.strands, self.threshold, self.want_q = int(strands), float(threshold), bool(want_q)
        self.capacity = int(hit_capacity)
        sizes = [m.span + 1 for m in self.motifs]
        self.off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        total = int(self.off[-1])
        self.hist = ctx.zeros(total, torch.int64) if want_q else None
        self.qtab = ctx.empty(total, torch.float64) if want_q else None
        self.rank = ctx.empty(total, torch.int32)
        self.totals = ctx.zeros(len(self.motifs), torch.int64)
        self.hits = ctx.empty(max(self.capacity, 1) * _HIT_BYTES, torch.uint8)
        self.counters = ctx.zeros(4, torch.int64)  # [0] hits found (all motifs), [1] kept
        self.row_limit = 0

    def _at(self, tensor, m, itemsize):
        return ctypes.c_void_p(tensor.data_ptr() + int(self.off[m]) * itemsize) if tensor is not None else None

    def score(self, m, packed, nmask=None, row_offset=0):
        """Queues K2 of motif number m over `packed` (rows row_offset .. row_offset+n-1 of the k-mer set of its width)."""
        ctx, mo = self.ctx, self.motifs[m]
        n = packed.shape[0]
        if n + int(row_offset) >= (1 << self.ROW_BITS):
            raise ValueError("at most 2^40 - 1 rows per motif")
        ctx.enter()
        packed.record_stream(ctx.stream)
        if nmask is not None:
            nmask.record_stream(ctx.stream)
        check(ctx.lib.gb2_score(ctx.h, mo.h, _ptr(packed), _ptr(nmask), n, (int(m) << self.ROW_BITS) + int(row_offset), self.strands,
                                self.threshold, self._at(self.hist, m, 8), _ptr(self.hits), self.capacity, _ptr(self.counters), None),
              "gb2_score", ctx.h)
        self.row_limit = max(self.row_limit, n + int(row_offset))

    def qvalues(self):
        """Queues K5 of every motif (its own histogram -> its own q-table and p-rank table): ONE launch, one CTA per
        motif (gb2_qvalues_from_hist_many)."""
        ctx = self.ctx
        ctx.enter()
        k = len(self.motifs)
        handles = (ctypes.c_void_p * k)(*[mo.h.value for mo in self.motifs])
        check(ctx.lib.gb2_qvalues_from_hist_many(ctx.h, k, handles, _np_ptr(self.off), _ptr(self.hist), _ptr(self.qtab),
                                                 _ptr(self.rank), _ptr(self.totals)), "gb2_qvalues_from_hist_many", ctx.h)

    def finalize_device(self, q_filter=False):
        """One sort for all motifs -> self.out (device tensors: motif, row, strand, iscore, score, p, q), rows ordered by
        (motif, p ascending, row, strand); returns the number of rows kept."""
        ctx = self.ctx
        ctx.sync()
        n = int(self.counters[0].item())
        if n > self.capacity:
            raise GrafimoB200Error(_lib.GB2_ERR_CAPACITY, "ManyScan.finalize", f"{n} hits exceed the capacity {self.capacity}")
        cap = max(n, 1)
        o = dict(motif=ctx.empty(cap, torch.int32), row=ctx.empty(cap, torch.int64), strand=ctx.empty(cap, torch.uint8),
                 iscore=ctx.empty(cap, torch.int32), score=ctx.empty(cap, torch.float64), p=ctx.empty(cap, torch.float64),
                 q=ctx.empty(cap, torch.float64) if self.want_q else None)
        k = len(self.motifs)
        handles = (ctypes.c_void_p * k)(*[mo.h.value for mo in self.motifs])
        ranks = (ctypes.c_void_p * k)(*[self.rank.data_ptr() + 4 * int(self.off[m]) for m in range(k)])
        qtabs = (ctypes.c_void_p * k)(*[self.qtab.data_ptr() + 8 * int(self.off[m]) for m in range(k)]) if self.want_q else None
        check(ctx.lib.gb2_finalize_hits_many(ctx.h, k, handles, qtabs, ranks, _ptr(self.hits), n, max(self.row_limit, 1), self.threshold,
                                             int(bool(q_filter)), self.threshold, _ptr(o["motif"]), _ptr(o["row"]), _ptr(o["strand"]),
                                             _ptr(o["iscore"]), _ptr(o["score"]), _ptr(o["p"]), _ptr(o["q"]),
                                             ctypes.c_void_p(self.counters.data_ptr() + 8)), "gb2_finalize_hits_many", ctx.h)
        ctx.sync()
        self.out = o
        self.hits_found = n
        return int(self.counters[1].item())

    def hits_overflowed(self):
        """True when more hits were found than the buffer holds (the scan has to be repeated with hits_needed())."""
        self.ctx.sync()
        return int(self.counters[0].item()) > self.capacity

    def hits_needed(self):
        self.ctx.sync()
        return int(self.counters[0].item())

    def split_by_motif(self, kept):
        """-> int64[n_motifs + 1] offsets of every motif's (contiguous) rows in self.out."""
        with torch.cuda.stream(self.ctx.stream):
            cnt = torch.bincount(self.out["motif"][:kept].to(torch.int64), minlength=len(self.motifs)).cpu().numpy()
        return np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)


def scan_host(ctx, motif, ascii_rows, strands=1, threshold=1e-4, q_filter=False, want_q=True, hit_capacity=None, out=None):
    """gb2_scan_host: numpy/pinned-torch uint8 [n, w] host k-mers -> dict of numpy columns (+ stats).  out: a HostTable."""
    if isinstance(ascii_rows, torch.Tensor):
        assert not ascii_rows.is_cuda
        base, n, stride = ascii_rows.data_ptr(), ascii_rows.shape[0], ascii_rows.stride(0)
        w = ascii_rows.shape[1]
    else:
        a = np.ascontiguousarray(ascii_rows, dtype=np.uint8)
        base, n, stride, w = a.ctypes.data, a.shape[0], a.strides[0], a.shape[1]
    cap = int(hit_capacity if hit_capacity is not None else max(1024, min(n * strands, 1 << 26)))
    row, strand, isc, score, p, q = _host_outputs(cap, want_q, out)
    nh = ctypes.c_uint64(0)
    stats = np.zeros(4, np.uint64)
    rc = ctx.lib.gb2_scan_host(ctx.h, motif.h, ctypes.c_void_p(base), n, w, stride, int(strands), float(threshold),
                               int(bool(q_filter)), int(bool(want_q)), cap, _np_ptr(row), _np_ptr(strand), _np_ptr(isc),
                               _np_ptr(score), _np_ptr(p), _np_ptr(q) if want_q else None, ctypes.byref(nh), _np_ptr(stats))
    check(rc, "gb2_scan_host", ctx.h)
    return _host_result(int(nh.value), row, strand, isc, score, p, q, want_q, stats, ("windows", "n_rows", "bad_rows", "hits"))


class HostTable:
    """Reusable PINNED host buffers for the hit table of the host-buffer entry points (scan_host*, `out=`).  The table comes
    back with one asynchronous copy per column at the PCIe rate; fresh pageable numpy arrays (the default) cost a staged
    copy plus a page fault per 4 KB touched -- measured 4 ms of a 16 ms call for 489 k hits.  The arrays a scan returns are
    views of these buffers: they are overwritten by the next scan that is given the same HostTable."""

    def __init__(self, capacity, want_q=True):
        self.capacity = int(capacity)
        mk = lambda dt: torch.empty(self.capacity, dtype=dt, pin_memory=True)  # noqa: E731
        self._t = [mk(torch.int64), mk(torch.uint8), mk(torch.int32), mk(torch.float64), mk(torch.float64),
                   mk(torch.float64) if want_q else None]
        self.want_q = bool(want_q)

    def arrays(self):
        row, strand, isc, score, p, q = [None if t is None else t.numpy() for t in self._t]
        return row.view(np.uint64), strand, isc, score, p, q


def _host_outputs(cap, want_q, out=None):
    if out is not None:
        if out.capacity < cap or (want_q and not out.want_q):
            raise ValueError(f"HostTable of {out.capacity} rows given to a scan with hit capacity {cap}")
        return out.arrays()
    row = np.empty(cap, np.uint64); strand = np.empty(cap, np.uint8); isc = np.empty(cap, np.int32)
    score = np.empty(cap, np.float64); p = np.empty(cap, np.float64); q = np.empty(cap, np.float64) if want_q else None
    return row, strand, isc, score, p, q


def _host_result(k, row, strand, isc, score, p, q, want_q, stats, names):
    out = {"row": row[:k], "strand": strand[:k], "int_score": isc[:k], "score": score[:k], "p-value": p[:k]}
    if want_q:
        out["q-value"] = q[:k]
    out["stats"] = {n: int(v) for n, v in zip(names, stats)}
    return out


def _host_base(a):
    if isinstance(a, torch.Tensor):
        assert not a.is_cuda and a.is_contiguous()
        return a.data_ptr()
    return a.ctypes.data


def scan_host_packed(ctx, motif, packed, nmask=None, strands=1, threshold=1e-4, q_filter=False, want_q=True, hit_capacity=None, out=None):
    """gb2_scan_host_packed: 2-bit packed k-mers in host memory (numpy uint64/int64 or pinned torch int64; [n] or [n, 2]
    for a motif wider than 32) -> dict of numpy columns (+ stats)."""
    n = int(packed.shape[0])
    cap = int(hit_capacity if hit_capacity is not None else max(1024, min(n * strands, 1 << 26)))
    row, strand, isc, score, p, q = _host_outputs(cap, want_q, out)
    nh = ctypes.c_uint64(0)
    stats = np.zeros(4, np.uint64)
    rc = ctx.lib.gb2_scan_host_packed(ctx.h, motif.h, ctypes.c_void_p(_host_base(packed)),
                                      ctypes.c_void_p(_host_base(nmask)) if nmask is not None else None, n, int(strands),
                                      float(threshold), int(bool(q_filter)), int(bool(want_q)), cap, _np_ptr(row), _np_ptr(strand),
                                      _np_ptr(isc), _np_ptr(score), _np_ptr(p), _np_ptr(q) if want_q else None, ctypes.byref(nh),
                                      _np_ptr(stats))
    check(rc, "gb2_scan_host_packed", ctx.h)
    return _host_result(int(nh.value), row, strand, isc, score, p, q, want_q, stats, ("windows", "n_rows", "bad_rows", "hits"))


def scan_host_sequences(ctx, motif, data, offsets, lens, fmt="ascii", nbits=None, strands=1, threshold=1e-4, q_filter=False,
                        want_q=True, hit_capacity=None, out=None):
    """gb2_scan_host_sequences: whole sequences in host memory -> hit table.  fmt "ascii": data = uint8 bytes, offsets =
    byte offset of every sequence; fmt "2bit": data = uint64/int64 words (32 bases each), offsets = word offsets, nbits =
    optional uint32 per word.  Rows are window indices (sequence-major, see include/grafimo_b200.h)."""
    offs = np.ascontiguousarray(offsets, dtype=np.int64)
    ln = np.ascontiguousarray(lens, dtype=np.int64)
    n_win = int(np.maximum(ln - motif.width + 1, 0).sum())
    cap = int(hit_capacity if hit_capacity is not None else max(1024, min(n_win * strands, 1 << 26)))
    row, strand, isc, score, p, q = _host_outputs(cap, want_q, out)
    nh = ctypes.c_uint64(0)
    stats = np.zeros(4, np.uint64)
    rc = ctx.lib.gb2_scan_host_sequences(ctx.h, motif.h, 0 if fmt == "ascii" else 1, ctypes.c_void_p(_host_base(data)),
                                         ctypes.c_void_p(_host_base(nbits)) if nbits is not None else None, offs.shape[0],
                                         _np_ptr(offs), _np_ptr(ln), int(strands), float(threshold), int(bool(q_filter)),
                                         int(bool(want_q)), cap, _np_ptr(row), _np_ptr(strand), _np_ptr(isc), _np_ptr(score),
                                         _np_ptr(p), _np_ptr(q) if want_q else None, ctypes.byref(nh), _np_ptr(stats))
    check(rc, "gb2_scan_host_sequences", ctx.h)
    return _host_result(int(nh.value), row, strand, isc, score, p, q, want_q, stats, ("windows", "n_bases", "bad_bases", "hits"))


def bh_from_pvalues(ctx, pvalues):
    """gb2_bh_pvalues: float64[n] p-values -> float64[n] Benjamini-Hochberg q-values (input order)."""
    p = np.ascontiguousarray(pvalues, dtype=np.float64)
    q = np.empty_like(p)
    check(ctx.lib.gb2_bh_pvalues(ctx.h, _np_ptr(p), p.shape[0], _np_ptr(q)), "gb2_bh_pvalues", ctx.h)
    return q
