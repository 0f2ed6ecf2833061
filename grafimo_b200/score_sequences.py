"""B1 seam: `compute_results` -- scoring of the k-mers extracted from the variation graph.

Same signature, flags, side effects, error behaviour and returned table as the reference's
`grafimo.score_sequences.compute_results` (src/grafimo/score_sequences.py:44-211), with the per-row Python /
numba loop (`score_seqs` :216-326, `compute_score_seq` :331-396), the statsmodels BH step (:401-428) and the
filter + sort of `ResultTmp.to_df` (src/grafimo/resultsTmp.py:241-314) replaced by calls into the CUDA
library: the TSV bytes are indexed and parsed on the device (K1b), scored (K2), the q-values come from the score
histogram (K5) and the hits are filtered and sorted there (K6).  The host keeps what is text: reading the files,
the two string columns of the reported rows and the DataFrame.

There is no CPU fallback: without the CUDA library / a GPU this module raises.
"""
import glob
import os
import time
from typing import List, Optional

import numpy as np
import pandas as pd

from . import engine
from .motif import Motif, is_motif, score_matrix_acgt
from .utils import exception_handler

COLUMNS = ["motif_id", "motif_alt_id", "sequence_name", "start", "stop", "strand", "score", "p-value", "q-value",
           "matched_sequence", "haplotype_frequency", "reference"]

_ctx = None
_LOCAL_ONLY = [False]


class local_only:
    """Inside this block the scoring seams ignore the other ranks of a torchrun job: used when a motif COLLECTION is
    sharded over the GPUs by motif (every rank scans all the rows of its own motifs, so its q-values are already global
    and there is nothing to exchange) -- as opposed to one motif whose rows are sharded over the ranks."""

    def __enter__(self):
        self.prev = _LOCAL_ONLY[0]
        _LOCAL_ONLY[0] = True

    def __exit__(self, *a):
        _LOCAL_ONLY[0] = self.prev


def _dist_world():
    """(world, rank) of the scoring seams: the torchrun job, or (1, 0) on one process / inside local_only()."""
    import torch.distributed as tdist
    if _LOCAL_ONLY[0] or not (tdist.is_available() and tdist.is_initialized()):
        return 1, 0
    return tdist.get_world_size(), tdist.get_rank()


def _context():
    global _ctx
    if _ctx is None:
        _ctx = engine.Context()
    return _ctx


class KmerTable:
    """Columns of the `vg find -x XG -H GBWT -K w -E -p REGION` output (7 whitespace-separated fields:
    region, k-mer, chr:start(+|-), chr:stop(+|-), haplotype count, ref|non.ref, node path), parsed the way
    score_seqs does (src/grafimo/score_sequences.py:279-293)."""

    def __init__(self, seqname, seq, start, stop, strand, freq, ref):
        self.seqname, self.seq, self.start, self.stop = seqname, seq, start, stop
        self.strand, self.freq, self.ref = strand, freq, ref

    def __len__(self):
        return len(self.seq)

    @staticmethod
    def read(files: List[str], noreverse: bool) -> "KmerTable":
        frames = []
        for fn in files:
            if os.stat(fn).st_size == 0:
                continue
            df = pd.read_csv(fn, sep=r"\s+", header=None, usecols=[0, 1, 2, 3, 4, 5], dtype=str, engine="c",
                             na_filter=False, quoting=3)
            frames.append(df)
        if not frames:
            empty = np.array([], dtype=object)
            return KmerTable(empty, empty, np.array([], np.int64), np.array([], np.int64), empty, np.array([], np.int64), empty)
        df = pd.concat(frames, ignore_index=True) if len(frames) > 1 else frames[0]
        pos2 = df[2].to_numpy(dtype=object)
        pos3 = df[3].to_numpy(dtype=object)
        strand = np.array([s[-1] for s in pos2], dtype=object)  # strand = data[2][-1]
        start = np.array([int(s.split(":")[1][:-1]) for s in pos2], dtype=np.int64)
        stop = np.array([int(s.split(":")[1][:-1]) for s in pos3], dtype=np.int64)
        t = KmerTable(df[0].to_numpy(dtype=object), df[1].to_numpy(dtype=object), start, stop, strand,
                      df[4].to_numpy(dtype=object).astype(np.int64), df[5].to_numpy(dtype=object))
        if noreverse:  # '-' rows are dropped BEFORE scoring and counting (score_sequences.py:281-282)
            keep = strand != "-"
            t = KmerTable(t.seqname[keep], t.seq[keep], t.start[keep], t.stop[keep], t.strand[keep], t.freq[keep], t.ref[keep])
        return t

    def ascii_matrix(self, width: int, debug: bool):
        joined = "".join(self.seq.tolist()).encode("ascii")
        if len(joined) != len(self.seq) * width:
            exception_handler(ValueError, f"Every k-mer must have exactly {width} symbols (the motif width).\n", debug)
        import torch
        buf = torch.empty((len(self.seq), width), dtype=torch.uint8, pin_memory=True)
        buf.numpy()[...] = np.frombuffer(joined, dtype=np.uint8).reshape(len(self.seq), width)
        return buf


def print_scoring_msg(motif: Motif, noreverse: bool, debug: bool) -> None:
    if not is_motif(motif):
        exception_handler(TypeError, f"Expected Motif, got {type(motif).__name__}.\n", debug)
    if not isinstance(noreverse, bool):
        exception_handler(TypeError, f"Expected bool, got {type(noreverse).__name__}.\n", debug)
    msg = "Scoring hits for motif {}."
    print(msg.format("+" + motif.motif_id))
    if not noreverse:
        print(msg.format("-" + motif.motif_id), end="\n\n")


def device_motif(motif: Motif, ctx=None):
    """Uploads a Motif (cached on the object): chunk LUTs + the p-value table (K4)."""
    ctx = ctx or _context()
    cached = getattr(motif, "_gb2_device", None)
    if cached is not None and cached.ctx is ctx and cached.h:
        return cached
    dm = ctx.motif(score_matrix_acgt(motif), motif.pval_matrix, motif.min_val, motif.scale, float(motif.offset))
    try:
        motif._gb2_device = dm
    except AttributeError:  # an object that does not take new attributes: no cache
        pass
    return dm


def device_motifs(motifs, ctx=None):
    """device_motif for a whole collection: the motifs not uploaded yet are created TOGETHER (gb2_motif_create_batched:
    one allocation, one upload, two K4 launches, one synchronisation -- not one of each per motif)."""
    ctx = ctx or _context()
    todo = [m for m in motifs if not (getattr(m, "_gb2_device", None) is not None and m._gb2_device.ctx is ctx and m._gb2_device.h)]
    made = engine.DeviceMotif.create_many(ctx, [(score_matrix_acgt(m), m.pval_matrix, m.min_val, m.scale, float(m.offset)) for m in todo])
    for m, dm in zip(todo, made):
        try:
            m._gb2_device = dm
        except AttributeError:
            pass
    lookup = {id(m): dm for m, dm in zip(todo, made)}
    return [lookup.get(id(m)) or m._gb2_device for m in motifs]


# The reference calls compute_results once per motif on the same `width_<w>/*.tsv` files (src/grafimo/grafimo.py:177-179);
# the parsed, device-resident rows of the most recent file set are kept so that the next motif of that width starts at
# the scoring kernel instead of re-reading and re-parsing gigabytes of text.  One entry per path, bounded in size.
_PARSED = {}
_PARSED_MAX_BYTES = 16 << 30


def _files_key(files, width, no_reverse, ctx):
    st = [os.stat(f) for f in files]
    return (tuple((f, t.st_ino, t.st_mtime_ns, t.st_ctime_ns, t.st_size) for f, t in zip(files, st)), int(width), bool(no_reverse), id(ctx))


def _parsed_get(kind, key):
    e = _PARSED.get(kind)
    return e[1] if e is not None and e[0] == key else None


def _parsed_put(kind, key, value, n_bytes):
    if n_bytes <= _PARSED_MAX_BYTES:
        _PARSED[kind] = (key, value)
    else:
        _PARSED.pop(kind, None)


def clear_parsed_cache():
    """Drops the cached k-mer rows (host text and device arrays) of the last compute_results / scan_dir_device call."""
    _PARSED.clear()


_DENSE_FROM = 0.006  # p-value threshold from which the dense form is the faster one


def _dense_rows(threshold, n_kmers, strands):
    """Unselective thresholds (`-t 1` in docs/paper_results/run_analysis.sh) report a large share of the windows: K2 then
    writes dense scores (4 bytes per k-mer) and gb2_finalize_dense builds the rows -- no 16-byte hit records appended
    through an atomic counter, no 41-bit sort keys.  Measured on 2^26 k-mers, both strands (tools/bench_configs.py, section
    "midrange", profiles/r02_configs_midrange.json): the dense form costs 0.93 ms whatever the threshold, hit records
    0.48 / 0.68 / 1.20 / 2.39 / 4.25 / 6.39 ms at p < 0.001 / 0.003 / 0.01 / 0.03 / 0.1 / 0.25.  -> dense_rows of Scan."""
    return int(n_kmers) if (threshold >= _DENSE_FROM and 0 < n_kmers * strands < (1 << 31)) else 0


_CHUNK_BYTES = 256 << 20  # TSV text goes to the device in chunks of at most 256 MiB (cut at line boundaries): two pinned
# staging buffers of that size are all the host memory the route needs, whatever the input size


_STAGING = [None, None]  # two pinned buffers reused by every call: pinning costs ~0.45 s per GB, reading ~0.1 s per GB


def _staging_buffer(slot: int, nbytes: int, pin: bool):
    import torch
    buf = _STAGING[slot]
    if buf is None or buf.shape[0] < nbytes or buf.is_pinned() != pin:
        buf = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, pin_memory=pin)
        _STAGING[slot] = buf
    return buf[:nbytes]


def _text_chunks(files: List[str], chunk_bytes: int = _CHUNK_BYTES, segments: Optional[list] = None, reuse: bool = False):
    """Yields (pinned) uint8 tensors holding whole lines of the concatenated files; a newline follows every file (the
    blank line this may add is ignored by the line index), chunks are cut at line boundaries and hold at most
    chunk_bytes.  The byte ranges of a chunk are read with a small thread pool (os.preadv releases the GIL).  When
    `segments` is a list, one entry per yielded chunk is appended to it: [(file index, byte offset in the chunk where
    that file's lines begin), ...].
    reuse=True: the chunks are views of two alternating, process-wide pinned staging buffers -- the consumer must be done
    with chunk k (its host-to-device copy complete) before it asks for chunk k + 2."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    pin = torch.cuda.is_available()
    slot = [0]

    def alloc(nbytes):
        if not reuse:
            return torch.empty(nbytes, dtype=torch.uint8, pin_memory=pin)
        slot[0] ^= 1
        return _staging_buffer(slot[0], nbytes, pin)

    sizes = [os.stat(f).st_size for f in files]
    left = sum(sizes) + len(files)  # bytes still to deliver, file-terminating newlines included
    piece = 32 << 20
    cur_f, cur_pos = 0, 0
    carry, carry_file = b"", 0
    pool = ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1))

    def read(job):
        fn, lo, nbytes, dst, view = job
        fd = os.open(fn, os.O_RDONLY)
        try:
            done = 0
            while done < nbytes:
                got = os.preadv(fd, [memoryview(view)[dst + done:dst + nbytes]], lo + done)
                if got <= 0:
                    raise IOError(f"short read on {fn}")
                done += got
        finally:
            os.close(fd)

    try:
        while cur_f < len(files) or carry:
            cap = max(2, min(int(chunk_bytes), left + len(carry) + 1))
            if len(carry) + 2 > cap:
                raise ValueError(f"{files[carry_file]}: a line longer than the chunk size")
            buf = alloc(cap)
            view = buf.numpy()
            fill = len(carry)
            segs = []
            if fill:
                view[:fill] = np.frombuffer(carry, dtype=np.uint8)
            if cur_pos > 0:  # the file being read goes on in this chunk (a carried partial line is always its own)
                segs.append((cur_f, 0))
            jobs = []
            while cur_f < len(files):
                room = cap - fill - 1  # one spare byte for the file-terminating newline
                if room <= 0:
                    break
                if cur_pos == 0:
                    segs.append((cur_f, fill))
                n = min(sizes[cur_f] - cur_pos, room)
                for lo in range(0, n, piece):
                    jobs.append((files[cur_f], cur_pos + lo, min(piece, n - lo), fill + lo, view))
                fill += n
                cur_pos += n
                left -= n
                if cur_pos == sizes[cur_f]:
                    view[fill] = 10
                    fill += 1
                    left -= 1
                    cur_f += 1
                    cur_pos = 0
            if len(jobs) > 1:
                list(pool.map(read, jobs))
            else:
                for job in jobs:
                    read(job)
            if cur_f >= len(files):  # everything is in: the last chunk
                carry = b""
                if segments is not None:
                    segments.append(segs)
                yield buf[:fill]
                break
            lo = max(0, fill - (1 << 20))  # emit the whole lines, carry the partial one
            cut = bytes(memoryview(view)[lo:fill]).rfind(b"\n")
            if cut < 0:
                raise ValueError(f"{files[cur_f]}: a line longer than the chunk size / 1 MiB")
            cut += lo + 1
            carry, carry_file = bytes(memoryview(view)[cut:fill]), cur_f
            if segments is not None:
                segments.append([sg for sg in segs if sg[1] < cut])
            if cut:
                yield buf[:cut]
    finally:
        pool.shutdown(wait=True)


def _gather_rows(text, offs: np.ndarray, length: int) -> np.ndarray:
    """uint8 [n, length] = text[offs[i] + j] (indices clamped to the text).  `text` is the host copy of the bytes (numpy)
    or the device copy (torch tensor: the gather runs on the GPU, in slices, and only the selected bytes come back)."""
    n = len(offs)
    last = text.shape[0] - 1
    if isinstance(text, np.ndarray):
        return text[np.minimum(offs[:, None] + np.arange(length, dtype=np.int64)[None, :], last)]
    import torch
    out = np.empty((n, length), dtype=np.uint8)
    col = torch.arange(length, device=text.device, dtype=torch.int64)[None, :]
    step = max(1, (64 << 20) // max(1, length))
    for lo in range(0, n, step):
        o = torch.from_numpy(np.ascontiguousarray(offs[lo:lo + step], dtype=np.int64)).to(text.device)
        out[lo:lo + step] = text[(o[:, None] + col).clamp_(max=last)].cpu().numpy()
    return out


def _fixed_strings(text, offs: np.ndarray, length: int) -> np.ndarray:
    """Object array of the ASCII strings text[o:o+length]: one gather, one decode, one slicing pass."""
    n = len(offs)
    if n == 0:
        return np.array([], dtype=object)
    flat = np.ascontiguousarray(_gather_rows(text, offs, length)).tobytes().decode("ascii")
    out = np.empty(n, dtype=object)
    out[:] = [flat[i:i + length] for i in range(0, n * length, length)]
    return out


def _var_strings(text, offs: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """Object array of text[o:o+l].  Region names repeat in long runs (one name per vg file), so rows are compared
    with their predecessor in offset order and only the heads of the runs are decoded."""
    n = len(offs)
    if n == 0:
        return np.array([], dtype=object)
    order = np.argsort(offs, kind="stable")
    o, l = offs[order], lens[order]
    mx = int(l.max())
    col = np.arange(mx, dtype=np.int64)[None, :]
    chars = _gather_rows(text, o, mx)
    chars[col >= l[:, None]] = 0
    head = np.ones(n, dtype=bool)
    if n > 1:
        head[1:] = (l[1:] != l[:-1]) | (chars[1:] != chars[:-1]).any(axis=1)
    heads = np.nonzero(head)[0]
    names = np.empty(len(heads), dtype=object)
    names[:] = [bytes(chars[h, :l[h]]).decode("ascii") for h in heads.tolist()]
    out = np.empty(n, dtype=object)
    out[order] = names[np.cumsum(head) - 1]
    return out


def _report_order(pval, start, stop, strand, seq, minus=None, presorted=False):
    """Row order of the report: p-value ascending, ties -- which the reference leaves undefined -- by (start, stop,
    strand, matched_sequence).  Numeric keys first; the sequence strings are only compared inside groups that tie on
    all of them.  `minus` (bool array): the caller already knows which rows are on the '-' strand and every other row is '+'."""
    n = len(pval)
    if n < 2:
        return np.arange(n)
    if minus is None:
        minus = (strand == "-")
        other = ~minus & (strand != "+")
    else:  # the caller knows the strands (strand may be None)
        other = np.zeros(n, dtype=bool)
    scode = minus.astype(np.int8) * 2 + other.astype(np.int8) * 3  # '+' < '-' < anything else, like the characters
    # presorted: the rows already come in (p, start, stop, strand) order (sorted on the device); only the ties remain
    order = np.arange(n) if presorted else np.lexsort((scode, stop, start, pval))
    p, a, b, c = pval[order], start[order], stop[order], scode[order]
    same = (p[1:] == p[:-1]) & (a[1:] == a[:-1]) & (b[1:] == b[:-1]) & (c[1:] == c[:-1])
    if same.any():
        grp = np.concatenate([[0], np.cumsum(~same)])
        tied = np.zeros(n, dtype=bool)
        tied[1:] |= same
        tied[:-1] |= same
        idx = np.nonzero(tied)[0]
        sub = np.lexsort((seq[order][idx].astype(str), grp[idx]))
        order[idx] = order[idx][sub]
    return order


def _build_table(motif, no_qvalue, keep, seqname, start, stop, strand, score, pval, qval, seq, freq, ref, world, minus=None):
    """The 12-column table (resultsTmp.py:269-313) from per-hit arrays: filter, merge over ranks, order, DataFrame."""
    arrays = [seqname, start, stop, strand, score, pval, qval, seq, freq, ref]
    if not keep.all():
        arrays = [None if a is None else a[keep] for a in arrays]
        minus = None if minus is None else minus[keep]
    if world > 1:  # every rank returns the whole table
        import torch.distributed as tdist
        parts = [None] * world
        tdist.all_gather_object(parts, arrays)
        arrays = [None if parts[0][k] is None else np.concatenate([p[k] for p in parts]) for k in range(len(arrays))]
    seqname, start, stop, strand, score, pval, qval, seq, freq, ref = arrays
    order = _report_order(pval, start, stop, strand, seq, minus if world == 1 else None)
    n = len(order)
    cols = {
        "motif_id": np.full(n, motif.motif_id, dtype=object),
        "motif_alt_id": np.full(n, motif.motif_name, dtype=object),
        "sequence_name": seqname[order],
        "start": start[order],
        "stop": stop[order],
        "strand": strand[order],
        "score": score[order],
        "p-value": pval[order],
    }
    if not no_qvalue:
        cols["q-value"] = qval[order]
    cols["matched_sequence"] = seq[order]
    cols["haplotype_frequency"] = freq[order]
    cols["reference"] = ref[order]
    return pd.DataFrame(cols)


def _arrow_string_dtype():
    """The dtype pandas gives a column of Python strings when it is left to infer it: an Arrow-backed string dtype from
    pandas 3 on (then the string columns of a table can be built from byte buffers, without a Python object per cell), or
    None where it is still `object`."""
    try:
        dt = pd.Series(np.array(["a"], dtype=object)).dtype
        if dt == object:
            return None
        import pyarrow  # noqa: F401
        return dt
    except Exception:
        return None


def _build_table_rows(motif, no_qvalue, keep, names, region, start, stop, minus, score, pval, qval, asc, freq, isref_fixed, presorted=False):
    """_build_table for rows that come from the device (graph path): the same table, but where pandas backs string columns by
    Arrow (pandas >= 3) they are assembled from byte buffers -- the k-mer letters, one byte per strand, dictionary look-ups for
    region names and the ref flag -- instead of 4 Python string objects per row that pandas would then convert again
    (229 k rows: 0.32 s -> ~0.05 s; with eight ranks building the table at once the saving is larger)."""
    dt = _arrow_string_dtype()
    n_all = len(minus)
    width = asc.shape[1] if asc.ndim == 2 else 0
    if dt is None or n_all == 0:
        seq = np.ascontiguousarray(asc).view(f"S{width}").ravel().astype(f"U{width}").astype(object) if n_all else np.array([], dtype=object)
        seqname = np.array(names, dtype=object)[region] if n_all else np.array([], dtype=object)
        ref = np.where(isref_fixed, "ref", "non.ref").astype(object)
        strand = np.where(minus, "-", "+").astype(object)
        return _build_table(motif, no_qvalue, keep, seqname, start, stop, strand, score, pval, qval, seq, freq, ref, 1, minus=minus)
    import pyarrow as pa
    if not keep.all():
        region, start, stop, minus, score, pval, asc, freq, isref_fixed = (a[keep] for a in (region, start, stop, minus, score, pval, asc, freq, isref_fixed))
        qval = None if qval is None else qval[keep]
    seq_bytes = np.ascontiguousarray(asc).view(f"S{width}").ravel()
    order = _report_order(pval, start, stop, None, seq_bytes, minus, presorted=presorted)
    n = len(order)

    def fixed(rows_u8):  # [n, k] bytes -> Arrow strings of k characters
        k = rows_u8.shape[1]
        off = np.arange(0, (n + 1) * k, k, dtype=np.int64)
        return pa.Array.from_buffers(pa.large_string(), n, [None, pa.py_buffer(off), pa.py_buffer(np.ascontiguousarray(rows_u8))])

    def const(text):
        b = np.frombuffer(text.encode("utf-8"), dtype=np.uint8)
        return fixed(np.broadcast_to(b, (n, len(b))))

    def take(values, idx):
        return pa.array(list(values), type=pa.large_string()).take(pa.array(np.ascontiguousarray(idx, dtype=np.int64)))

    col = lambda arr: pd.array(arr, dtype=dt)  # noqa: E731
    cols = {
        "motif_id": col(const(str(motif.motif_id))),
        "motif_alt_id": col(const(str(motif.motif_name))),
        "sequence_name": col(take(names, region[order])),
        "start": start[order],
        "stop": stop[order],
        "strand": col(fixed(np.where(minus[order], np.uint8(45), np.uint8(43)).astype(np.uint8)[:, None])),
        "score": score[order],
        "p-value": pval[order],
    }
    if not no_qvalue:
        cols["q-value"] = qval[order]
    cols["matched_sequence"] = col(fixed(asc[order]))
    cols["haplotype_frequency"] = freq[order]
    cols["reference"] = col(take(["non.ref", "ref"], isref_fixed[order].astype(np.int64)))
    return pd.DataFrame(cols)


def compute_results(motif: Motif, sequence_loc: str, debug: bool, args_obj=None,
                    testmode: Optional[bool] = False) -> pd.DataFrame:
    """Scores every k-mer row under `<sequence_loc>/width_<w>/*.tsv` and returns the report table
    (columns and semantics of src/grafimo/resultsTmp.py:269-313).

    args_obj needs the attributes the reference reads (score_sequences.py:93-99): cores (ignored: the GPU
    does the work), threshold, noqvalue, qvalueT, noreverse, recomb, verbose.
    The TSV bytes go to the GPU as they are: lines are indexed and parsed there (K1b), k-mers packed, scored (K2),
    q-values derived from the score histogram (K5) and the hits filtered and sorted (K6); only the reported rows
    come back, and only their two string fields are sliced from the host copy of the text.
    Rows are ordered by p-value ascending; ties -- whose order the reference leaves undefined -- by
    (start, stop, strand, matched_sequence)."""
    if not is_motif(motif):
        exception_handler(TypeError, f"Expected Motif, got {type(motif).__name__}.\n", debug)
    if not isinstance(sequence_loc, str):
        exception_handler(TypeError, f"Expected str, got {type(sequence_loc).__name__}.\n", debug)
    if not os.path.isdir(sequence_loc):
        exception_handler(FileNotFoundError, f"Unable to locate {sequence_loc}.\n", debug)
    if not testmode:
        needed = ("threshold", "noqvalue", "qvalueT", "noreverse", "recomb", "verbose")
        if args_obj is None or not all(hasattr(args_obj, a) for a in needed):
            exception_handler(TypeError, f"Expected Findmotif, got {type(args_obj).__name__}.\n", debug)
        threshold, no_qvalue, qval_t = args_obj.threshold, args_obj.noqvalue, args_obj.qvalueT
        no_reverse, recomb, verbose = args_obj.noreverse, args_obj.recomb, args_obj.verbose
    else:  # the reference's pytest mode (score_sequences.py:100-107)
        threshold, recomb, no_qvalue, qval_t, no_reverse, verbose = float(1), True, False, False, False, False
    assert threshold > 0 and threshold <= 1
    if qval_t:
        assert not no_qvalue
    print_scoring_msg(motif, no_reverse, debug)
    if not motif.is_scaled:
        exception_handler(AssertionError, "The motif has not been scaled.\n", debug)
    width = motif.width
    files = sorted(glob.glob(os.path.join(sequence_loc, f"width_{width}", "*.tsv")))
    files = [f for f in files if os.stat(f).st_size > 0]
    # several GPUs (one process per GPU under torchrun): every rank takes every world-th file, like the reference
    # splits its files over `--cores` processes (score_sequences.py:120-147)
    import torch
    import torch.distributed as tdist
    from . import dist as gdist
    world, rank = _dist_world()
    files = files[rank::world]
    t0 = time.time()
    ctx = _context()
    dm = device_motif(motif, ctx)
    key = _files_key(files, width, no_reverse, ctx)
    cached = _parsed_get("table", key)
    if cached is not None:
        chunks, n_local = cached
    else:
        chunks = []  # (device text, DeviceRows, first local row)
        n_local = 0
        # the text goes through two reusable pinned staging buffers and stays on the device: the string columns of the
        # reported rows are gathered there, the host keeps no copy
        for text in _text_chunks(files, _CHUNK_BYTES, reuse=True) if files else ():
            rows = ctx.parse_kmer_tsv(text, width, skip_minus=no_reverse)
            chunks.append((rows.d_text, rows, n_local))
            n_local += rows.n
        _parsed_put("table", key, (chunks, n_local), sum(int(c[0].shape[0]) for c in chunks))
    if world > 1:
        counts = [None] * world
        tdist.all_gather_object(counts, n_local)
        rank_base, n = sum(counts[:rank]), sum(counts)
    else:
        rank_base, n = 0, n_local
    if n == 0:  # score_sequences.py:189-192
        errmsg = "No result retrieved. Unable to proceed.\n"
        errmsg += "\nAre you using the correct VGs and searching on the right chromosomes?\n"
        exception_handler(ValueError, errmsg, debug)
    stats = [c[1].stats() for c in chunks]
    bad = sum(st["malformed"] for st in stats)
    odd = sum(st["bad_rows"] for st in stats)
    if world > 1:  # every rank must fail together, or the others would wait in the exchange step until NCCL times out
        both = [None] * world
        tdist.all_gather_object(both, (bad, odd))
        bad, odd = sum(b for b, _ in both), sum(o for _, o in both)
    if bad:
        exception_handler(ValueError, f"{bad} k-mer rows are malformed (six fields and a k-mer of exactly {width} "
                          "symbols are required).\n", debug)
    if odd and rank == 0:
        import warnings
        warnings.warn(f"{odd} k-mer rows hold symbols other than A, C, G, T, N (IUPAC codes?): the reference leaves them "
                      "undefined; they are scored like N rows (p-value 1)")
    # every row is scored as given: `vg find -E` already emits the reverse-strand rows.  The hit buffer starts
    # small for selective thresholds and the (cheap) scoring pass is repeated in the rare case it overflows.
    cap = n_local if threshold >= 0.25 else min(n_local, max(1 << 20, n_local // 8))
    dense = _dense_rows(threshold, n_local, 1)
    while True:
        scan = engine.Scan(ctx, dm, strands=1, threshold=float(threshold), want_q=not no_qvalue, hit_capacity=cap, dense_rows=dense)
        for (_, rows, base), st in zip(chunks, stats):
            if rows.n:  # the N mask is only read by the kernel when some row of the chunk needs it
                scan.score(rows.packed, rows.nmask if st["n_rows"] else None, row_base=rank_base + base)
        found = scan.n_hits()
        if found <= cap:
            break
        cap = found
    if world > 1 and not no_qvalue:  # the one exchange step: global score histogram -> global q-values
        with torch.cuda.stream(ctx.stream):
            gdist.allreduce_histogram(scan.histogram(), ctx=ctx)
    kept = scan.finalize_device(q_filter=bool(qval_t))
    if rank == 0:
        if verbose:
            print("Sequences scored in %.2fs" % (time.time() - t0))
        if not no_qvalue:
            print("\nComputing q-values...\n")
        print(f"Scanned sequences:\t{n}")
        print(f"Scanned nucleotides:\t{n * width}")
    t1 = time.time()
    with torch.cuda.stream(ctx.stream):
        sel = scan.out["row"][:kept] - rank_base
        score = scan.out["score"][:kept].cpu().numpy()
        pval = scan.out["p"][:kept].cpu().numpy()
        qval = scan.out["q"][:kept].cpu().numpy() if not no_qvalue else None
        sel_h = sel.cpu().numpy()
    bases = np.array([c[2] for c in chunks] + [n_local], dtype=np.int64)
    which = np.searchsorted(bases, sel_h, side="right") - 1
    seqname = np.empty(kept, dtype=object); seq = np.empty(kept, dtype=object); strand = np.empty(kept, dtype=object)
    start = np.empty(kept, dtype=np.int64); stop = np.empty(kept, dtype=np.int64); freq = np.empty(kept, dtype=np.int64)
    ref = np.empty(kept, dtype=object)
    for k, (text, rows, base) in enumerate(chunks):
        m = np.nonzero(which == k)[0]
        if len(m) == 0:
            continue
        with torch.cuda.stream(ctx.stream):
            g = rows.gather(torch.from_numpy(sel_h[m] - base).to(ctx.device))
        off = g["line_off"].astype(np.int64)
        # leading blanks of a line belong to no field: the name starts at the first non-blank byte
        lead = np.zeros(len(m), dtype=np.int64)
        if len(off):
            first = _gather_rows(text, off, 1)[:, 0]
            while True:
                blank = (first == 32) | (first == 9)
                if not blank.any():
                    break
                lead[blank] += 1
                first = _gather_rows(text, off + lead, 1)[:, 0]
        seqname[m] = _var_strings(text, off + lead, g["name_len"].astype(np.int64))
        seq[m] = _fixed_strings(text, off + g["seq_off"].astype(np.int64), width)
        strand[m] = g["strand"].view("S1").astype("U1").astype(object)
        start[m], stop[m], freq[m] = g["start"], g["stop"], g["freq"]
        r = np.where(g["ref"] == 1, "ref", "non.ref").astype(object)
        other = np.nonzero(g["ref"] == 2)[0]
        for i in other:  # a sixth field that is neither "ref" nor "non.ref" is passed through verbatim
            r[i] = bytes(_gather_rows(text, off[i:i + 1], 4096)[0]).split(b"\n", 1)[0].split()[5].decode("ascii")
        ref[m] = r
    keep = np.ones(kept, dtype=bool) if recomb else freq > 0  # resultsTmp.py:309-310
    ref[(ref == "ref") & (np.abs(stop - start) != width)] = "non.ref"  # score_sequences.py:305-307
    df = _build_table(motif, no_qvalue, keep, seqname, start, stop, strand, score, pval, qval, seq, freq, ref, world)
    if verbose and rank == 0:
        print("\nResults summary built in %.2fs" % (time.time() - t1))
    return df


LAST_PHASES = {}  # GB2_PHASES=1: wall seconds of the phases of the last compute_results_rows call (stream-synchronised)


def _phase(ctx, name, t_prev):
    """Phase timer of compute_results_rows (only with GB2_PHASES=1: it synchronises the stream)."""
    if not os.environ.get("GB2_PHASES"):
        return t_prev
    ctx.sync()
    now = time.perf_counter()
    LAST_PHASES[name] = LAST_PHASES.get(name, 0.0) + now - t_prev
    return now


def compute_results_rows(motif: Motif, rows, debug: bool, args_obj=None, testmode: Optional[bool] = False) -> pd.DataFrame:
    """compute_results for k-mers that are already on the device: `rows` is a GraphRows (or a list of them, one per
    chromosome) from extract_regions.DeviceGraph.extract -- the forward walks of every region, with start/stop,
    haplotype frequency and ref flag as side arrays.  Returns the table compute_results would return for the TSVs
    `vg find -K w -E` writes for the same regions (GraphRows.to_vg_tsv): both strands are scored from the one
    packed k-mer (the '-' row of a walk is its reverse complement with start and stop swapped), every
    row of both strands counts in the q-values, and the same flags apply (score_sequences.py:93-107).
    Under torchrun (one process per GPU, e.g. chromosomes sharded over the ranks) every rank passes its own rows: the
    score histograms are all-reduced so the q-values are global, and every rank returns the whole table."""
    if not is_motif(motif):
        exception_handler(TypeError, f"Expected Motif, got {type(motif).__name__}.\n", debug)
    if not testmode:
        needed = ("threshold", "noqvalue", "qvalueT", "noreverse", "recomb", "verbose")
        if args_obj is None or not all(hasattr(args_obj, a) for a in needed):
            exception_handler(TypeError, f"Expected Findmotif, got {type(args_obj).__name__}.\n", debug)
        threshold, no_qvalue, qval_t = args_obj.threshold, args_obj.noqvalue, args_obj.qvalueT
        no_reverse, recomb, verbose = args_obj.noreverse, args_obj.recomb, args_obj.verbose
    else:
        threshold, recomb, no_qvalue, qval_t, no_reverse, verbose = float(1), True, False, False, False, False
    assert threshold > 0 and threshold <= 1
    if qval_t:
        assert not no_qvalue
    print_scoring_msg(motif, no_reverse, debug)
    if not motif.is_scaled:
        exception_handler(AssertionError, "The motif has not been scaled.\n", debug)
    import torch
    batches = [r for r in (rows if isinstance(rows, (list, tuple)) else [rows])]
    width = motif.width
    for b in batches:
        if b.width != width:
            exception_handler(ValueError, f"k-mers of width {b.width} given to a motif of width {width}.\n", debug)
    strands = 1 if no_reverse else 2
    n_kmers = sum(b.n for b in batches)
    import torch.distributed as tdist
    from . import dist as gdist
    world, rank = _dist_world()
    ctx = batches[0].ctx if batches else _context()
    LAST_PHASES.clear()
    tp = time.perf_counter()
    if world > 1:
        if getattr(ctx, "world", 1) == world:  # the context owns the communicator: 8 bytes per rank through the library
            counts = ctx.allgather(torch.tensor([n_kmers], dtype=torch.int64, device=ctx.device)).view(-1).cpu().tolist()
        else:
            counts = [None] * world
            tdist.all_gather_object(counts, n_kmers)
        rank_base, n_all = sum(counts[:rank]), sum(counts)
    else:
        rank_base, n_all = 0, n_kmers
    tp = _phase(ctx, "row_counts", tp)
    n = n_all * strands
    if n == 0:  # score_sequences.py:189-192
        errmsg = "No result retrieved. Unable to proceed.\n"
        errmsg += "\nAre you using the correct VGs and searching on the right chromosomes?\n"
        exception_handler(ValueError, errmsg, debug)
    t0 = time.time()
    dm = device_motif(motif, ctx)
    tp = _phase(ctx, "motif", tp)
    bases = (rank_base + np.concatenate([[0], np.cumsum([b.n for b in batches])])).astype(np.int64)
    n_local = n_kmers * strands
    cap = max(1, n_local if threshold >= 0.25 else min(n_local, max(1 << 20, n_local // 8)))
    dense = _dense_rows(threshold, n_kmers, strands)
    while True:
        scan = engine.Scan(ctx, dm, strands=strands, threshold=float(threshold), want_q=not no_qvalue, hit_capacity=cap, dense_rows=dense)
        for b, base in zip(batches, bases[:-1]):
            if b.n:
                scan.score(b.packed, b.nmask if b.n_masked() else None, row_base=int(base))
        found = scan.n_hits()
        if found <= cap:
            break
        cap = found
    tp = _phase(ctx, "score", tp)
    if world > 1 and not no_qvalue:  # the one exchange step: global score histogram -> global q-values
        with torch.cuda.stream(ctx.stream):
            gdist.allreduce_histogram(scan.histogram(), ctx=ctx)
    tp = _phase(ctx, "allreduce", tp)
    kept = scan.finalize_device(q_filter=bool(qval_t))
    tp = _phase(ctx, "finalize", tp)
    if rank == 0:
        if verbose:
            print("Sequences scored in %.2fs" % (time.time() - t0))
        if not no_qvalue:
            print("\nComputing q-values...\n")
        print(f"Scanned sequences:\t{n}")
        print(f"Scanned nucleotides:\t{n * width}")
    t1 = time.time()
    # ---- the fixed-width columns of this rank's hits, gathered on the device in hit order
    wide = width > 32
    with torch.cuda.stream(ctx.stream):
        sel = scan.out["row"][:kept]
        bases_dev = torch.from_numpy(bases[1:].copy()).to(ctx.device)
        which = torch.bucketize(sel, bases_dev, right=True)
        dev = {"packed": torch.empty((kept, 2) if wide else (kept,), dtype=torch.int64, device=ctx.device),
               "start": torch.empty(kept, dtype=torch.int64, device=ctx.device), "stop": torch.empty(kept, dtype=torch.int64, device=ctx.device),
               "freq": torch.empty(kept, dtype=torch.int64, device=ctx.device), "isref": torch.empty(kept, dtype=torch.uint8, device=ctx.device),
               "region": torch.empty(kept, dtype=torch.int64, device=ctx.device)}
        name_base = np.concatenate([[0], np.cumsum([len(b.regions) for b in batches])]).astype(np.int64)
        for k, b in enumerate(batches):
            if not b.n:
                continue
            m = torch.nonzero(which == k).view(-1)
            if m.numel() == 0:
                continue
            idx = sel[m] - int(bases[k])
            dev["packed"][m] = b.packed[idx]
            dev["start"][m], dev["stop"][m] = b.start[idx], b.stop[idx]
            dev["freq"][m] = b.freq[idx].to(torch.int64)
            dev["isref"][m] = b.isref[idx].to(torch.uint8)
            dev["region"][m] = b.region[idx].to(torch.int64) + int(name_base[k])
        dev["minus"] = scan.out["strand"][:kept].to(torch.uint8)
        dev["score"], dev["p"] = scan.out["score"][:kept], scan.out["p"][:kept]
        if not no_qvalue:
            dev["q"] = scan.out["q"][:kept]
    names = [b.region_name(r) for b in batches for r in range(len(b.regions))]
    tp = _phase(ctx, "hit_columns", tp)
    if world > 1:
        dev, names = _gather_hit_columns(ctx, dev, names, kept, world)
    tp = _phase(ctx, "gather_ranks", tp)
    with torch.cuda.stream(ctx.stream):
        # report order (p, start, stop, strand) on the device: four stable sorts of the (merged) hit columns, least significant
        # key first -- the host then only looks at rows that tie on all four (np.lexsort took 70-80 ms per rank for 229 k rows)
        n_rows = int(dev["p"].shape[0])
        if n_rows > 1:
            mi = dev["minus"].to(torch.bool)
            r_start = torch.where(mi, dev["stop"], dev["start"])
            r_stop = torch.where(mi, dev["start"], dev["stop"])
            idx = torch.arange(n_rows, device=ctx.device)
            for key in (dev["minus"].to(torch.int16), r_stop, r_start, dev["p"]):
                idx = idx[torch.sort(key[idx], stable=True).indices]
            dev = {k: v[idx] for k, v in dev.items()}
        dev["letters"] = _letters_device(dev.pop("packed"), dev["minus"], width) if n_rows else torch.zeros((0, width), dtype=torch.uint8, device=ctx.device)
        host = {k: v.cpu().numpy() for k, v in dev.items()}
    tp = _phase(ctx, "to_host", tp)
    minus = host["minus"].astype(bool)
    asc = host["letters"]
    # the '-' row of a walk starts where the walk stops
    start = np.where(minus, host["stop"], host["start"])
    stop = np.where(minus, host["start"], host["stop"])
    freq = host["freq"]
    isref = host["isref"].astype(bool) & (np.abs(stop - start) == width)  # score_sequences.py:305-307
    keep = np.ones(len(minus), dtype=bool) if recomb else freq > 0  # resultsTmp.py:309-310
    tp = _phase(ctx, "strings", tp)
    df = _build_table_rows(motif, no_qvalue, keep, names, host["region"], start, stop, minus, host["score"], host["p"], host.get("q"),
                           asc, freq, isref, presorted=True)
    tp = _phase(ctx, "dataframe", tp)
    if verbose and rank == 0:
        print("\nResults summary built in %.2fs" % (time.time() - t1))
    return df


def _gather_hit_columns(ctx, dev, names, kept, world):
    """Multi-rank merge of the report rows ON THE DEVICE: the fixed-width hit columns of every rank (packed k-mer, start,
    stop, frequency, ref flag, region id, strand, score, p, q) are padded to the largest per-rank count, laid back to back
    in one byte buffer and exchanged by ONE all-gather inside the library (gb2_allgather_bytes -> ncclAllGather; gloo /
    torch.distributed when the context owns no communicator); the strings are decoded once from the gathered columns.
    Replaces the pickled object arrays of all_gather_object (and the reference's Manager-dict funnel,
    score_sequences.py:115-118,171-188).  Region ids become global (rank-major); -> (columns, all region names)."""
    import torch
    import torch.distributed as tdist
    meta = [None] * world
    tdist.all_gather_object(meta, (int(kept), list(names)))  # a few bytes per rank: counts and region names
    counts = [m[0] for m in meta]
    name_off = np.concatenate([[0], np.cumsum([len(m[1]) for m in meta])])
    all_names = [nm for m in meta for nm in m[1]]
    rank = tdist.get_rank()
    cap = max(max(counts), 1)
    keys = list(dev.keys())
    with torch.cuda.stream(ctx.stream):
        dev["region"] = dev["region"] + int(name_off[rank])
        parts, layout = [], []
        for k in keys:
            v = dev[k].contiguous()
            row_bytes = v.element_size() * (v.shape[1] if v.dim() == 2 else 1)
            buf = torch.zeros(cap * row_bytes, dtype=torch.uint8, device=ctx.device)
            buf[:kept * row_bytes] = v.view(torch.uint8).view(-1)
            parts.append(buf)
            layout.append((k, v.dtype, row_bytes, v.dim() == 2))
        send = torch.cat(parts)
    if getattr(ctx, "world", 1) == world:
        gathered = ctx.allgather(send)
    else:
        gathered = torch.empty((world, send.numel()), dtype=torch.uint8, device=ctx.device)
        with torch.cuda.stream(ctx.stream):
            tdist.all_gather_into_tensor(gathered.view(-1), send)
    out = {}
    with torch.cuda.stream(ctx.stream):
        off = 0
        for k, dt, row_bytes, two in layout:
            pieces = [gathered[r, off:off + counts[r] * row_bytes] for r in range(world)]
            col = torch.cat(pieces).view(dt)
            out[k] = col.view(-1, 2) if two else col
            off += cap * row_bytes
    return out, all_names


def _revcomp_packed(packed, width):
    """Reverse complement of packed k-mers on the device: int64[n], or int64[n, 2] for widths above 32."""
    import torch
    wide = packed.dim() == 2
    words = [packed[:, 0], packed[:, 1]] if wide else [packed]
    out = [torch.zeros_like(w_) for w_ in words]
    for i in range(width):
        j = width - 1 - i
        out[j >> 5] |= (3 - ((words[i >> 5] >> (2 * (i & 31))) & 3)) << (2 * (j & 31))
    return torch.stack(out, dim=1) if wide else out[0]


def _letters_device(packed, minus, width):
    """Packed k-mers of the reported rows -> uint8 [n, width] ASCII on the device, the '-' rows as the reverse complement
    (what the report prints as matched_sequence): one small gather instead of a host pass over every hit."""
    import torch
    wide = packed.dim() == 2
    words = [packed[:, 0], packed[:, 1]] if wide else [packed]
    fwd = torch.stack([(words[int(i) >> 5] >> (2 * (int(i) & 31))) & 3 for i in range(width)], dim=1)  # [n, width] codes 0..3
    rc = 3 - fwd.flip(1)
    codes = torch.where(minus.to(torch.bool)[:, None], rc, fwd)
    return torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=packed.device)[codes]


def scan_rows_device(motif: Motif, rows, debug: bool, args_obj):
    """The scoring of compute_results_rows with the report left on the device: -> res_writer.DeviceReport (hit columns
    in HBM + per-bin score / p / q values) for write_results_device (K8).  Same flags and semantics; rows are ordered by
    (p-value, row, strand) -- deterministic; the reference leaves ties unordered.  Single process (one GPU)."""
    import torch
    from .res_writer import DeviceReport
    if not is_motif(motif):
        exception_handler(TypeError, f"Expected Motif, got {type(motif).__name__}.\n", debug)
    threshold, no_qvalue, qval_t = args_obj.threshold, args_obj.noqvalue, args_obj.qvalueT
    no_reverse, recomb = args_obj.noreverse, args_obj.recomb
    print_scoring_msg(motif, no_reverse, debug)
    if not motif.is_scaled:
        exception_handler(AssertionError, "The motif has not been scaled.\n", debug)
    batches = [r for r in (rows if isinstance(rows, (list, tuple)) else [rows])]
    width = motif.width
    strands = 1 if no_reverse else 2
    n_kmers = sum(b.n for b in batches)
    n = n_kmers * strands
    if n == 0:
        exception_handler(ValueError, "No result retrieved. Unable to proceed.\n", debug)
    ctx = batches[0].ctx
    dev = ctx.device
    dm = device_motif(motif, ctx)
    bases = np.concatenate([[0], np.cumsum([b.n for b in batches])]).astype(np.int64)
    cap = max(1, n if threshold >= 0.25 else min(n, max(1 << 20, n // 8)))
    dense = _dense_rows(threshold, n_kmers, strands)
    while True:
        scan = engine.Scan(ctx, dm, strands=strands, threshold=float(threshold), want_q=not no_qvalue, hit_capacity=cap, dense_rows=dense)
        for b, base in zip(batches, bases[:-1]):
            if b.n:
                scan.score(b.packed, b.nmask if b.n_masked() else None, row_base=int(base))
        found = scan.n_hits()
        if found <= cap:
            break
        cap = found
    kept = scan.finalize_device(q_filter=bool(qval_t), index_only=True)  # K8 prints score / p / q from per-bin tables
    if not no_qvalue:
        print("\nComputing q-values...\n")
    print(f"Scanned sequences:\t{n}")
    print(f"Scanned nucleotides:\t{n * width}")
    with torch.cuda.stream(ctx.stream):
        qtab = scan.qtab.cpu().numpy() if not no_qvalue else None
    return _device_report(ctx, motif, dm, batches, scan.out["row"][:kept], scan.out["strand"][:kept], scan.out["iscore"][:kept],
                          qtab, recomb, no_qvalue)


def _device_report(ctx, motif, dm, batches, row, strand_u8, iscore, qtab, recomb, no_qvalue):
    """res_writer.DeviceReport from the device-resident hit columns of one motif (rows index the concatenation of
    `batches`): side arrays gathered on the device, '-' hits reverse-complemented with start/stop swapped,
    the ref rewrite of score_sequences.py:305-307 and the recomb filter of resultsTmp.py:309-310."""
    import torch
    from .res_writer import DeviceReport
    dev, width = ctx.device, motif.width
    with torch.cuda.stream(ctx.stream):
        minus = strand_u8.to(torch.bool)
        bin_ = (iscore - int(dm.lo)).to(torch.int32)
        cat = lambda name, dt: torch.cat([getattr(b, name)[:b.n].to(dt) for b in batches]) if len(batches) > 1 else getattr(batches[0], name)[:batches[0].n].to(dt)  # noqa: E731
        packed, start, stop = cat("packed", torch.int64)[row], cat("start", torch.int64)[row], cat("stop", torch.int64)[row]
        freq, isref = cat("freq", torch.int64)[row], cat("isref", torch.bool)[row]
        name_base = np.concatenate([[0], np.cumsum([len(b.regions) for b in batches])])
        region = torch.cat([b.region[:b.n].to(torch.int32) + int(nb) for b, nb in zip(batches, name_base[:-1])])[row]
        # '-' hits: reverse complement of the k-mer, start/stop swapped
        rc = _revcomp_packed(packed, width)
        kmer = torch.where(minus[:, None] if packed.dim() == 2 else minus, rc, packed)
        s2 = torch.where(minus, stop, start)
        e2 = torch.where(minus, start, stop)
        ref = (isref & ((e2 - s2).abs() == width)).to(torch.uint8)  # score_sequences.py:305-307
        strand = torch.where(minus, torch.tensor(45, dtype=torch.uint8, device=dev), torch.tensor(43, dtype=torch.uint8, device=dev))
        keep = torch.ones_like(minus) if recomb else freq > 0  # resultsTmp.py:309-310
        cols = [c[keep].contiguous() for c in (kmer, strand, s2, e2, freq, ref, bin_, region)]
    ctx.sync()
    seqnames = [b.region_name(r) for b in batches for r in range(len(b.regions))]
    span = int(dm.span)
    score_by_bin = (np.arange(span, dtype=np.int64) + int(dm.lo)) / np.float64(motif.scale) + np.float64(width) * np.float64(motif.offset)
    return DeviceReport(ctx, motif, width, not no_qvalue, *cols, seqnames, score_by_bin, dm.ptable,
                        qtab[:span] if qtab is not None else None)


def scan_rows_device_many(motifs, rows_of_width, debug: bool, args_obj):
    """scan_rows_device for a motif COLLECTION over the same k-mers (BASELINE config 3; replaces the per-motif loop of
    src/grafimo/grafimo.py:177-183): the motifs are uploaded together (gb2_motif_create_batched), K2 of every motif is
    queued into one shared hit buffer, K5 of all motifs is one launch and ONE sort finalizes the hits of all of them
    (engine.ManyScan) -- no host round trip per motif.  `rows_of_width`: {width: [GraphRows, ...]}.  -> list of
    DeviceReport, one per motif, each identical to what scan_rows_device returns for that motif alone.  Unselective
    thresholds (>= 0.25: every window is a report row) take the per-motif dense route."""
    import torch
    threshold, no_qvalue, qval_t = args_obj.threshold, args_obj.noqvalue, args_obj.qvalueT
    no_reverse, recomb = args_obj.noreverse, args_obj.recomb
    if not motifs:
        return []
    if threshold >= 0.25 or len(motifs) == 1:
        return [scan_rows_device(m, rows_of_width[m.width], debug, args_obj) for m in motifs]
    for m in motifs:
        if not is_motif(m):
            exception_handler(TypeError, f"Expected Motif, got {type(m).__name__}.\n", debug)
        if not m.is_scaled:
            exception_handler(AssertionError, "The motif has not been scaled.\n", debug)
    strands = 1 if no_reverse else 2
    ctx = next(iter(rows_of_width.values()))[0].ctx
    dms = device_motifs(motifs, ctx)
    n_of_width = {w: sum(b.n for b in bs) for w, bs in rows_of_width.items()}
    windows = sum(n_of_width[m.width] * strands for m in motifs)
    if any(n_of_width[m.width] == 0 for m in motifs):
        exception_handler(ValueError, "No result retrieved. Unable to proceed.\n", debug)
    cap = max(1 << 20, min(windows, int(4.0 * threshold * windows) + (1 << 20)))
    while True:
        many = engine.ManyScan(ctx, dms, strands=strands, threshold=float(threshold), want_q=not no_qvalue, hit_capacity=cap)
        for k, m in enumerate(motifs):
            base = 0
            for b in rows_of_width[m.width]:
                if b.n:
                    many.score(k, b.packed, b.nmask if b.n_masked() else None, row_offset=base)
                base += b.n
        if not many.hits_overflowed():
            break
        cap = many.hits_needed()
    many.qvalues()
    kept = many.finalize_device(q_filter=bool(qval_t))
    off = many.split_by_motif(kept)
    with torch.cuda.stream(ctx.stream):
        qtab_all = many.qtab.cpu().numpy() if not no_qvalue else None
    reports = []
    for k, (m, dm) in enumerate(zip(motifs, dms)):
        print_scoring_msg(m, no_reverse, debug)
        if not no_qvalue:
            print("\nComputing q-values...\n")
        n = n_of_width[m.width] * strands
        print(f"Scanned sequences:\t{n}")
        print(f"Scanned nucleotides:\t{n * m.width}")
        lo, hi = int(off[k]), int(off[k + 1])
        o = many.out
        qtab = qtab_all[int(many.off[k]):int(many.off[k + 1])] if qtab_all is not None else None
        reports.append(_device_report(ctx, m, dm, rows_of_width[m.width], o["row"][lo:hi], o["strand"][lo:hi], o["iscore"][lo:hi],
                                      qtab, recomb, no_qvalue))
    return reports


_GENERAL = "general path"  # cached verdict of _parse_dir_for_report: the input needs compute_results


def _parse_dir_for_report(files, width, no_reverse, ctx, debug):
    """Parses the TSV files for scan_dir_device: -> (chunks [(DeviceRows with .name_id, first row, stats)], region names,
    number of rows), or _GENERAL when the input needs the DataFrame path."""
    import torch
    dev = ctx.device
    segments, chunks, names, n = [], [], [], 0
    name_of_file = {}
    for text in _text_chunks(files, _CHUNK_BYTES, segments, reuse=True) if files else ():
        rows = ctx.parse_kmer_tsv(text, width, skip_minus=no_reverse)
        st = rows.stats()
        if st["malformed"]:
            exception_handler(ValueError, f"{st['malformed']} k-mer rows are malformed (six fields and a k-mer of exactly {width} "
                              "symbols are required).\n", debug)
        if st["bad_rows"] or (rows.n and bool((rows.ref[:rows.n] == 2).any().item())):
            return _GENERAL
        host = text.numpy()
        segs = segments[-1]
        with torch.cuda.stream(ctx.stream):
            seg_start = torch.tensor([o for _, o in segs], dtype=torch.int64, device=dev)
            file_of_row = torch.bucketize(rows.line_off[:rows.n], seg_start, right=True) - 1  # index into segs
            # every line of a file must begin with the name its first line has (vg writes one file per region)
            if rows.n:
                first_row = torch.full((len(segs),), rows.n, dtype=torch.int64, device=dev)
                first_row.scatter_reduce_(0, file_of_row, torch.arange(rows.n, device=dev), reduce="amin")
                has = first_row < rows.n
                fr = first_row.clamp(max=max(rows.n - 1, 0))
                nl = rows.name_len[:rows.n].to(torch.int64)
                off = rows.line_off[:rows.n]
                ok = bool((nl == nl[fr][file_of_row]).all().item())
                first_byte = rows.d_text[off]
                ok = ok and not bool(((first_byte == 32) | (first_byte == 9)).any().item())  # leading blanks: general path
                if ok:  # k-mers are printed from their packed form, i.e. in upper case: lower-case input takes the general path
                    kcol = torch.arange(width, device=dev)[None, :]
                    kb = rows.d_text[(off + rows.seq_off[:rows.n].to(torch.int64))[:, None] + kcol]
                    ok = not bool((kb >= 97).any().item())
                if ok:
                    mx = int(nl.max().item())
                    col = torch.arange(mx, device=dev)[None, :]
                    a = rows.d_text[(off[:, None] + col).clamp(max=rows.d_text.shape[0] - 1)]
                    b = rows.d_text[(off[fr][file_of_row][:, None] + col).clamp(max=rows.d_text.shape[0] - 1)]
                    ok = bool(((a == b) | (col >= nl[:, None])).all().item())
                if not ok:
                    return _GENERAL
                fr_h, has_h = fr.cpu().numpy(), has.cpu().numpy()
                off_h, nl_h = off[fr].cpu().numpy(), nl[fr].cpu().numpy()
                local = np.full(len(segs), -1, dtype=np.int64)
                for k, (fi, _) in enumerate(segs):
                    if not has_h[k]:
                        continue
                    nm = bytes(host[off_h[k]:off_h[k] + nl_h[k]]).decode("ascii")
                    if fi in name_of_file and names[name_of_file[fi]] != nm:
                        return _GENERAL
                    if fi not in name_of_file:
                        name_of_file[fi] = len(names)
                        names.append(nm)
                    local[k] = name_of_file[fi]
                name_id = torch.from_numpy(local).to(dev)[file_of_row].to(torch.int32)
            else:
                name_id = torch.zeros(0, dtype=torch.int32, device=dev)
        rows.d_text = None
        rows.name_id = name_id
        chunks.append((rows, n, st))
        n += rows.n
    return chunks, names, n


def scan_dir_device(motif: Motif, sequence_loc: str, debug: bool, args_obj):
    """compute_results with the report left on the device: the `vg find` TSVs under `<sequence_loc>/width_<w>/` are
    parsed, scored and finalized on the GPU and the hit columns stay there -> res_writer.DeviceReport for
    write_results_device (K8), no DataFrame.  Returns None when the input needs the general path (compute_results):
    a file whose lines do not all carry the same region name (vg writes one file per region), lines with leading blanks,
    a reference column that is neither `ref` nor `non.ref`, lower-case or non-ACGTN k-mers, or more than one process.  Rows are ordered by (p-value, row, strand)."""
    import torch
    import torch.distributed as tdist
    from .res_writer import DeviceReport
    if _dist_world()[0] > 1:
        return None
    if not is_motif(motif):
        exception_handler(TypeError, f"Expected Motif, got {type(motif).__name__}.\n", debug)
    if not os.path.isdir(sequence_loc):
        exception_handler(FileNotFoundError, f"Unable to locate {sequence_loc}.\n", debug)
    threshold, no_qvalue, qval_t = args_obj.threshold, args_obj.noqvalue, args_obj.qvalueT
    no_reverse, recomb = args_obj.noreverse, args_obj.recomb
    if not motif.is_scaled:
        exception_handler(AssertionError, "The motif has not been scaled.\n", debug)
    width = motif.width
    files = sorted(glob.glob(os.path.join(sequence_loc, f"width_{width}", "*.tsv")))
    files = [f for f in files if os.stat(f).st_size > 0]
    ctx = _context()
    dev = ctx.device
    dm = device_motif(motif, ctx)
    key = _files_key(files, width, no_reverse, ctx)
    parsed = _parsed_get("report", key)
    if parsed is None:
        parsed = _parse_dir_for_report(files, width, no_reverse, ctx, debug)
        nbytes = 0 if parsed is _GENERAL else sum(r.packed.numel() * 8 + r.n * 40 for r, _, _ in parsed[0])
        _parsed_put("report", key, parsed, nbytes)
    if parsed is _GENERAL:
        return None
    chunks, names, n = parsed
    print_scoring_msg(motif, no_reverse, debug)
    if n == 0:
        errmsg = "No result retrieved. Unable to proceed.\n"
        errmsg += "\nAre you using the correct VGs and searching on the right chromosomes?\n"
        exception_handler(ValueError, errmsg, debug)
    cap = max(1, n if threshold >= 0.25 else min(n, max(1 << 20, n // 8)))
    dense = _dense_rows(threshold, n, 1)
    while True:
        scan = engine.Scan(ctx, dm, strands=1, threshold=float(threshold), want_q=not no_qvalue, hit_capacity=cap, dense_rows=dense)
        for rows, base, st in chunks:
            if rows.n:
                scan.score(rows.packed, rows.nmask if st["n_rows"] else None, row_base=base)
        found = scan.n_hits()
        if found <= cap:
            break
        cap = found
    kept = scan.finalize_device(q_filter=bool(qval_t), index_only=True)
    if not no_qvalue:
        print("\nComputing q-values...\n")
    print(f"Scanned sequences:\t{n}")
    print(f"Scanned nucleotides:\t{n * width}")
    with torch.cuda.stream(ctx.stream):
        row = scan.out["row"][:kept]
        bin_ = (scan.out["iscore"][:kept] - int(dm.lo)).to(torch.int32)
        cat = lambda name, dt: torch.cat([getattr(r, name)[:r.n].to(dt) for r, _, _ in chunks])  # noqa: E731
        kmer, start, stop = cat("packed", torch.int64)[row], cat("start", torch.int64)[row], cat("stop", torch.int64)[row]
        freq, refc, strand = cat("freq", torch.int64)[row], cat("ref", torch.uint8)[row], cat("strand", torch.uint8)[row]
        name_id = cat("name_id", torch.int32)[row]
        ref = ((refc == 1) & ((stop - start).abs() == width)).to(torch.uint8)  # score_sequences.py:305-307
        keep = torch.ones_like(ref, dtype=torch.bool) if recomb else freq > 0  # resultsTmp.py:309-310
        cols = [c[keep].contiguous() for c in (kmer, strand, start, stop, freq, ref, bin_, name_id)]
        qtab = scan.qtab.cpu().numpy() if not no_qvalue else None
    ctx.sync()
    span = int(dm.span)
    score_by_bin = (np.arange(span, dtype=np.int64) + int(dm.lo)) / np.float64(motif.scale) + np.float64(width) * np.float64(motif.offset)
    return DeviceReport(ctx, motif, width, not no_qvalue, *cols, names, score_by_bin, dm.ptable,
                        qtab[:span] if qtab is not None else None)


def compute_qvalues(pvalues: List[float], debug: bool) -> List[float]:
    """B3 seam (src/grafimo/score_sequences.py:401-428): Benjamini-Hochberg q-values of a list of p-values,
    same order in and out.  Inside compute_results the q-values come from the score histogram (K5); this
    stand-alone form bins the distinct p-values and runs the same kernel."""
    if not isinstance(pvalues, list):
        exception_handler(TypeError, f"Expected list, got {type(pvalues).__name__}.\n", debug)
    print("\nComputing q-values...\n")
    p = np.asarray(pvalues, dtype=np.float64)
    q = engine.bh_from_pvalues(_context(), p)
    assert len(q) == len(pvalues)
    return list(q)
