"""Multi-GPU plumbing: one process per GPU, rows sharded by region / chromosome, ONE exchange step.

The reference merges the per-process results and runs Benjamini-Hochberg on all p-values at once
(src/grafimo/score_sequences.py:171-198).  Here every rank scores its own shard; because the p-value is a
function of the integer score, the global multiset of p-values is the element-wise sum of the per-rank score
histograms, so the only cross-GPU traffic is one all-reduce (sum, int64) of `span + 1` counters per motif --
NCCL over NVLink on GPUs, gloo in the CPU tests.  Every rank then derives the identical q-value table (K5) and
finalizes its own hits; rank 0 (or the caller) concatenates the per-rank tables and merges them by p-value.
"""
import os
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Dict[str, int]:
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return dict(rank=rank, world=world, local=local)


def init_comm(ctx, group=None):
    """Gives the library context `ctx` (engine.Context) its own NCCL communicator over the ranks of the current
    torch.distributed job: rank 0 creates the NCCL unique id, the process group only carries those 128 bytes (the
    rendezvous), and every later exchange of the scan runs inside the C ABI (gb2_allreduce_hist / gb2_allgather_bytes,
    csrc/comm.cu).  No-op on one rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        ctx.comm_init(None, 0, 1)
        return ctx
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.comm_init(box[0], rank, world)
    return ctx


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous, balanced split of n_items rows over the ranks (32-aligned so N-mask words are not shared)."""
    per = (n_items + world - 1) // world
    per = (per + 31) // 32 * 32
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def assign_chromosomes(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-first assignment of chromosomes (or regions) to ranks; returns the index list per rank."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: load[k])
        out[r].append(i)
        load[r] += lengths[i]
    return [sorted(x) for x in out]


def allreduce_histogram(hist: torch.Tensor, group=None, ctx=None) -> torch.Tensor:
    """Sum the per-rank score histograms in place (int64[span+1], last bin = N rows).  No-op on one rank.
    With a context that owns a communicator (init_comm) the all-reduce is issued by the library on the context's
    stream (gb2_allreduce_hist); otherwise torch.distributed does it (gloo in the CPU tests)."""
    if ctx is not None and getattr(ctx, "world", 1) > 1:
        return ctx.allreduce_hist(hist)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


def allreduce_max(value: float, device=None, group=None) -> float:
    """Max over ranks of a host scalar (used for device-timed durations)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item())
    return float(value)


def allreduce_sum(value: int, device=None, group=None) -> int:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.tensor([value], dtype=torch.int64, device=device or "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return int(t.item())
    return int(value)


def merge_hit_tables(tables: List[Dict[str, np.ndarray]]) -> Dict[str, np.ndarray]:
    """Concatenate per-rank hit tables (each sorted by p) and order by (p-value, row, strand): rows carry global
    row indices, so the result does not depend on how the rows were sharded."""
    keys = [k for k in tables[0].keys() if isinstance(tables[0][k], np.ndarray)]
    cat = {k: np.concatenate([t[k] for t in tables]) for k in keys}
    order = np.lexsort((cat["strand"], cat["row"], cat["p-value"]))
    return {k: v[order] for k, v in cat.items()}


def gather_hit_tables(table: Dict[str, np.ndarray], group=None) -> List[Dict[str, np.ndarray]]:
    """all_gather_object of the (small) per-rank hit tables."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        out = [None] * dist.get_world_size(group)
        dist.all_gather_object(out, table, group=group)
        return out
    return [table]
