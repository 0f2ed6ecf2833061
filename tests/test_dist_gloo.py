"""world_size-2 tests of the multi-GPU plumbing on CPU (gloo): sharding, the histogram all-reduce that makes
the q-values global, and the merge of per-rank hit tables.  The GPU kernels are not involved here; the BH
arithmetic from a histogram is restated in numpy and checked against the oracle's row-wise BH."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_util as gu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _bh_from_hist(ptab, hist):
    """numpy restatement of K5 (grafimo_b200/csrc/qvalue.cu): bins sorted by p ascending, running count,
    p / (C / float(N)), reverse running minimum, clip."""
    order = np.argsort(ptab, kind="stable")
    c = np.cumsum(hist[order])
    n = float(hist.sum())
    with np.errstate(divide="ignore", invalid="ignore"):
        raw = np.where(hist[order] > 0, ptab[order] / (c / n), np.inf)
    q = np.minimum.accumulate(raw[::-1])[::-1]
    q[q > 1] = 1
    out = np.empty_like(q)
    out[order] = q
    return out


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from grafimo_b200 import dist as gdist
    from oracle import oracle as orc
    info = gdist.init_from_env("gloo")
    assert info == dict(rank=rank, world=world, local=rank)
    m = gu.load_motif("ctcf_meme__bgnt")
    rng = np.random.default_rng(5)
    n = 6001
    seqs = ["".join(rng.choice(list("ACGT"), size=19)) for _ in range(n)]
    a = orc.kmers_to_matrix(seqs, 19)
    isc, lo, pv = orc.score_rows(a, m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    q_global = orc.bh(pv)
    nz = np.nonzero(m["pval_mat"])[0]
    lo_s, hi_s = nz[0], nz[-1]
    ptab = orc.pvalue_table(m["pval_mat"])[lo_s:hi_s + 1]
    # shard rows, build the local histogram, all-reduce, BH from the global histogram
    b, e = gdist.shard_bounds(n, rank, world)
    assert b % 32 == 0
    hist = torch.from_numpy(np.bincount(isc[b:e] - lo_s, minlength=hi_s - lo_s + 1).astype(np.int64))
    local_sum = int(hist.sum())
    gdist.allreduce_histogram(hist)
    assert int(hist.sum()) == n and local_sum == e - b
    assert gdist.allreduce_sum(e - b) == n
    assert gdist.allreduce_max(float(rank)) == float(world - 1)
    qtab = _bh_from_hist(ptab, hist.numpy())
    assert np.array_equal(qtab[isc[b:e] - lo_s], q_global[b:e])  # exact global q from the reduced histogram
    # per-rank hit tables -> merged table identical to the single-process one
    keep = np.nonzero(pv[b:e] < 0.05)[0] + b
    order = np.lexsort((keep, pv[keep]))
    keep = keep[order]
    table = {"row": keep.astype(np.uint64), "strand": np.zeros(len(keep), np.uint8), "p-value": pv[keep], "q-value": q_global[keep]}
    merged = gdist.merge_hit_tables(gdist.gather_hit_tables(table))
    allk = np.nonzero(pv < 0.05)[0]
    allk = allk[np.lexsort((allk, pv[allk]))]
    assert np.array_equal(merged["row"].astype(np.int64), allk) and np.array_equal(merged["q-value"], q_global[allk])
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_global_qvalues_over_two_ranks(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_shard_bounds_and_chromosome_assignment():
    from grafimo_b200 import dist as gdist
    for n, world in ((0, 2), (1, 2), (1000, 3), (2503954928, 8)):
        cover = []
        for r in range(world):
            lo, hi = gdist.shard_bounds(n, r, world)
            assert lo % 32 == 0 or lo == n
            cover.append((lo, hi))
        assert cover[0][0] == 0 and cover[-1][1] == n and all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
    hg38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
            135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167,
            46709983, 50818468, 156040895, 57227415]
    parts = gdist.assign_chromosomes(hg38, 8)
    assert sorted(i for p in parts for i in p) == list(range(24))
    loads = [sum(hg38[i] for i in p) for p in parts]
    assert max(loads) / (sum(hg38) / 8) < 1.08  # greedy by length balances the whole genome within 8 %
