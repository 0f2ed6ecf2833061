"""torchrun worker: the C-ABI communicator (csrc/comm.cu) over real GPUs, through ctypes.
    python -m torch.distributed.run --nproc-per-node 2 tests/dist_comm_worker.py <tmpdir>
torch.distributed only carries the 128-byte NCCL id (dist.init_comm); then
  * gb2_allreduce_hist sums known per-rank uint64 counters, gb2_allreduce_max_f64 and gb2_allgather_bytes are checked
    against closed forms;
  * bench.parity_multi_gpu: rows split over the ranks + all-reduced histogram + per-rank finalize + gathered columns ==
    one single-GPU scan of the same rows, bit for bit (q-table of every rank, row / strand / score / p / q of every hit)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    import numpy as np
    import torch
    import torch.distributed as tdist
    import bench
    import golden_util as gu
    from grafimo_b200 import dist as gdist
    from grafimo_b200 import engine
    tmp = sys.argv[1]
    info = gdist.init_from_env("nccl")
    rank, world = info["rank"], info["world"]
    torch.cuda.set_device(info["local"])
    ctx = engine.Context(info["local"])
    gdist.init_comm(ctx)
    assert (ctx.rank, ctx.world) == (rank, world)
    n = 7425
    h = (torch.arange(n, dtype=torch.int64, device=ctx.device) * (rank + 1) + (1 << 40) * rank)
    ctx.allreduce_hist(h)
    ctx.sync()
    tri, sq = world * (world + 1) // 2, world * (world - 1) // 2
    assert torch.equal(h.cpu(), torch.arange(n, dtype=torch.int64) * tri + (1 << 40) * sq)
    assert ctx.allreduce_max([float(rank), -1.0 * rank, 3.5]) == [float(world - 1), 0.0, 3.5]
    g = ctx.allgather(torch.full((5, 3), rank, dtype=torch.int32, device=ctx.device))
    ctx.sync()
    assert g.shape == (world, 5, 3) and all(int(g[r].min()) == r == int(g[r].max()) for r in range(world))
    m = gu.load_motif("ctcf_meme__unif")
    dm = ctx.motif(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"])
    res = bench.parity_multi_gpu(ctx, dm, rank, world)
    if rank == 0:
        assert res["ok"] and res["hits"] > 1000 and res["qtable_equal_on_every_rank"], res
        print("parity", res)
    tdist.barrier()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    ctx.close()
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
