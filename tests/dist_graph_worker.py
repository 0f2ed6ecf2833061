"""torchrun worker: the graph path over several GPUs.  Two synthetic chromosomes, one per rank: every rank extracts
and scores its own (K7 -> K2), the score histograms are all-reduced, and the table every rank returns must equal the
table one process computes from both chromosomes (written to <tmpdir>/expected.pkl by the launching test).
    python -m torch.distributed.run --nproc-per-node 2 tests/dist_graph_worker.py <tmpdir>"""
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import golden_util as gu  # noqa: E402
import graph_util as gr  # noqa: E402


class Args:
    cores, threshold, noqvalue, qvalueT, noreverse, recomb, verbose = 1, 0.05, False, False, False, False, False


def chromosomes():
    return [("1",) + gr.random_case(900, length=4000, n_var=200, n_hap=64), ("2",) + gr.random_case(901, length=2500, n_var=120, n_hap=64)]


def build_motif(tmp):
    from grafimo_b200 import motif_ops as mo
    p = os.path.join(tmp, "ctcf.meme")
    if not os.path.exists(p):
        open(p, "w").write(gu.fixtures()["ctcf_meme"])
    return mo.build_motif_meme(p, "unfrm_dst", 0.1, False, 1, False, True)[0]


def main():
    import torch.distributed as tdist
    from grafimo_b200 import dist as gdist
    from grafimo_b200 import score_sequences as ss
    from grafimo_b200.extract_regions import DeviceGraph
    tmp = sys.argv[1]
    info = gdist.init_from_env("nccl")
    rank, world = info["rank"], info["world"]
    ctx = ss._context()
    gdist.init_comm(ctx)  # histogram all-reduce and hit-column all-gather run inside the C ABI (csrc/comm.cu)
    assert ctx.world == world and ctx.rank == rank
    motif = build_motif(os.path.join(tmp, f"r{rank}"))
    mine = [c for i, c in enumerate(chromosomes()) if i % world == rank]
    rows = []
    for name, ref, vs, gt in mine:
        dg = DeviceGraph.build(ctx, name, ref, vs, gt=gt)
        rows.append(dg.extract([(0, len(ref) // 2), (len(ref) // 2 - 10, len(ref))], motif.width))
    df = ss.compute_results_rows(motif, rows, True, Args)
    exp = pickle.load(open(os.path.join(tmp, "expected.pkl"), "rb"))
    cols = [c for c in df.columns if c not in ("motif_id", "motif_alt_id")]
    gu.assert_tables_equal({c: df[c].to_numpy() for c in cols}, exp, cols)
    tdist.barrier()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
