"""torchrun worker: compute_results over several GPUs (files split over the ranks, histogram all-reduced) must
return, on every rank, the same table as the reference produced for all files together.
    python -m torch.distributed.run --nproc-per-node 2 tests/dist_compute_results_worker.py <tmpdir>"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import golden_util as gu  # noqa: E402


class Args:
    def __init__(self, o):
        self.cores, self.threshold, self.noqvalue, self.qvalueT = 1, float(o["threshold"]), o["noqvalue"], o["qvalueT"]
        self.noreverse, self.recomb, self.verbose = o["noreverse"], o["recomb"], False


def main():
    import torch
    import torch.distributed as tdist
    from grafimo_b200 import dist as gdist
    from grafimo_b200 import motif_ops as mo
    from grafimo_b200.score_sequences import compute_results
    tmp = sys.argv[1]
    info = gdist.init_from_env("nccl")
    rank, world = info["rank"], info["world"]
    fx = gu.fixtures()
    for tag in ("fixture_plus_N_2files", "fixture_qvalT", "synth_w8"):
        c = gu.load_scoring(tag)
        g = gu.load_motif(c["motif_tag"])
        root = os.path.join(tmp, tag)
        d = os.path.join(root, "seqs", f"width_{g['width']}")
        if rank == 0:
            os.makedirs(d, exist_ok=True)
            lines = [ln for f in c["files"] for ln in f]
            k = 5  # more files than ranks, uneven sizes
            cuts = [0] + sorted(np.random.default_rng(1).choice(np.arange(1, len(lines)), size=k - 1, replace=False).tolist()) + [len(lines)]
            for i in range(k):
                with open(os.path.join(d, f"part{i}.tsv"), "w") as fh:
                    fh.write("\n".join(lines[cuts[i]:cuts[i + 1]]) + "\n")
            key = g["source"] if g["source"].endswith("_" + g["fmt"]) else g["source"] + "_meme"
            open(os.path.join(root, "m.meme"), "w").write(fx[key])
            open(os.path.join(root, "bg_nt"), "w").write(fx["bg_nt"])
        tdist.barrier()
        bg = "unfrm_dst" if g["bgfile"] == "unif" else os.path.join(root, "bg_nt")
        m = mo.build_motif_meme(os.path.join(root, "m.meme"), bg, g["pseudo"], g["no_reverse"], 1, False, True)[0]
        df = compute_results(m, os.path.join(root, "seqs"), True, Args(c["options"]))
        got = {col: df[col].to_numpy() for col in df.columns}
        gu.assert_tables_equal(got, c["table"], c["columns"])
        tdist.barrier()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
