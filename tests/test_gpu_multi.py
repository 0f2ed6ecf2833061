"""Multi-GPU tests on hardware (skipped on a one-GPU box; `bench.py --gpus N` carries the same parity check in its
`parity` key so that the driver's scaling runs prove it too): the C-ABI communicator through ctypes at world 2, the
`--gpus 2` command line (motif collection sharded by motif) against the one-GPU run, and the batched motif entry points."""
import os
import subprocess
import sys

import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ctx():
    from grafimo_b200.engine import Context
    from grafimo_b200 import score_sequences as ss
    c = Context(0)
    yield c
    if ss._ctx is c:  # tests below hand this context to the scoring seams: do not leave a closed one behind
        ss._ctx = None
    c.close()


def _torchrun(worker, tmp_path, port, n=2):
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(HERE, worker), str(tmp_path)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r


def test_c_abi_communicator_world_2(tmp_path):
    """gb2_comm_init / gb2_allreduce_hist / gb2_allreduce_max_f64 / gb2_allgather_bytes over two GPUs + the rank-merged hit
    table == the single-GPU table, bit for bit"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _torchrun("dist_comm_worker.py", tmp_path, 29581)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_comm_world_1_is_a_no_op(ctx):
    """one GPU: gb2_comm_init without an id, collectives are no-ops / local copies (through ctypes)"""
    ctx.comm_init(None, 0, 1)
    h = torch.arange(100, dtype=torch.int64, device=ctx.device)
    ctx.allreduce_hist(h)
    g = ctx.allgather(h)
    ctx.sync()
    assert torch.equal(h.cpu(), torch.arange(100)) and g.shape == (1, 100) and torch.equal(g[0], h)
    assert ctx.allreduce_max([2.5, -1.0]) == [2.5, -1.0]
    from grafimo_b200._lib import GrafimoB200Error
    with pytest.raises(GrafimoB200Error):
        ctx.comm_init(None, 3, 2)


def test_k4_exact_host_path_equals_chain_kernels(ctx, monkeypatch):
    """K4 takes integer suffix sums on the host when no addition can round (uniform background: values are multiples of
    4^-w and the total fits 53 bits) and the order-preserving device kernels otherwise; forcing every motif through the
    kernels must give the same tables, and both are the oracle's (sequential sums per start score)"""
    from grafimo_b200.engine import DeviceMotif
    from oracle import oracle as orc
    tags = ["ctcf_meme__unif", "gata1_meme__unif", "synth_w8_meme__unif", "ctcf_meme__bgnt", "synth_w35_meme__unif_norev", "synth_w30_meme__bgnt"]
    ms = [gu.load_motif(t) for t in tags]
    items = [(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"]) for m in ms]
    monkeypatch.delenv("GB2_K4_FORCE_CHAIN", raising=False)
    fast = DeviceMotif.create_many(ctx, items)
    monkeypatch.setenv("GB2_K4_FORCE_CHAIN", "1")
    chain = DeviceMotif.create_many(ctx, items)
    for m, a, b in zip(ms, fast, chain):
        tab = orc.pvalue_table(m["pval_mat"])[a.lo:a.hi + 1]
        assert np.array_equal(a.ptable, b.ptable) and np.array_equal(a.ptable, tab), m["tag"]
        assert a.info.total == b.info.total and a.info.monotone == b.info.monotone


def test_batched_motif_create_equals_single(ctx):
    """gb2_motif_create_batched (one allocation / upload / K4 launch pair for the collection) == gb2_motif_create per motif:
    p-value tables bit-equal, same chunk plan; K4 stays pinned on the oracle for every golden motif"""
    from grafimo_b200.engine import DeviceMotif
    from oracle import oracle as orc
    ms = [gu.load_motif(t) for t in gu.motif_tags()]
    items = [(m["score_matrix"], m["pval_mat"], m["min_val"], m["scale"], m["offset"]) for m in ms]
    many = DeviceMotif.create_many(ctx, items)
    for m, it, dm in zip(ms, items, many):
        one = ctx.motif(*it)
        assert np.array_equal(dm.ptable, one.ptable), m["tag"]
        tab = orc.pvalue_table(m["pval_mat"])
        assert np.array_equal(dm.ptable, tab[dm.lo:dm.hi + 1]), m["tag"]
        for f in ("width", "n_chunks", "chunk_bases", "lut_replicas", "monotone", "lo", "hi", "span", "smem_bytes", "total"):
            assert getattr(dm.info, f) == getattr(one.info, f), (m["tag"], f)
        assert dm.info.chunk_bases in (3, 4) and dm.info.n_chunks == -(-dm.info.width // dm.info.chunk_bases)
        if dm.info.width > 32:
            assert dm.info.chunk_bases == 3
    del many[3]  # handles share one allocation: dropping one must not free the others' tables
    assert np.array_equal(many[3].ptable, ctx.motif(*items[4]).ptable)
    assert DeviceMotif.create_many(ctx, []) == []
    from grafimo_b200._lib import GrafimoB200Error
    bad = list(items[:3])
    bad[1] = (items[1][0], items[1][1], items[1][2], 0, items[1][4])  # not scaled
    with pytest.raises(GrafimoB200Error):
        DeviceMotif.create_many(ctx, bad)


def test_many_motif_cli_route_equals_per_motif_runs(ctx, tmp_path):
    """`findmotif` on a MEME file with three motifs and a selective threshold: ONE many-motif scan (ManyScan: batched upload,
    one K5 launch, one sort) writes the same report files as three single-motif runs"""
    import gzip
    from grafimo_b200 import score_sequences as ss
    from grafimo_b200.__main__ import main
    ss._ctx = ctx
    fx = gu.fixtures()
    (tmp_path / "test.fa").write_text(fx["test_fa"])
    with gzip.open(tmp_path / "test.vcf.gz", "wt") as fh:
        fh.write(fx["test_vcf"])
    (tmp_path / "r.bed").write_text("chrx\t0\t50\n")
    blocks = {"ctcf": fx["ctcf_meme"], "atf3": fx["atf3_meme"], "w8": fx["synth_w8_meme"]}
    head = blocks["ctcf"][:blocks["ctcf"].index("MOTIF")]
    body = {k: v[v.index("MOTIF"):].rstrip("\n") for k, v in blocks.items()}
    (tmp_path / "three.meme").write_text(head + "\n\n".join(body.values()) + "\n")
    common = ["-l", str(tmp_path / "test.fa"), "-v", str(tmp_path / "test.vcf.gz"), "-b", str(tmp_path / "r.bed"), "-t", "0.2",
              "--recomb", "--debug"]
    assert main(["findmotif", "-m", str(tmp_path / "three.meme"), "-o", str(tmp_path / "many")] + common) == 0
    files = sorted(p.name for p in (tmp_path / "many").iterdir() if p.suffix in (".tsv", ".gff"))
    assert len(files) == 6
    for k in body:
        (tmp_path / f"{k}.meme").write_text(head + body[k] + "\n")
        assert main(["findmotif", "-m", str(tmp_path / f"{k}.meme"), "-o", str(tmp_path / f"one_{k}")] + common) == 0
        for ext in ("tsv", "gff"):
            one = (tmp_path / f"one_{k}" / f"grafimo_out.{ext}").read_bytes()
            match = [f for f in files if f.endswith(ext) and (tmp_path / "many" / f).read_bytes() == one]
            assert len(match) == 1 and len(one) > 200, (k, ext)


def test_cli_gpus_2_shards_the_collection_by_motif(tmp_path):
    """`findmotif --gpus 2` (self-launched torchrun, motifs greedily assigned to the ranks, every rank writes its own
    reports) == the one-GPU run, file for file"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import gzip
    fx = gu.fixtures()
    (tmp_path / "test.fa").write_text(fx["test_fa"])
    with gzip.open(tmp_path / "test.vcf.gz", "wt") as fh:
        fh.write(fx["test_vcf"])
    (tmp_path / "r.bed").write_text("chrx\t0\t50\n")
    blocks = [fx["ctcf_meme"], fx["atf3_meme"], fx["synth_w8_meme"], fx["gata1_meme"], fx["synth_w11_meme"]]
    head = blocks[0][:blocks[0].index("MOTIF")]
    (tmp_path / "five.meme").write_text(head + "\n\n".join(b[b.index("MOTIF"):].rstrip("\n") for b in blocks) + "\n")
    common = ["-m", str(tmp_path / "five.meme"), "-l", str(tmp_path / "test.fa"), "-v", str(tmp_path / "test.vcf.gz"),
              "-b", str(tmp_path / "r.bed"), "-t", "0.2", "--recomb"]
    env = dict(os.environ, PYTHONPATH=os.path.dirname(HERE))
    for gpus in (1, 2):
        r = subprocess.run([sys.executable, "-m", "grafimo_b200", "findmotif", "--gpus", str(gpus), "-o", str(tmp_path / f"g{gpus}")] + common,
                           capture_output=True, text=True, timeout=900, env=env, cwd=os.path.dirname(HERE))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    a = sorted(p.name for p in (tmp_path / "g1").iterdir() if p.suffix in (".tsv", ".gff"))
    b = sorted(p.name for p in (tmp_path / "g2").iterdir() if p.suffix in (".tsv", ".gff"))
    assert a == b and len(a) == 10
    for f in a:
        assert (tmp_path / "g1" / f).read_bytes() == (tmp_path / "g2" / f).read_bytes(), f
