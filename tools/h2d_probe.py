#!/usr/bin/env python
"""Raw host->device copy rate of the box under N concurrent ranks -- the wall the end-to-end (`e2e`) numbers sit on.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py [--mb 1024]

Every rank copies the same amount from its own pinned buffer to its own GPU at the same time (barrier, CUDA events, max over
ranks), for three kinds of pinned memory: torch's (cudaHostAlloc default), write-combined (cudaHostAllocWriteCombined) and
cudaHostRegister over malloc'd pages first touched by this rank, and with the copy split over two streams.  No product code is
involved: if the per-GPU rate falls as N grows here, no staging scheme inside the library can win it back.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from grafimo_b200 import dist as gdist
    info = gdist.init_from_env("nccl")
    rank, world, local = info["rank"], info["world"], info["local"]
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nbytes = a.mb << 20
    rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(ptr, split):
        best = None
        for rep in range(a.reps):
            torch.cuda.synchronize()
            barrier()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(s1)
            if split:
                half = nbytes // 2
                s2.wait_event(e0)
                rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(half), 1, ctypes.c_void_p(s1.cuda_stream))
                rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr() + half), ctypes.c_void_p(ptr + half), ctypes.c_size_t(nbytes - half), 1, ctypes.c_void_p(s2.cuda_stream))
                e2.record(s2)
                s1.wait_event(e2)
            else:
                rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(nbytes), 1, ctypes.c_void_p(s1.cuda_stream))
            e1.record(s1)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if rep:  # first repetition is warm-up
                best = ms if best is None else min(best, ms)
        t = torch.tensor([best], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return nbytes / (float(t.item()) * 1e-3) / 1e9

    out = {"n_gpus": world, "mb_per_rank": a.mb}
    pinned = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    pinned.fill_(65)
    out["torch_pinned_gbs_per_gpu"] = timed(pinned.data_ptr(), False)
    out["torch_pinned_two_streams_gbs_per_gpu"] = timed(pinned.data_ptr(), True)
    del pinned
    p = ctypes.c_void_p()
    if rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), 4) == 0:  # cudaHostAllocWriteCombined
        ctypes.memset(p, 65, nbytes)
        out["write_combined_gbs_per_gpu"] = timed(p.value, False)
        rt.cudaFreeHost(p)
    libc = ctypes.CDLL("libc.so.6")
    libc.aligned_alloc.restype = ctypes.c_void_p
    q = libc.aligned_alloc(ctypes.c_size_t(2 << 20), ctypes.c_size_t(nbytes))
    if q:
        ctypes.memset(ctypes.c_void_p(q), 65, nbytes)  # first touch by this rank's thread
        if rt.cudaHostRegister(ctypes.c_void_p(q), ctypes.c_size_t(nbytes), 0) == 0:
            out["host_registered_gbs_per_gpu"] = timed(q, False)
            rt.cudaHostUnregister(ctypes.c_void_p(q))
        libc.free(ctypes.c_void_p(q))
    out["aggregate_gbs"] = out["torch_pinned_gbs_per_gpu"] * world
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
    except Exception:
        bus = None
    out["cpus_allowed"] = len(os.sched_getaffinity(0))
    if rank == 0:
        line = json.dumps(out)
        print(line)
        if a.out:
            with open(a.out, "a") as fh:
                fh.write(line + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
