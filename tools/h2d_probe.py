#!/usr/bin/env python
"""Host-to-device ceiling of ONE box: N ranks (one per GPU), each with a pinned host buffer the size of a
1080p / 60 s clip (11.2 GB; --gb to change), concurrent plain cudaMemcpyAsync copies in 256 MiB chunks
and NO kernels.  Explains (or refutes) the end-to-end scaling of bench.py, whose per-clip time at N > 2
is H2D time:  aggregate GB/s here = the most any e2e path can move on this host.

    torchrun --nproc-per-node N tools/h2d_probe.py [--gb 11.2] [--reps 3] [--wc] [--affinity]

  --wc         allocate the pinned buffer write-combined (cudaHostAllocWriteCombined)
  --affinity   pin each rank's process to its share of the host cores before allocating (first-touch locality)
One JSON line on rank 0: per-rank GB/s (min / max), aggregate GB/s (sum of bytes / max time over ranks).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=11.2)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--wc", action="store_true")
    ap.add_argument("--affinity", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncpu = os.cpu_count() or 1
    if args.affinity:
        per = max(1, ncpu // max(world, 1))
        os.sched_setaffinity(0, set(range(local * per, min(ncpu, (local + 1) * per))))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = int(args.gb * 1e9) // 4096 * 4096
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    host_ptr = None
    if args.wc:
        # the CUDA runtime torch itself loaded (torch/lib or nvidia/cuda_runtime/lib): find it among the mapped libraries
        path = next(l.split()[-1] for l in open("/proc/self/maps") if "libcudart" in l)
        lib = ctypes.CDLL(path)
        p = ctypes.c_void_p()
        rc = lib.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04))      # cudaHostAllocWriteCombined
        assert rc == 0, f"cudaHostAlloc(write-combined) -> {rc}"
        host_ptr = p.value
        ctypes.memset(host_ptr, 1, nbytes)
        copy_lib = lib
    else:
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host.fill_(1)                                            # first touch by this (possibly core-pinned) process
        host_ptr = host.data_ptr()
        copy_lib = None
    stream = torch.cuda.Stream()
    chunk = 256 << 20

    def one_pass():
        with torch.cuda.stream(stream):
            for off in range(0, nbytes, chunk):
                n = min(chunk, nbytes - off)
                if copy_lib is None:
                    dev[off:off + n].copy_(host[off:off + n], non_blocking=True)     # one plain cudaMemcpyAsync per chunk
                else:
                    rc = copy_lib.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr() + off), ctypes.c_void_p(host_ptr + off),
                                                  ctypes.c_size_t(n), ctypes.c_int(1), ctypes.c_void_p(stream.cuda_stream))
                    assert rc == 0
        stream.synchronize()

    one_pass()
    times = []
    for _ in range(args.reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        one_pass()
        times.append(time.perf_counter() - t0)
    best = min(times)
    t = torch.tensor([best], dtype=torch.float64, device="cuda")
    tmax, tmin = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"probe": "h2d", "n_gpus": world, "gb_per_rank": nbytes / 1e9, "write_combined": bool(args.wc),
                          "affinity": bool(args.affinity), "host_cpus": ncpu,
                          "per_rank_gbs_min": nbytes / 1e9 / float(tmax.item()), "per_rank_gbs_max": nbytes / 1e9 / float(tmin.item()),
                          "aggregate_gbs": world * nbytes / 1e9 / float(tmax.item()), "ms_per_clip_slowest_rank": 1e3 * float(tmax.item())}),
              flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
