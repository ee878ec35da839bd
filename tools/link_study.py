#!/usr/bin/env python
"""Host-link study: pinned-host -> device copy rate per GPU with 1 / 2 / 4 / all ranks copying at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/link_study.py > profiles/r2_link_study.json

For every subset size k (ranks 0..k-1 copy, the others wait at the barrier) and every variant

    default     cudaHostAlloc default flags, one stream, the whole 3.6 GB in one cudaMemcpyAsync
    wc          write-combined pinned memory (cudaHostAllocWriteCombined)
    streams2    two streams, each one half of the buffer
    chunk1024   one stream, back-to-back copies of 1024 windows (369 MB), as the host-fed calls issue them
    chunk2500   the same with 2500 windows per copy

rank 0 prints one JSON object: GB/s per active rank (CUDA events on the copy stream, max over three repeats).
This separates what the box gives (links shared between GPUs, a socket hop) from what the library's
copy schedule could change.  torch is plumbing (process group, barrier); the copies go through cudart.
"""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

FL, NWIN = 45000, 10000
BYTES = FL * 8 * NWIN


def main():
    # stdout carries the JSON only: libraries (NCCL's version line) write to fd 1, which is pointed at stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rt = C.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else C.CDLL("libcudart.so")
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    rt.cudaFreeHost.argtypes = [C.c_void_p]
    dst = torch.empty(BYTES, dtype=torch.uint8, device=dev)
    bufs = {}
    for name, flags in (("default", 0), ("wc", 4)):
        p = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(p), BYTES, flags) == 0
        C.memset(p, 1, BYTES)
        bufs[name] = p
    s0, s1 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def copy(host, off, n, stream):
        assert rt.cudaMemcpyAsync(C.c_void_p(dst.data_ptr() + off), C.c_void_p(host.value + off), n, 1, C.c_void_p(stream.cuda_stream)) == 0

    def variant(name):
        if name == "default" or name == "wc":
            copy(bufs[name], 0, BYTES, s0)
        elif name == "streams2":
            h = BYTES // 2
            copy(bufs["default"], 0, h, s0)
            copy(bufs["default"], h, BYTES - h, s1)
        else:
            step = int(name[5:]) * FL * 8
            for off in range(0, BYTES, step):
                copy(bufs["default"], off, min(step, BYTES - off), s0)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    table = {}
    ks = sorted({k for k in (1, 2, 4, 8, world) if k <= world})
    for k in ks:
        for name in ("default", "wc", "streams2", "chunk1024", "chunk2500"):
            best = 0.0
            for _ in range(3):
                barrier()
                if rank < k:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(s0)
                    s1.wait_event(e0)
                    variant(name)
                    ej = torch.cuda.Event()
                    ej.record(s1)
                    s0.wait_event(ej)
                    e1.record(s0)
                    torch.cuda.synchronize(dev)
                    best = max(best, BYTES / (e0.elapsed_time(e1) * 1e-3) / 1e9)
                barrier()
            if world > 1:
                t = torch.tensor([best], device=dev, dtype=torch.float64)
                out = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(out, t)
                rates = [round(float(v.item()), 2) for v in out][:k]
            else:
                rates = [round(best, 2)]
            table.setdefault(str(k), {})[name] = rates
    if rank == 0:
        print(json.dumps(dict(bytes_per_copy=BYTES, world=world, gbs_per_active_rank=table,
                              note="ranks 0..k-1 copy at once; CUDA events on the copy stream; best of 3")))
    for p in bufs.values():
        rt.cudaFreeHost(p)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
