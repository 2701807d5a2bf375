"""Where does the end-to-end host->device ceiling come from?  

Run under torchrun with N ranks (one per GPU).  Every rank copies a 512 MiB pinned host buffer to its GPU, alone and
then all ranks at once, for four kinds of page-locked memory:
  torch      torch.empty(pin_memory=True)                (cudaHostAlloc, default flags)
  portable   cudaHostAlloc(cudaHostAllocPortable)
  wc         cudaHostAlloc(cudaHostAllocPortable | cudaHostAllocWriteCombined)
  registered numpy buffer + cudaHostRegister
and prints one JSON object: per-rank GB/s alone, per-rank GB/s concurrent, aggregate, CPU affinity / NUMA node of the
rank and of the buffer's pages.  usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

BYTES = 1 << 29
REPS = 4


def numa_of_cpu(cpu):
    base = "/sys/devices/system/node"
    try:
        for n in os.listdir(base):
            if n.startswith("node") and os.path.exists(f"{base}/{n}/cpu{cpu}"):
                return int(n[4:])
    except OSError:
        pass
    return None


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rt = C.CDLL("libcudart.so.12")
    dst = torch.empty(BYTES, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    bufs = {}
    t = torch.empty(BYTES, dtype=torch.uint8, pin_memory=True)
    t.fill_(1)
    bufs["torch"] = (t.data_ptr(), t)
    for name, flags in (("portable", 1), ("wc", 1 | 4)):
        p = C.c_void_p()
        rc = rt.cudaHostAlloc(C.byref(p), C.c_size_t(BYTES), C.c_uint(flags))
        if rc == 0:
            C.memset(p, 1, BYTES)
            bufs[name] = (p.value, None)
    a = np.ones(BYTES, dtype=np.uint8)
    if rt.cudaHostRegister(C.c_void_p(a.ctypes.data), C.c_size_t(BYTES), C.c_uint(0)) == 0:
        bufs["registered"] = (a.ctypes.data, a)

    stream = torch.cuda.current_stream(dev).cuda_stream

    def copy_gbs(ptr):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rt.cudaMemcpyAsync(C.c_void_p(dst.data_ptr()), C.c_void_p(ptr), C.c_size_t(BYTES), C.c_int(1), C.c_void_p(stream))
        torch.cuda.synchronize()
        e0.record()
        for _ in range(REPS):
            rt.cudaMemcpyAsync(C.c_void_p(dst.data_ptr()), C.c_void_p(ptr), C.c_size_t(BYTES), C.c_int(1), C.c_void_p(stream))
        e1.record()
        torch.cuda.synchronize()
        return REPS * BYTES / (e0.elapsed_time(e1) * 1e-3) / 1e9

    res = {}
    for name, (ptr, _) in bufs.items():
        alone = None
        for r in range(world):           # one rank at a time
            barrier()
            if r == rank:
                alone = copy_gbs(ptr)
        barrier()
        together = copy_gbs(ptr)         # all ranks at once
        barrier()
        res[name] = {"alone_gbs": alone, "concurrent_gbs": together}
    aff = sorted(os.sched_getaffinity(0))
    cpu_now = os.sched_getcpu() if hasattr(os, "sched_getcpu") else None
    mine = {"rank": rank, "gpu": local, "cpu_affinity": f"{aff[0]}-{aff[-1]} ({len(aff)} cpus)", "running_on_cpu": cpu_now,
            "numa_node_of_cpu": numa_of_cpu(cpu_now) if cpu_now is not None else None, "h2d": res}
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    if rank == 0:
        nodes = sorted(n for n in os.listdir("/sys/devices/system/node") if n.startswith("node")) if os.path.isdir("/sys/devices/system/node") else []
        out = {"n_gpus": world, "bytes_per_copy": BYTES, "host_cpus": os.cpu_count(), "numa_nodes": nodes, "ranks": allr,
               "aggregate_concurrent_gbs": {k: sum(r["h2d"][k]["concurrent_gbs"] for r in allr if k in r["h2d"]) for k in res},
               "mean_alone_gbs": {k: float(np.mean([r["h2d"][k]["alone_gbs"] for r in allr if k in r["h2d"]])) for k in res}}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
