"""Short driver for ncu captures: launches each hot kernel a few times at profile-friendly sizes.
usage: python tools/profile_kernels.py [hash] [scan] [stream] [tiled] [exact] [tensor]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yume_pdq_b200 as yp

what = sys.argv[1:] or ["hash", "scan", "stream", "tiled"]
rng = np.random.default_rng(0)
dev = torch.device("cuda", 0)
if "hash" in what:
    n = 1184   # 2 full waves of 4 CTAs/SM x 148 SMs
    frames = torch.empty((n, 512, 512), dtype=torch.float32, device=dev)
    yp.synth_frames(frames, n, kind=1)
    k = yp.CudaKernel(0)
    outs = {"hashes": torch.empty((n, 8, 32), dtype=torch.uint8, device=dev),
            "quality": torch.empty((n, 8), dtype=torch.float32, device=dev),
            "thresholds": torch.empty((n, 8), dtype=torch.float32, device=dev)}
    for _ in range(3):
        k.hash_batch(frames, out=outs)
    torch.cuda.synchronize()
if {"scan", "stream", "tiled", "exact", "tensor", "tensor128", "tensor32k", "rows8"} & set(what):
    n = (10_000_000 // 2048 + 1) * 2048
    db = yp.CudaVectorDatabase(synthetic_rows=n)
    m = yp.CudaMatcher(db, 31)
    out = torch.zeros((1 << 16, 16), dtype=torch.uint8, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    n8 = torch.from_numpy(rng.integers(0, 256, size=(8, 32), dtype=np.uint8)).cuda()
    nbig = torch.from_numpy(rng.integers(0, 256, size=(2048, 32), dtype=np.uint8)).cuda()
    nbig4 = torch.from_numpy(rng.integers(0, 256, size=(4096, 32), dtype=np.uint8)).cuda()
    for _ in range(3):
        if "scan" in what:
            m.find_device(n8, 8, out, 1 << 16, cnt, algo=yp.ALGO_TILED)
        if "rows8" in what:       # rows-in-M tensor scan (matcher_tc1.cu)
            m.find_device(n8, 8, out, 1 << 16, cnt, algo=yp.ALGO_TENSOR_ROWS)
        if "exact" in what:
            m.find_device(n8, 8, out, 1 << 16, cnt, algo=yp.ALGO_EXACT)
        if "tiled" in what:
            m.find_device(nbig, 2048, out, 1 << 16, cnt, algo=yp.ALGO_TILED)
        if "tensor" in what:
            m.find_device(nbig4, 4096, out, 1 << 16, cnt, algo=yp.ALGO_TENSOR)
        if "tensor32k" in what:     # the bench's needle batch
            if "nb32" not in globals():
                nb32 = torch.from_numpy(rng.integers(0, 256, size=(32768, 32), dtype=np.uint8)).cuda()
            m.find_device(nb32, 32768, out, 1 << 16, cnt, algo=yp.ALGO_TENSOR)
        if "tensor128" in what:     # one needle tile per unit: the X2Small configuration of the paired kernel
            m.find_device(nbig[:128].contiguous(), 128, out, 1 << 16, cnt, algo=yp.ALGO_TENSOR_PAIRED)
    torch.cuda.synchronize()
print("done")

if "one" in what:        # single-frame latency kernel (hash_one_kernel), frame already in device memory
    f = torch.empty((1, 512, 512), dtype=torch.float32, device=dev)
    yp.synth_frames(f, 1, kind=1)
    torch.cuda.synchronize()
    for _ in range(4):
        yp.hash_get_threshold(f[0])
if "mailbox1" in what:   # the peer-memory query folded into the rows-in-M scan, a world of one rank
    n = (10_000_000 // 2048 + 1) * 2048
    dbm = yp.CudaVectorDatabase(synthetic_rows=n)
    mb = yp.PeerMailbox(0, 0, 1)
    mb.connect([mb.handle])
    q8 = rng.integers(0, 256, size=(8, 32), dtype=np.uint8)
    for _ in range(4):
        mb.scan(dbm, q8, 8, 31, root=0)
    mb.close()
    dbm.close()
