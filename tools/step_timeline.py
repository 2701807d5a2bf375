"""Developer tool: CUDA-event timeline of the bench step (hash -> match) in steady state."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yume_pdq_b200 as yp
B = 4096
n = (10_000_000 // 2048 + 1) * 2048
dev = torch.device("cuda", 0)
frames = torch.empty((B, 512, 512), dtype=torch.float32, device=dev)
yp.synth_frames(frames, B, kind=1)
kern = yp.CudaKernel(0)
db = yp.CudaVectorDatabase(synthetic_rows=n); m = yp.CudaMatcher(db, 31)
outs = {"hashes": torch.empty((B, 8, 32), dtype=torch.uint8, device=dev), "quality": torch.empty((B, 8), dtype=torch.float32, device=dev),
        "thresholds": torch.empty((B, 8), dtype=torch.float32, device=dev)}
out = torch.zeros((1 << 16, 16), dtype=torch.uint8, device=dev); cnt = torch.zeros(1, dtype=torch.int32, device=dev)
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(12)]
for it in range(12):
    e = ev[it]
    e[0].record(); kern.hash_batch(frames, out=outs); e[1].record()
    nd = outs["hashes"].reshape(-1, 32); cnt.zero_()
    e[2].record(); m.find_device(nd, B * 8, out, 1 << 16, cnt, algo=yp.ALGO_AUTO); e[3].record()
torch.cuda.synchronize()
for it in range(4, 12):
    e = ev[it]
    nxt = ev[it + 1][0] if it + 1 < 12 else None
    print(f"step {it}: hash {e[0].elapsed_time(e[1]):7.3f} ms  gap {e[1].elapsed_time(e[2]):6.3f}  match {e[2].elapsed_time(e[3]):7.3f}"
          + (f"  step {e[0].elapsed_time(nxt):7.3f}" if nxt else ""))
