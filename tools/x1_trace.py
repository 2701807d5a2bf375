"""Developer tool: pipeline timeline of the rows-in-M tensor scan (CTA 0, worker group 0), from a -DX1_TRACE build.
usage: python -m yume_pdq_b200.build variant trace1 X1_TRACE; YUME_PDQ_CUDA_LIB=.../libyume_pdq_cuda_trace1.so python tools/x1_trace.py [q]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yume_pdq_b200 as yp

q = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = (10_000_000 // 2048 + 1) * 2048
db = yp.CudaVectorDatabase(synthetic_rows=n)
m = yp.CudaMatcher(db, 31)
out = torch.zeros((1 << 16, 16), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
needles = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(q, 32), dtype=np.uint8)).cuda()
for _ in range(3):
    m.find_device(needles, q, out, 1 << 16, cnt, algo=yp.ALGO_TENSOR_ROWS)
torch.cuda.synchronize()
tr = np.zeros((8, 256), dtype=np.int64)
fn = yp.lib.yume_pdq_cuda_debug_x1_trace
fn.restype = C.c_int
fn.argtypes = [C.c_void_p]
assert fn(tr.ctypes.data) == 0
t0 = tr[5, 0]
print("tile: prod r_empty-ok | worker: start r_full-ok published | mma: a_full-ok committed | worker: d_full-ok ld-done")
for u in list(range(0, 8)) + list(range(60, 72)):
    print(f"tile {u:3d}: prod {tr[0,u]-t0:7d} | w start {tr[5,u]-t0:7d} rfull {tr[6,u]-t0:7d} pub {tr[7,u]-t0:7d} | "
          f"mma afull {tr[1,u]-t0:7d} commit {tr[2,u]-t0:7d} | w dfull {tr[3,u]-t0:7d} ld {tr[4,u]-t0:7d}")
print("per tile of group 0 (clk):", (tr[7, 100] - tr[7, 40]) / 60)
