"""Developer tool: pipeline timeline of the <=128-needle tensor scan (CTA 0), from a -DX2_TRACE build.
usage: python -m yume_pdq_b200.build variant trace X2_TRACE; YUME_PDQ_CUDA_LIB=.../libyume_pdq_cuda_trace.so python tools/x2_trace.py [q]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yume_pdq_b200 as yp

q = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = (10_000_000 // 2048 + 1) * 2048
db = yp.CudaVectorDatabase(synthetic_rows=n)
m = yp.CudaMatcher(db, 31)
out = torch.zeros((1 << 16, 16), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
needles = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(q, 32), dtype=np.uint8)).cuda()
for _ in range(3):
    m.find_device(needles, q, out, 1 << 16, cnt, algo=yp.ALGO_TENSOR_PAIRED)
torch.cuda.synchronize()
tr = np.zeros((8, 256), dtype=np.int64)
fn = yp.lib.yume_pdq_cuda_debug_x2_trace
fn.restype = C.c_int
fn.argtypes = [C.c_void_p]
assert fn(tr.ctypes.data) == 0
t0 = tr[1, 0]
names = {0: "exp: b_empty ok", 7: "exp: published", 1: "mma: b_full ok", 3: "mma: d_empty ok (2i+t)", 2: "mma: issued (2i+t)",
         4: "epi: d_full ok (2i+t)", 5: "epi: ld done", 6: "epi: tree done"}
for u in range(20, 28):
    print(f"unit {u}: exp wait-ok {tr[0,u]-t0:7d} pub {tr[7,u]-t0:7d} | mma bfull {tr[1,u]-t0:7d} "
          f"dE0 {tr[3,2*u]-t0:7d} iss0 {tr[2,2*u]-t0:7d} dE1 {tr[3,2*u+1]-t0:7d} iss1 {tr[2,2*u+1]-t0:7d} | "
          f"epi0 full {tr[4,2*u]-t0:7d} ld {tr[5,2*u]-t0:7d} tree {tr[6,2*u]-t0:7d}  epi1 full {tr[4,2*u+1]-t0:7d} ld {tr[5,2*u+1]-t0:7d} tree {tr[6,2*u+1]-t0:7d}")
print("per unit (clk):", (tr[1, 60] - tr[1, 20]) / 40)
