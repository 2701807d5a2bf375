import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import yume_pdq_b200 as yp
q = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
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
fn.restype = C.c_int; fn.argtypes = [C.c_void_p]
assert fn(tr.ctypes.data) == 0
t0 = tr[1, 0]
T = (q + 127) // 128
print("q", q, "tiles/unit", T)
for u in range(10, 18):
    print(f"unit {u}: exp wait-ok {tr[0,u]-t0:8d} pub {tr[7,u]-t0:8d} | mma bfull {tr[1,u]-t0:8d}")
print("per unit (clk):", (tr[1, 30] - tr[1, 10]) / 20)
# stage-level: role 3 (d_empty ok), 2 (issued), 4 (d_full), 5 (ld done), 6 (tree done) indexed by 2*ia+t, ia = tile counter
base = 2 * T * 10
for k in range(base, min(base + 2 * T + 4, 256)):
    print(f"  stage {k}: mma dE {tr[3,k]-t0:8d} iss {tr[2,k]-t0:8d} | epi full {tr[4,k]-t0:8d} ld {tr[5,k]-t0:8d} tree {tr[6,k]-t0:8d}")
