"""Per-call latency of the reference's single-frame C ABI (yume_pdq_hash_smart_kernel) through ctypes, for a frame in
pageable host memory, pinned host memory and device memory.  Developer tool; bench.py reports the same figures."""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import yume_pdq_b200 as yp  # noqa: E402

dev = torch.empty((4, 512, 512), dtype=torch.float32, device="cuda:0")
yp.synth_frames(dev, 4, first_frame=0, kind=1)
torch.cuda.synchronize()
one = dev[0].cpu().numpy()
pinned = torch.empty((512, 512), dtype=torch.float32).pin_memory()
pinned.copy_(dev[0])
torch.cuda.synchronize()
o_thr = ctypes.c_float(0.0)
o_hash = np.zeros(32, dtype=np.uint8)
o_buf1 = np.zeros((128, 128), dtype=np.float32)
o_tmp = np.zeros(128, dtype=np.float32)
o_pdqf = np.zeros((16, 16), dtype=np.float32)
fn = yp.lib.yume_pdq_hash_smart_kernel
outs = (ctypes.byref(o_thr), ctypes.c_void_p(o_hash.ctypes.data), ctypes.c_void_p(o_buf1.ctypes.data),
        ctypes.c_void_p(o_tmp.ctypes.data), ctypes.c_void_p(o_pdqf.ctypes.data))


def per_call(ptr, n=500):
    src = ctypes.c_void_p(ptr)
    for _ in range(20):
        fn(src, *outs)
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn(src, *outs)
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    st = (ctypes.c_double * 4)()
    kc = (ctypes.c_longlong * 7)()
    yp.lib.yume_pdq_cuda_one_frame_last_timing(st, kc)
    print("   last call: copies issued %.1f us, launched %.1f, answer seen %.1f, outputs copied %.1f" % tuple(st))
    print("   finishing CTA, SM clocks since its start: elected %d, staged %d, dct1 %d, dct2 %d, median %d, hash %d, gradient sum %d" % tuple(kc))
    return "median %.1f us  mean %.1f  p10 %.1f  p90 %.1f" % (np.median(ts), ts.mean(), np.percentile(ts, 10), np.percentile(ts, 90))


print("pageable input:", per_call(one.ctypes.data))
h0 = o_hash.copy()
print("pinned input:  ", per_call(pinned.data_ptr()))
assert (o_hash == h0).all()
print("device input:  ", per_call(dev[0].data_ptr()))
assert (o_hash == h0).all()
# host memcpy of one frame into a pinned block, for scale
dst = torch.empty((512, 512), dtype=torch.float32).pin_memory().numpy()
ts = []
for _ in range(200):
    t0 = time.perf_counter()
    np.copyto(dst, one)
    ts.append(time.perf_counter() - t0)
print("host memcpy of 1 MiB into pinned memory: median %.1f us" % (np.median(ts) * 1e6))

# the second function of the reference ABI: 7 callbacks per call
seen = []
p0 = o_pdqf.copy()
thr = float(o_thr.value)
for _ in range(20):
    p = p0.copy()
    yp.visit_dihedrals(p, thr, lambda *a: True)
ts = []
for _ in range(300):
    p = p0.copy()
    t0 = time.perf_counter()
    yp.visit_dihedrals(p, thr, lambda *a: True)
    ts.append(time.perf_counter() - t0)
print("visit_dihedrals (7 Python callbacks): median %.1f us" % (np.median(ts) * 1e6))
