#!/usr/bin/env python3
"""A ctypes user of the reference's two-function C ABI, making the calls the reference's own
integration/hash.py:32-92 makes (same library lookup through $LIBYUME_PDQ_PATH, same argtypes, ctypes
buffers owned by the caller), but reading a prepared 512x512 luma frame (.npz with key "luma", or a raw
f32 file) instead of decoding a JPEG, so it runs where PIL and the reference tree are absent.

    LIBYUME_PDQ_PATH=yume_pdq_b200/libyume_pdq.so python integration/python/ref_abi_client.py frame.npz [iterations]
"""
import os
import sys
import time
from ctypes import CDLL, POINTER, byref, c_float, c_uint8, memmove, sizeof

import numpy as np

lib = CDLL(os.environ.get("LIBYUME_PDQ_PATH", "yume_pdq_b200/libyume_pdq.so"))
fn = lib.yume_pdq_hash_smart_kernel
fn.restype = c_float
fn.argtypes = [POINTER(c_float), POINTER(c_float), POINTER(c_uint8), POINTER(c_float), POINTER(c_float), POINTER(c_float)]

path = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
if path.endswith(".npz"):
    z = np.load(path)
    frame = np.ascontiguousarray(z[z.files[0]], dtype=np.float32).reshape(512, 512)
else:
    frame = np.fromfile(path, dtype=np.float32).reshape(512, 512)

threshold = c_float(0.0)
inp = (c_float * (512 * 512))()
buf1 = (c_float * (128 * 128))()
tmp = (c_float * 128)()
pdqf = (c_float * 256)()
out = (c_uint8 * 32)()
memmove(inp, frame.ctypes.data_as(POINTER(c_float)), sizeof(inp))


def once():
    q = fn(inp, byref(threshold), out, buf1, tmp, pdqf)
    return q, bytes(out).hex()


quality, hash_hex = once()
print(f"Quality: {quality:.3f}")
print(f"Threshold: {threshold.value:.3f}")
print(f"Hash: {hash_hex if quality > 0.5 else None}")
t0 = time.time()
for i in range(iters):
    q, h = once()
    assert h == hash_hex, f"hash changed at iteration {i}"
dt = time.time() - t0
print(f"Average hash time: {dt / max(iters, 1) * 1e6:.3f} us over {iters} calls")
