"""ctypes loader for the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by the
product package yume_pdq_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")


def build(force: bool = False) -> None:
    """Compile the oracle with the committed Makefile (gcc only)."""
    if force:
        subprocess.check_call(["make", "-C", HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HERE, "all"], stdout=subprocess.DEVNULL)


class Match(C.Structure):
    _fields_ = [("row", C.c_uint64), ("needle", C.c_uint32), ("dist", C.c_uint32)]


MATCH_DTYPE = np.dtype([("row", "<u8"), ("needle", "<u4"), ("dist", "<u4")])

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_DIH_CB = C.CFUNCTYPE(C.c_bool, C.c_void_p, C.c_uint32, C.c_float, C.c_float)


def _fp(a):
    return a.ctypes.data_as(_f32p)


def _bp(a):
    return a.ctypes.data_as(_u8p)


class Oracle:
    def __init__(self, path: str | None = None):
        path = path or os.path.join(BUILD, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        L.oracle_jarosz_compress.argtypes = [_f32p, _f32p, C.c_size_t]
        L.oracle_dct2d.argtypes = [_f32p, C.c_size_t, _f32p, _f32p]
        L.oracle_sum_of_gradients.argtypes = [_f32p]
        L.oracle_sum_of_gradients.restype = C.c_float
        L.oracle_adjust_quality.argtypes = [C.c_float]
        L.oracle_adjust_quality.restype = C.c_float
        L.oracle_torben_median.argtypes = [_f32p]
        L.oracle_torben_median.restype = C.c_float
        L.oracle_quantize.argtypes = [_f32p, _f32p, _u8p]
        L.oracle_threshold_2d_f32.argtypes = [_f32p, _u8p, C.c_float]
        for n in ("oracle_negate_alt_cols", "oracle_negate_alt_rows",
                  "oracle_negate_alt_rows_trait_default"):
            getattr(L, n).argtypes = [_f32p, C.c_int]
        L.oracle_pdqf_t.argtypes = [_f32p]
        L.oracle_negate_off_diagonals.argtypes = [_f32p]
        L.oracle_hash_get_threshold.argtypes = [_f32p, _f32p, _u8p, _f32p, _f32p, _f32p]
        L.oracle_hash_get_threshold.restype = C.c_float
        L.oracle_visit_dihedrals.argtypes = [C.c_void_p, C.c_float, _u8p, _f32p, _DIH_CB]
        L.oracle_visit_dihedrals.restype = C.c_bool
        L.oracle_hash_all_dihedrals.argtypes = [_f32p, _u8p, _f32p, _f32p, _f32p]
        L.oracle_matcher_find.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.c_uint32,
                                          C.POINTER(Match), C.c_size_t]
        L.oracle_matcher_find.restype = C.c_size_t
        L.oracle_matcher_scan.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.c_uint32]
        L.oracle_matcher_scan.restype = C.c_int
        L.oracle_matcher_find_min.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.c_uint32]
        L.oracle_matcher_find_min.restype = C.c_uint64
        L.oracle_synth_db.argtypes = [_u8p, C.c_size_t, C.c_uint64, C.c_uint64]
        L.oracle_synth_frames.argtypes = [_f32p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int]
        L.oracle_synth_pixels.argtypes = [_u8p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int]
        L.oracle_cvt_to_luma8f.argtypes = [_u8p, C.c_int, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32,
                                           C.c_int, _f32p]
        self.luma_table_itu = np.array(list((C.c_float * 3).in_dll(L, "ORACLE_RGB8_TO_LUMA8_TABLE_ITU")),
                                       dtype=np.float32)
        self.dihedral_order = list((C.c_uint32 * 8).in_dll(L, "ORACLE_DIHEDRAL_ORDER"))

    # ---- hashing ---------------------------------------------------------
    def jarosz_compress(self, frame: np.ndarray, stride: int = 128) -> np.ndarray:
        frame = np.ascontiguousarray(frame, dtype=np.float32).reshape(512, 512)
        out = np.zeros((128 if stride == 128 else 127, stride), dtype=np.float32)
        self.lib.oracle_jarosz_compress(_fp(frame), _fp(out), stride)
        return out

    def dct2d(self, buf: np.ndarray) -> np.ndarray:
        buf = np.ascontiguousarray(buf, dtype=np.float32)
        tmp = np.zeros(128, dtype=np.float32)
        pdqf = np.zeros((16, 16), dtype=np.float32)
        self.lib.oracle_dct2d(_fp(buf), buf.shape[1], _fp(tmp), _fp(pdqf))
        return pdqf

    def sum_of_gradients(self, pdqf) -> float:
        pdqf = np.ascontiguousarray(pdqf, dtype=np.float32)
        return float(self.lib.oracle_sum_of_gradients(_fp(pdqf)))

    def adjust_quality(self, g: float) -> float:
        return float(self.lib.oracle_adjust_quality(C.c_float(g)))

    def torben_median(self, pdqf) -> float:
        pdqf = np.ascontiguousarray(pdqf, dtype=np.float32)
        return float(self.lib.oracle_torben_median(_fp(pdqf)))

    def quantize(self, pdqf):
        pdqf = np.ascontiguousarray(pdqf, dtype=np.float32)
        thr = C.c_float(0)
        out = np.zeros(32, dtype=np.uint8)
        self.lib.oracle_quantize(_fp(pdqf), C.byref(thr), _bp(out))
        return out, float(thr.value)

    def threshold_2d_f32(self, pdqf, threshold: float):
        pdqf = np.ascontiguousarray(pdqf, dtype=np.float32)
        out = np.zeros(32, dtype=np.uint8)
        self.lib.oracle_threshold_2d_f32(_fp(pdqf), _bp(out), C.c_float(threshold))
        return out

    def _inplace(self, name, pdqf, *args):
        p = np.array(pdqf, dtype=np.float32, order="C", copy=True)
        getattr(self.lib, name)(_fp(p), *args)
        return p

    def negate_alt_cols(self, pdqf, negate: bool):
        return self._inplace("oracle_negate_alt_cols", pdqf, int(negate))

    def negate_alt_rows(self, pdqf, negate: bool):
        return self._inplace("oracle_negate_alt_rows", pdqf, int(negate))

    def negate_alt_rows_trait_default(self, pdqf, negate: bool):
        return self._inplace("oracle_negate_alt_rows_trait_default", pdqf, int(negate))

    def pdqf_t(self, pdqf):
        return self._inplace("oracle_pdqf_t", pdqf)

    def negate_off_diagonals(self, pdqf):
        return self._inplace("oracle_negate_off_diagonals", pdqf)

    def cvt_to_luma8f(self, pixels: np.ndarray, channels: int, coeff=None, fma: bool = False) -> np.ndarray:
        """u8 [..., channels] (channels 3/4) or u8 [...] (channels 1) -> f32 luma, mod.rs:303-340 /
        main.rs:508-510; fma=True restates the AVX2 override x86.rs:686-753."""
        pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
        shape = pixels.shape[:-1] if channels > 1 else pixels.shape
        cf = self.luma_table_itu if coeff is None else np.asarray(coeff, dtype=np.float32)
        bits = [int(b) for b in cf.view(np.uint32)]
        out = np.zeros(shape, dtype=np.float32)
        self.lib.oracle_cvt_to_luma8f(_bp(pixels), channels, out.size, bits[0], bits[1], bits[2], int(fma),
                                      _fp(out))
        return out

    def hash_get_threshold(self, frame):
        """-> (quality, threshold, hash[32] u8, pdqf[16,16]) like lib.rs:427-455."""
        frame = np.ascontiguousarray(frame, dtype=np.float32).reshape(512, 512)
        buf1 = np.zeros((128, 128), dtype=np.float32)
        tmp = np.zeros(128, dtype=np.float32)
        pdqf = np.zeros((16, 16), dtype=np.float32)
        out = np.zeros(32, dtype=np.uint8)
        thr = C.c_float(0)
        q = self.lib.oracle_hash_get_threshold(_fp(frame), C.byref(thr), _bp(out), _fp(buf1),
                                               _fp(tmp), _fp(pdqf))
        return float(q), float(thr.value), out, pdqf

    def visit_dihedrals(self, pdqf, threshold: float, stop_after: int | None = None):
        """-> list of (dihedral, threshold, quality, hash, pdqf_state) for the 7 callbacks."""
        p = np.array(pdqf, dtype=np.float32, order="C", copy=True)
        out = np.zeros(32, dtype=np.uint8)
        seen = []

        def cb(_ctx, dihedral, thr, quality):
            seen.append((int(dihedral), float(thr), float(quality), out.copy(), p.copy()))
            return stop_after is None or len(seen) < stop_after

        ok = self.lib.oracle_visit_dihedrals(None, C.c_float(threshold), _bp(out), _fp(p),
                                             _DIH_CB(cb))
        return bool(ok), seen, p

    def hash_all_dihedrals(self, frames: np.ndarray):
        """frames [n,512,512] f32 -> hashes [n,8,32] u8, quality [n,8], thr [n,8], pdqf [n,16,16]."""
        frames = np.ascontiguousarray(frames, dtype=np.float32).reshape(-1, 512, 512)
        n = frames.shape[0]
        hashes = np.zeros((n, 8, 32), dtype=np.uint8)
        quality = np.zeros((n, 8), dtype=np.float32)
        thr = np.zeros((n, 8), dtype=np.float32)
        pdqf = np.zeros((n, 16, 16), dtype=np.float32)
        for i in range(n):
            self.lib.oracle_hash_all_dihedrals(_fp(frames[i]), _bp(hashes[i]), _fp(quality[i]),
                                               _fp(thr[i]), _fp(pdqf[i]))
        return hashes, quality, thr, pdqf

    # ---- matching --------------------------------------------------------
    def matcher_find(self, rows: np.ndarray, needles: np.ndarray, threshold: int,
                     cap: int = 1 << 20) -> np.ndarray:
        rows = np.ascontiguousarray(rows, dtype=np.uint8).reshape(-1, 32)
        needles = np.ascontiguousarray(needles, dtype=np.uint8).reshape(-1, 32)
        out = np.zeros(cap, dtype=MATCH_DTYPE)
        n = self.lib.oracle_matcher_find(_bp(rows), rows.shape[0], _bp(needles), needles.shape[0],
                                         threshold, out.ctypes.data_as(C.POINTER(Match)), cap)
        if n > cap:
            return self.matcher_find(rows, needles, threshold, cap=int(n))
        return out[:n].copy()

    def matcher_scan(self, rows, needles, threshold: int) -> bool:
        rows = np.ascontiguousarray(rows, dtype=np.uint8).reshape(-1, 32)
        needles = np.ascontiguousarray(needles, dtype=np.uint8).reshape(-1, 32)
        return bool(self.lib.oracle_matcher_scan(_bp(rows), rows.shape[0], _bp(needles),
                                                 needles.shape[0], threshold))

    def matcher_find_min(self, rows, needles, threshold: int):
        rows = np.ascontiguousarray(rows, dtype=np.uint8).reshape(-1, 32)
        needles = np.ascontiguousarray(needles, dtype=np.uint8).reshape(-1, 32)
        v = self.lib.oracle_matcher_find_min(_bp(rows), rows.shape[0], _bp(needles),
                                             needles.shape[0], threshold)
        return None if v == 0xFFFFFFFFFFFFFFFF else (int(v) // needles.shape[0],
                                                     int(v) % needles.shape[0])

    # ---- synthetic workloads --------------------------------------------
    def synth_db(self, n_rows: int, first_row: int = 0, seed: int = 0x79756D652D706471):
        rows = np.zeros((n_rows, 32), dtype=np.uint8)
        self.lib.oracle_synth_db(_bp(rows), n_rows, first_row, seed)
        return rows

    def synth_frames(self, n: int, first_frame: int = 0, seed: int = 0x79756D652D706471,
                     kind: int = 1):
        frames = np.zeros((n, 512, 512), dtype=np.float32)
        self.lib.oracle_synth_frames(_fp(frames), n, first_frame, seed, kind)
        return frames


    def synth_pixels(self, n: int, channels: int, first_frame: int = 0, seed: int = 0x79756D652D706471):
        shape = (n, 512, 512) if channels == 1 else (n, 512, 512, channels)
        frames = np.zeros(shape, dtype=np.uint8)
        self.lib.oracle_synth_pixels(_bp(frames), n, first_frame, seed, channels)
        return frames


_ORACLE = None


def get() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE


def readme_doctest_frame() -> np.ndarray:
    """The README doctest input, README.md:313-322, in strict f32.

    `frequency` starts at 4.0 and is incremented by 4/512 BEFORE every element,
    carried across rows; value = sin(j*2*pi*frequency/512)*127+128, all f32.
    """
    f32 = np.float32
    modulation = f32(4.0) / f32(512.0)
    inc = np.full(512 * 512, modulation, dtype=np.float32)
    inc[0] = f32(4.0) + modulation
    freq = np.add.accumulate(inc, dtype=np.float32).reshape(512, 512)  # sequential f32 adds
    j = np.arange(512, dtype=np.float32)[None, :]
    two_pi = f32(2.0) * f32(np.pi)
    arg = ((j * f32(2.0)) * f32(np.pi) * freq) / f32(512.0)
    del two_pi
    return (np.sin(arg.astype(np.float32)).astype(np.float32) * f32(127.0) + f32(128.0)).astype(
        np.float32)


class CpuBaseline:
    """ctypes view of oracle/_build/libcpu_baseline.so (AVX2 hasher + AVX-512 matcher port)."""

    def __init__(self, path: str | None = None):
        path = path or os.path.join(BUILD, "libcpu_baseline.so")
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        L.cpu_baseline_isa.restype = C.c_int
        L.cpu_baseline_hw_threads.restype = C.c_int
        L.cpu_baseline_hash_batch.argtypes = [_f32p, C.c_size_t, _u8p, _f32p, _f32p, _f32p, C.c_int]
        L.cpu_baseline_hash_batch.restype = C.c_int
        L.cpu_baseline_match.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.c_uint32, C.c_int,
                                         C.POINTER(Match), C.c_size_t, C.c_int, C.c_int]
        L.cpu_baseline_match.restype = C.c_size_t
        L.cpu_baseline_quantize_avx2.argtypes = [_f32p, _f32p, _u8p]
        L.cpu_baseline_quantize_avx2.restype = C.c_int
        self.isa = L.cpu_baseline_isa()
        self.hw_threads = L.cpu_baseline_hw_threads()

    @property
    def has_avx2(self):
        return bool(self.isa & 1)

    @property
    def has_avx512_popcnt(self):
        return bool(self.isa & 2)

    def hash_batch(self, frames, threads: int = 1, want_pdqf: bool = True):
        frames = np.ascontiguousarray(frames, dtype=np.float32).reshape(-1, 512, 512)
        n = frames.shape[0]
        hashes = np.zeros((n, 8, 32), dtype=np.uint8)
        quality = np.zeros((n, 8), dtype=np.float32)
        thr = np.zeros((n, 8), dtype=np.float32)
        pdqf = np.zeros((n, 16, 16), dtype=np.float32) if want_pdqf else None
        rc = self.lib.cpu_baseline_hash_batch(_fp(frames), n, _bp(hashes), _fp(quality), _fp(thr),
                                              _fp(pdqf) if want_pdqf else None, threads)
        if rc != 0:
            raise RuntimeError("host CPU has no AVX2+FMA")
        return hashes, quality, thr, pdqf

    def quantize_avx2(self, pdqf):
        """Avx2F32Kernel::quantize (x86.rs:440-621): (hash [32] u8, bisection threshold)."""
        pdqf = np.ascontiguousarray(pdqf, dtype=np.float32).reshape(256)
        thr = np.zeros(1, dtype=np.float32)
        out = np.zeros(32, dtype=np.uint8)
        if self.lib.cpu_baseline_quantize_avx2(_fp(pdqf), _fp(thr), _bp(out)) != 0:
            raise RuntimeError("host CPU has no AVX2+FMA")
        return out, np.float32(thr[0])

    def match(self, rows, needles, threshold: int, find: bool = True, threads: int = 1,
              force_scalar: bool = False, cap: int = 1 << 20):
        rows = np.ascontiguousarray(rows, dtype=np.uint8).reshape(-1, 32)
        needles = np.ascontiguousarray(needles, dtype=np.uint8).reshape(-1, 32)
        out = np.zeros(cap if find else 1, dtype=MATCH_DTYPE)
        n = self.lib.cpu_baseline_match(_bp(rows), rows.shape[0], _bp(needles), needles.shape[0],
                                        threshold, int(find), out.ctypes.data_as(C.POINTER(Match)),
                                        cap if find else 0, threads, int(force_scalar))
        if not find:
            return bool(n)
        if n > cap:
            return self.match(rows, needles, threshold, True, threads, force_scalar, cap=int(n))
        return out[:n].copy()
