"""yume_pdq_b200 -- host-side mirror of the reference's plugin surface over the C ABI.

This package is a thin ctypes binding of ``libyume_pdq_cuda.so``
(include/yume_pdq_cuda.h).  Names follow the reference crate:

  ``CudaKernel``          the ``Kernel`` trait          (src/kernel/mod.rs:268-497)
  ``hash`` / ``hash_get_threshold`` / ``hash_float`` / ``visit_dihedrals``
                          the drivers                   (src/lib.rs:322-479)
  ``Dihedrals``           dihedral constants            (src/lib.rs:85-134)
  ``CudaVectorDatabase``  like ``VulkanVectorDatabase`` (src/matching/vulkan.rs:40-65)
  ``CudaMatcher``         the ``PDQMatcher`` trait      (src/matching/mod.rs:50-105)

There is NO CPU fallback: importing works anywhere the shared library exists
(so the symbol table can be checked), but every compute call raises
``YumePdqCudaError`` unless a B200-class (sm_100) device is present.  The CPU
oracle under ``oracle/`` is test infrastructure and is never imported here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# YUME_PDQ_CUDA_LIB: developer knob, an experimental build of the same library (yume_pdq_b200/build.py variant)
LIB_PATH = os.environ.get("YUME_PDQ_CUDA_LIB") or os.path.join(HERE, "libyume_pdq_cuda.so")

MATCH_DTYPE = np.dtype([("row", "<u8"), ("needle", "<u4"), ("dist", "<u4")])

ALGO_AUTO, ALGO_EXACT, ALGO_TILED, ALGO_TENSOR, ALGO_TENSOR_I8 = 0, 1, 2, 4, 5
ALGO_TENSOR_FULL, ALGO_TENSOR_PREFIX, ALGO_TENSOR_PAIRED, ALGO_TENSOR_ROWS = 6, 7, 8, 9
MAILBOX_AUTO, MAILBOX_SEPARATE, MAILBOX_FUSED = 0, 1, 2
EXCHANGE_KERNEL, EXCHANGE_COPY_ENGINE = 0, 1

OP_T = 0
OP_NEGATE_ALT_COLS_FALSE = 1
OP_NEGATE_ALT_COLS_TRUE = 2
OP_NEGATE_ALT_ROWS_TRUE = 3
OP_NEGATE_ALT_ROWS_FALSE = 4
OP_NEGATE_OFF_DIAGONALS = 5

PIXEL_F32, PIXEL_LUMA8, PIXEL_RGB8, PIXEL_RGBA8, PIXEL_FMA = 0, 1, 2, 3, 0x100
# RGB8_TO_LUMA8_TABLE_ITU, src/kernel/constants.rs:23
RGB8_TO_LUMA8_TABLE_ITU = (0.299, 0.587, 0.114)

E_INVALID, E_NO_DEVICE, E_CAPACITY, E_NOMEM = -1, -2, -3, -4
COUNT_SATURATED = 0xF0000000    # the device-side match counter stops here (csrc/common.cuh)


class YumePdqCudaError(RuntimeError):
    def __init__(self, code: int, what: str):
        self.code = code
        super().__init__(f"{what}: {error_string(code)} (code {code})")


class Match(C.Structure):
    _fields_ = [("row", C.c_uint64), ("needle", C.c_uint32), ("dist", C.c_uint32)]


DihedralCallback = C.CFUNCTYPE(C.c_bool, C.c_void_p, C.c_uint32, C.c_float, C.c_float)

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m yume_pdq_b200.build` "
        "(nvcc, sm_100a).  yume_pdq_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

# every symbol include/yume_pdq_cuda.h declares: (restype, argtypes)
SIGNATURES = {
    "yume_pdq_hash_smart_kernel": (C.c_float, [_vp, _f32p, _vp, _vp, _vp, _vp]),
    "yume_pdq_visit_dihedrals_smart_kernel": (C.c_bool, [_vp, C.c_float, _vp, _vp, DihedralCallback]),
    "yume_pdq_cuda_error_string": (C.c_char_p, [C.c_int]),
    "yume_pdq_cuda_device_count": (C.c_int, []),
    "yume_pdq_cuda_kernel_ident": (C.c_char_p, []),
    "yume_pdq_cuda_host_alloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    "yume_pdq_cuda_host_free": (C.c_int, [C.c_void_p]),
    "yume_pdq_cuda_matcher_ident": (C.c_char_p, []),
    "yume_pdq_cuda_launch_count": (C.c_uint64, []),
    "yume_pdq_cuda_hash_batch": (C.c_int, [C.c_int, _vp, C.c_size_t, _vp, _vp, _vp, _vp, _vp]),
    "yume_pdq_cuda_frame_bytes": (C.c_size_t, [C.c_int]),
    "yume_pdq_cuda_hash_batch_pixels": (C.c_int, [C.c_int, _vp, C.c_int, C.c_size_t, _vp, _vp, _vp, _vp, _vp]),
    "yume_pdq_cuda_cvt_to_luma8f": (C.c_int, [C.c_int, _vp, C.c_int, C.c_size_t, _vp, _vp, _vp]),
    "yume_pdq_cuda_jarosz_compress": (C.c_int, [C.c_int, _vp, C.c_size_t, _vp, _vp]),
    "yume_pdq_cuda_dct2d": (C.c_int, [C.c_int, _vp, C.c_size_t, _vp, _vp]),
    "yume_pdq_cuda_sum_of_gradients": (C.c_int, [C.c_int, _vp, C.c_size_t, _vp, _vp, _vp]),
    "yume_pdq_cuda_quantize": (C.c_int, [C.c_int, _vp, C.c_size_t, _vp, _vp, _vp]),
    "yume_pdq_cuda_threshold_2d_f32": (C.c_int, [C.c_int, _vp, C.c_size_t, _vp, _vp, _vp]),
    "yume_pdq_cuda_pdqf_transform": (C.c_int, [C.c_int, _vp, C.c_size_t, C.c_int, _vp]),
    "yume_pdq_cuda_dihedrals_batch": (C.c_int, [C.c_int, _vp, _vp, C.c_size_t, _vp, _vp, _vp, _vp]),
    "yume_pdq_cuda_db_create": (C.c_int, [C.c_int, _vp, C.c_size_t, C.c_uint64, C.POINTER(_vp)]),
    "yume_pdq_cuda_db_create_synthetic": (C.c_int, [C.c_int, C.c_size_t, C.c_uint64, C.c_uint64,
                                                    C.POINTER(_vp)]),
    "yume_pdq_cuda_db_destroy": (C.c_int, [_vp]),
    "yume_pdq_cuda_db_len": (C.c_size_t, [_vp]),
    "yume_pdq_cuda_db_device": (C.c_int, [_vp]),
    "yume_pdq_cuda_db_planes": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "yume_pdq_cuda_mailbox_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_vp), _vp]),
    "yume_pdq_cuda_mailbox_connect": (C.c_int, [_vp, _vp]),
    "yume_pdq_cuda_mailbox_destroy": (C.c_int, [_vp]),
    "yume_pdq_cuda_mailbox_last_timing": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "yume_pdq_cuda_mailbox_set_mode": (C.c_int, [_vp, C.c_int]),
    "yume_pdq_cuda_exchange_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(_vp), _vp]),
    "yume_pdq_cuda_exchange_connect": (C.c_int, [_vp, C.c_char_p]),
    "yume_pdq_cuda_exchange_post": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "yume_pdq_cuda_exchange_post_bytes": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, _vp]),
    "yume_pdq_cuda_exchange_wait": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "yume_pdq_cuda_exchange_failed": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "yume_pdq_cuda_exchange_destroy": (C.c_int, [_vp]),
    "yume_pdq_cuda_one_frame_last_timing": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "yume_pdq_cuda_mailbox_scan": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_uint32, C.c_int, C.POINTER(C.c_uint32), _vp]),
    "yume_pdq_cuda_db_auto_state": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_uint32)]),
    "yume_pdq_cuda_db_write_rows": (C.c_int, [_vp, C.c_size_t, _vp, C.c_size_t]),
    "yume_pdq_cuda_matcher_scan": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32, C.POINTER(C.c_int), _vp]),
    "yume_pdq_cuda_matcher_find": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32, _vp, C.c_size_t,
                                             C.POINTER(C.c_size_t), _vp]),
    "yume_pdq_cuda_matcher_find_min": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32,
                                                 C.POINTER(C.c_uint64), _vp]),
    "yume_pdq_cuda_matcher_find_device": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32, _vp, C.c_uint32,
                                                    _vp, C.c_int, _vp]),
    "yume_pdq_cuda_matcher_find_device_base": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32, _vp, C.c_uint32,
                                                         _vp, C.c_int, C.c_uint32, _vp]),
    "yume_pdq_cuda_matcher_find_device_split": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32, _vp, C.c_uint32,
                                                          _vp, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "yume_pdq_cuda_screen_batch": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32, _vp, _vp, _vp, C.c_size_t,
                                             C.POINTER(C.c_size_t), _vp]),
    "yume_pdq_cuda_screen_batch_pixels": (C.c_int, [_vp, _vp, C.c_int, C.c_size_t, C.c_uint32, _vp, _vp, _vp,
                                                    C.c_size_t, C.POINTER(C.c_size_t), _vp]),
    "yume_pdq_cuda_measure_mxf4_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "yume_pdq_cuda_synth_pixels": (C.c_int, [C.c_int, _vp, C.c_int, C.c_size_t, C.c_uint64, C.c_uint64, _vp]),
    "yume_pdq_cuda_synth_frames": (C.c_int, [C.c_int, _vp, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int, _vp]),
    "yume_pdq_cuda_synth_rows": (C.c_int, [C.c_int, _vp, C.c_size_t, C.c_uint64, C.c_uint64, _vp]),
}
DATA_SYMBOLS = [
    "YUME_PDQ_VERSION_MAJOR", "YUME_PDQ_VERSION_MINOR", "YUME_PDQ_VERSION_PATCH",
    "YUME_PDQ_VERSION_STRING", "YUME_PDQ_DIHEDRAL_NORMAL", "YUME_PDQ_DIHEDRAL_FLIPPED",
    "YUME_PDQ_DIHEDRAL_FLOPPED", "YUME_PDQ_DIHEDRAL_ROTATED_180", "YUME_PDQ_DIHEDRAL_ROTATED_90",
    "YUME_PDQ_DIHEDRAL_ROTATED_270", "YUME_PDQ_DIHEDRAL_ROTATED_90_FLOPPED",
    "YUME_PDQ_DIHEDRAL_FLOPPED_ROTATED_270", "YUME_PDQ_CUDA_DIHEDRAL_ORDER",
]
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def error_string(code: int) -> str:
    return lib.yume_pdq_cuda_error_string(code).decode()


def _check(code: int, what: str) -> None:
    if code != 0:
        raise YumePdqCudaError(code, what)


def device_count() -> int:
    return lib.yume_pdq_cuda_device_count()


def launch_count() -> int:
    return int(lib.yume_pdq_cuda_launch_count())


def version() -> str:
    return C.string_at(C.addressof(C.c_char.in_dll(lib, "YUME_PDQ_VERSION_STRING"))).decode()


def _ptr(x):
    """numpy array -> host pointer, torch CUDA/CPU tensor -> data_ptr, int -> as is, None -> NULL."""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return C.c_void_p(x.data_ptr())
    raise TypeError(type(x))


def _stream(stream):
    if stream is None:
        return None
    if isinstance(stream, int):
        return C.c_void_p(stream)
    return C.c_void_p(stream.cuda_stream)  # torch.cuda.Stream


class Dihedrals:
    """src/lib.rs:85-134."""

    @staticmethod
    def from_tuples(dx, dy) -> int:
        return int.from_bytes(bytes([dx[0] & 0xFF, dx[1] & 0xFF, dy[0] & 0xFF, dy[1] & 0xFF]), "big")

    @staticmethod
    def into_tuples(packed: int):
        b = packed.to_bytes(4, "big")
        s = lambda v: v - 256 if v > 127 else v
        return (s(b[0]), s(b[1])), (s(b[2]), s(b[3]))

    NORMAL = C.c_uint32.in_dll(lib, "YUME_PDQ_DIHEDRAL_NORMAL").value
    FLIPPED = C.c_uint32.in_dll(lib, "YUME_PDQ_DIHEDRAL_FLIPPED").value
    FLOPPED = C.c_uint32.in_dll(lib, "YUME_PDQ_DIHEDRAL_FLOPPED").value
    ROTATED_180 = C.c_uint32.in_dll(lib, "YUME_PDQ_DIHEDRAL_ROTATED_180").value
    ROTATED_90 = C.c_uint32.in_dll(lib, "YUME_PDQ_DIHEDRAL_ROTATED_90").value
    ROTATED_270 = C.c_uint32.in_dll(lib, "YUME_PDQ_DIHEDRAL_ROTATED_270").value
    ROTATED_90_FLOPPED = C.c_uint32.in_dll(lib, "YUME_PDQ_DIHEDRAL_ROTATED_90_FLOPPED").value
    FLOPPED_ROTATED_270 = C.c_uint32.in_dll(lib, "YUME_PDQ_DIHEDRAL_FLOPPED_ROTATED_270").value


DIHEDRAL_ORDER = list((C.c_uint32 * 8).in_dll(lib, "YUME_PDQ_CUDA_DIHEDRAL_ORDER"))


# ----------------------------------------------------------------------------
# Kernel trait
class CudaKernel:
    """Batched CUDA implementation of the ``Kernel`` trait (src/kernel/mod.rs:268-497).

    Every method takes arrays with a leading batch dimension; numpy arrays are
    staged through the device, torch CUDA tensors are used in place
    (asynchronously on ``stream``).
    """

    InputDimension = 512
    OutputDimension = 16
    Buffer1WidthX = 128
    Buffer1LengthY = 128
    InternalFloat = np.float32
    RequiredHardwareFeature = "CudaAccelerator(sm_100)"

    def __init__(self, device: int = 0, stream=None):
        self.device = device
        self.stream = stream

    def ident(self) -> str:
        return lib.yume_pdq_cuda_kernel_ident().decode()

    @staticmethod
    def adjust_quality(gradient):
        """Kernel::adjust_quality (mod.rs:611-615) is a pure scalar map; it is applied on the
        device by sum_of_gradients(..., quality=True); this helper exists for API parity."""
        return np.minimum(np.float32(gradient) / np.float32(180.0), np.float32(1.0))

    def _out(self, like, shape, dtype):
        if isinstance(like, np.ndarray):
            return np.empty(shape, dtype=dtype)
        import torch
        return torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), device=like.device)

    # input conversion (mod.rs:303-340); rows [..., 3] / [..., 4] u8 -> f32 luma of the leading shape
    def _cvt(self, pixels, fmt, channels, coeff, fma):
        if channels > 1 and pixels.shape[-1] != channels:
            raise ValueError(f"last dimension must be {channels}")
        shape = tuple(pixels.shape[:-1]) if channels > 1 else tuple(pixels.shape)
        n_pixels = int(np.prod(shape)) if shape else 1
        out = self._out(pixels, shape, np.float32)
        cf = None if coeff is None else (C.c_float * 3)(*coeff)
        _check(lib.yume_pdq_cuda_cvt_to_luma8f(self.device, _ptr(pixels), fmt | (PIXEL_FMA if fma else 0),
                                               n_pixels, cf, _ptr(out), _stream(self.stream)), "cvt_to_luma8f")
        return out

    def cvt_rgb8_to_luma8f(self, rows, coeff=None, fma=False):
        """Kernel::cvt_rgb8_to_luma8f (mod.rs:303-321): u8 [..., 3] -> f32 [...]; coeff = (R, G, B) like the
        trait's const generics (default: the ITU table); fma=True: AVX2 rounding (x86.rs:686-753)."""
        return self._cvt(rows, PIXEL_RGB8, 3, coeff, fma)

    def cvt_rgba8_to_luma8f(self, rows, coeff=None, fma=False):
        """Kernel::cvt_rgba8_to_luma8f (mod.rs:324-340): u8 [..., 4] -> f32 [...], alpha ignored."""
        return self._cvt(rows, PIXEL_RGBA8, 4, coeff, fma)

    def cvt_luma8_to_luma8f(self, rows):
        """`row_buf[j] as f32` of the CLI's LUMA8 input (main.rs:508-510)."""
        return self._cvt(rows, PIXEL_LUMA8, 1, None, False)

    def jarosz_compress(self, frames, out=None):
        n = frames.shape[0]
        out = self._out(frames, (n, 128, 128), np.float32) if out is None else out
        _check(lib.yume_pdq_cuda_jarosz_compress(self.device, _ptr(frames), n, _ptr(out),
                                                 _stream(self.stream)), "jarosz_compress")
        return out

    def dct2d(self, buf1, out=None):
        n = buf1.shape[0]
        out = self._out(buf1, (n, 16, 16), np.float32) if out is None else out
        _check(lib.yume_pdq_cuda_dct2d(self.device, _ptr(buf1), n, _ptr(out), _stream(self.stream)),
               "dct2d")
        return out

    def sum_of_gradients(self, pdqf, quality: bool = False):
        n = pdqf.shape[0]
        g = self._out(pdqf, (n,), np.float32)
        q = self._out(pdqf, (n,), np.float32) if quality else None
        _check(lib.yume_pdq_cuda_sum_of_gradients(self.device, _ptr(pdqf), n, _ptr(g), _ptr(q),
                                                  _stream(self.stream)), "sum_of_gradients")
        return (g, q) if quality else g

    def quantize(self, pdqf):
        n = pdqf.shape[0]
        thr = self._out(pdqf, (n,), np.float32)
        out = self._out(pdqf, (n, 32), np.uint8)
        _check(lib.yume_pdq_cuda_quantize(self.device, _ptr(pdqf), n, _ptr(thr), _ptr(out),
                                          _stream(self.stream)), "quantize")
        return out, thr

    def threshold_2d_f32(self, pdqf, thresholds):
        n = pdqf.shape[0]
        out = self._out(pdqf, (n, 32), np.uint8)
        _check(lib.yume_pdq_cuda_threshold_2d_f32(self.device, _ptr(pdqf), n, _ptr(thresholds),
                                                  _ptr(out), _stream(self.stream)), "threshold_2d_f32")
        return out

    def _transform(self, pdqf, op):
        _check(lib.yume_pdq_cuda_pdqf_transform(self.device, _ptr(pdqf), pdqf.shape[0], op,
                                                _stream(self.stream)), "pdqf_transform")
        return pdqf

    def pdqf_t(self, pdqf):
        return self._transform(pdqf, OP_T)

    def pdqf_negate_alt_cols(self, pdqf, negate: bool):
        return self._transform(pdqf, OP_NEGATE_ALT_COLS_TRUE if negate else OP_NEGATE_ALT_COLS_FALSE)

    def pdqf_negate_alt_rows(self, pdqf, negate: bool):
        return self._transform(pdqf, OP_NEGATE_ALT_ROWS_TRUE if negate else OP_NEGATE_ALT_ROWS_FALSE)

    def pdqf_negate_off_diagonals(self, pdqf):
        return self._transform(pdqf, OP_NEGATE_OFF_DIAGONALS)

    # fused hot path
    def hash_batch(self, frames, want_quality=True, want_thresholds=True, want_pdqf=False, out=None,
                   pixel_format: int = PIXEL_F32):
        """frames [n,512,512] f32 (or u8 [n,512,512] / [n,512,512,3] / [n,512,512,4] with the matching
        ``pixel_format``) -> dict(hashes [n,8,32] u8, quality [n,8], thresholds [n,8], pdqf [n,16,16])
        in DIHEDRAL_ORDER."""
        n = frames.shape[0]
        _check_frames(frames, pixel_format)
        out = out or {}
        hashes = out.get("hashes")
        if hashes is None:
            hashes = self._out(frames, (n, 8, 32), np.uint8)
        quality = out.get("quality")
        if quality is None and want_quality:
            quality = self._out(frames, (n, 8), np.float32)
        thr = out.get("thresholds")
        if thr is None and want_thresholds:
            thr = self._out(frames, (n, 8), np.float32)
        pdqf = out.get("pdqf")
        if pdqf is None and want_pdqf:
            pdqf = self._out(frames, (n, 16, 16), np.float32)
        _check(lib.yume_pdq_cuda_hash_batch_pixels(self.device, _ptr(frames), pixel_format, n, _ptr(hashes),
                                                   _ptr(quality), _ptr(thr), _ptr(pdqf), _stream(self.stream)),
               "hash_batch")
        return {"hashes": hashes, "quality": quality, "thresholds": thr, "pdqf": pdqf}

    def dihedrals_batch(self, pdqf, thr0):
        n = pdqf.shape[0]
        hashes = self._out(pdqf, (n, 8, 32), np.uint8)
        quality = self._out(pdqf, (n, 8), np.float32)
        thr = self._out(pdqf, (n, 8), np.float32)
        _check(lib.yume_pdq_cuda_dihedrals_batch(self.device, _ptr(pdqf), _ptr(thr0), n, _ptr(hashes),
                                                 _ptr(quality), _ptr(thr), _stream(self.stream)),
               "dihedrals_batch")
        return {"hashes": hashes, "quality": quality, "thresholds": thr}


def _check_frames(frames, pixel_format: int) -> None:
    """The byte size of ``frames`` must be n frames of ``pixel_format`` (a wrong format would read out of bounds)."""
    fb = int(lib.yume_pdq_cuda_frame_bytes(pixel_format))
    if fb == 0:
        raise ValueError(f"unknown pixel format {pixel_format:#x}")
    nbytes = frames.nbytes if isinstance(frames, np.ndarray) else frames.numel() * frames.element_size()
    if nbytes != int(frames.shape[0]) * fb:
        raise ValueError(f"frames hold {nbytes} bytes, expected {frames.shape[0]} x {fb} for pixel format "
                         f"{pixel_format:#x}")


def smart_kernel(device: int = 0) -> CudaKernel:
    """kernel::smart_kernel (mod.rs:96-103): here always the CUDA kernel."""
    return CudaKernel(device)


# ---- drivers over the reference's own two-function C ABI (single frame) ----
def hash_get_threshold(frame: np.ndarray):
    """lib.rs:427-455 through ``yume_pdq_hash_smart_kernel`` -> (quality, threshold, hash, pdqf, buf1).  `frame`: a
    numpy array (pageable host memory: copied through the library's pinned block while the kernel already runs) or a
    contiguous f32 torch tensor -- pinned host memory is read in place over PCIe, device memory as is."""
    if isinstance(frame, np.ndarray) or not hasattr(frame, "data_ptr"):
        frame = np.ascontiguousarray(frame, dtype=np.float32).reshape(512, 512)
    elif frame.numel() != 512 * 512 or frame.element_size() != 4:
        raise ValueError("frame must hold 512 x 512 f32")
    thr = C.c_float(0.0)
    out = np.zeros(32, dtype=np.uint8)
    buf1 = np.zeros((128, 128), dtype=np.float32)
    tmp = np.zeros(128, dtype=np.float32)
    pdqf = np.zeros((16, 16), dtype=np.float32)
    q = lib.yume_pdq_hash_smart_kernel(_ptr(frame), C.byref(thr), _ptr(out), _ptr(buf1), _ptr(tmp),
                                       _ptr(pdqf))
    if q != q:
        raise YumePdqCudaError(E_NO_DEVICE, "yume_pdq_hash_smart_kernel")
    return float(q), float(thr.value), out, pdqf, buf1


def hash(frame: np.ndarray):  # noqa: A001 - mirrors yume_pdq::hash (lib.rs:322-351)
    q, _thr, out, pdqf, _ = hash_get_threshold(frame)
    return q, out, pdqf


def hash_float(frame: np.ndarray):
    """lib.rs:464-479: PDQF and quality without quantisation."""
    q, _thr, _out, pdqf, _ = hash_get_threshold(frame)
    return q, pdqf


def visit_dihedrals(pdqf: np.ndarray, threshold: float, f):
    """lib.rs:358-420 through ``yume_pdq_visit_dihedrals_smart_kernel``.

    ``f(dihedral, threshold, quality, pdqf_view, hash_view) -> bool`` is called 7 times; pdqf is
    transformed in place exactly like the reference.  Returns True iff all were visited."""
    if pdqf.dtype != np.float32 or not pdqf.flags["C_CONTIGUOUS"]:
        raise ValueError("pdqf must be a C-contiguous float32 array (it is modified in place)")
    out = np.zeros(32, dtype=np.uint8)

    def cb(_ctx, dihedral, thr, quality):
        return bool(f(int(dihedral), float(thr), float(quality), pdqf, out))

    return bool(lib.yume_pdq_visit_dihedrals_smart_kernel(None, C.c_float(threshold), _ptr(out),
                                                          _ptr(pdqf), DihedralCallback(cb)))


# ----------------------------------------------------------------------------
# matcher
class CudaVectorDatabase:
    """Haystack resident in HBM (mirrors VulkanVectorDatabase::new, vulkan.rs:50-64)."""

    def __init__(self, rows=None, device: int = 0, global_row_offset: int = 0, *,
                 synthetic_rows: int | None = None, seed: int = 0x79756D652D706471):
        self._h = _vp()
        if synthetic_rows is not None:
            _check(lib.yume_pdq_cuda_db_create_synthetic(device, synthetic_rows, global_row_offset, seed,
                                                         C.byref(self._h)), "db_create_synthetic")
        else:
            if isinstance(rows, np.ndarray):
                rows = np.ascontiguousarray(rows, dtype=np.uint8).reshape(-1, 32)
            n = rows.shape[0]
            _check(lib.yume_pdq_cuda_db_create(device, _ptr(rows), n, global_row_offset,
                                               C.byref(self._h)), "db_create")
        self.device = device
        self.global_row_offset = global_row_offset

    def __len__(self):
        return int(lib.yume_pdq_cuda_db_len(self._h))

    @property
    def planes_ptr(self) -> tuple[int, int]:
        """Device addresses of the prefix plane (bytes 0..15 of every row) and the suffix plane (bytes 16..31)."""
        pre, suf = _vp(), _vp()
        _check(lib.yume_pdq_cuda_db_planes(self._h, C.byref(pre), C.byref(suf)), "db_planes")
        return int(pre.value or 0), int(suf.value or 0)

    @property
    def auto_state(self) -> tuple[bool, int]:
        """(fallback, probe_interval) of ALGO_AUTO on this haystack: fallback is True while large needle batches run the
        unfiltered tensor kernel because the prefix filter met too many candidates."""
        fb, iv = C.c_int(0), C.c_uint32(0)
        _check(lib.yume_pdq_cuda_db_auto_state(self._h, C.byref(fb), C.byref(iv)), "db_auto_state")
        return bool(fb.value), int(iv.value)

    def write_rows(self, first_local_row: int, rows):
        if isinstance(rows, np.ndarray):
            rows = np.ascontiguousarray(rows, dtype=np.uint8).reshape(-1, 32)
        _check(lib.yume_pdq_cuda_db_write_rows(self._h, first_local_row, _ptr(rows), rows.shape[0]),
               "db_write_rows")

    def close(self):
        if self._h:
            lib.yume_pdq_cuda_db_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CudaMatcher:
    """``PDQMatcher`` for the GPU (matching/mod.rs:50-105): the haystack lives on the device and a
    query is a set of needles, like ``VulkanMatcher`` (vulkan.rs:335-472) -- but exact AND complete."""

    InputDimension = 32
    NoFalseNegative = True
    NoFalsePositive = True
    FindsAllMatches = True

    def __init__(self, haystack: CudaVectorDatabase, threshold: int, stream=None):
        if not 0 <= threshold <= 256:
            raise ValueError("threshold must be in [0, 256]")
        self.haystack = haystack
        self.threshold = threshold
        self.stream = stream

    @staticmethod
    def ident() -> str:
        return lib.yume_pdq_cuda_matcher_ident().decode()

    @staticmethod
    def _needles(query):
        if isinstance(query, np.ndarray):
            query = np.ascontiguousarray(query, dtype=np.uint8).reshape(-1, 32)
        return query, query.shape[0]

    def scan(self, query) -> bool:
        query, q = self._needles(query)
        m = C.c_int(0)
        _check(lib.yume_pdq_cuda_matcher_scan(self.haystack._h, _ptr(query), q, self.threshold,
                                              C.byref(m), _stream(self.stream)), "matcher_scan")
        return bool(m.value)

    def find_all(self, query, cap: int = 1 << 16) -> np.ndarray:
        """All matches as a MATCH_DTYPE array sorted by (row, needle)."""
        query, q = self._needles(query)
        while True:
            out = np.zeros(cap, dtype=MATCH_DTYPE)
            n = C.c_size_t(0)
            rc = lib.yume_pdq_cuda_matcher_find(self.haystack._h, _ptr(query), q, self.threshold,
                                                _ptr(out), cap, C.byref(n), _stream(self.stream))
            if rc == E_CAPACITY:
                if n.value >= COUNT_SATURATED:
                    raise YumePdqCudaError(rc, "matcher_find (more than 4e9 matches: the 32-bit device counter saturated)")
                cap = int(n.value)
                continue
            _check(rc, "matcher_find")
            return out[: n.value].copy()

    def find(self, query, f):
        """PDQMatcher::find: call ``f(hay_idx, needle_idx)`` for every match in (row, needle) order
        until it returns something other than None; return that value (or None)."""
        for m in self.find_all(query):
            r = f(int(m["row"]), int(m["needle"]))
            if r is not None:
                return r
        return None

    def find_min(self, query):
        """VulkanMatcher::find semantics: (local_row, needle) of the minimum needle + row*q, or None."""
        query, q = self._needles(query)
        k = C.c_uint64(0)
        _check(lib.yume_pdq_cuda_matcher_find_min(self.haystack._h, _ptr(query), q, self.threshold,
                                                  C.byref(k), _stream(self.stream)), "matcher_find_min")
        if k.value == 0xFFFFFFFFFFFFFFFF:
            return None
        return int(k.value) // q, int(k.value) % q

    def screen(self, frames, hashes=None, quality=None, cap: int = 65536, pixel_format: int = PIXEL_F32):
        """Hash ``frames`` (n x 512 x 512 f32 -- or 8-bit frames with ``pixel_format`` --, numpy or torch CUDA) and match all 8 n dihedral hashes
        against the haystack in one pipelined call (examples/end_to_end.rs:224-256 for a batch).
        Returns the matches (needle = 8 * frame + dihedral) sorted by (row, needle); fills ``hashes``
        (n x 8 x 32) / ``quality`` (n x 8) when given."""
        n = int(frames.shape[0])
        _check_frames(frames, pixel_format)
        while True:
            out = np.zeros(max(cap, 1), dtype=MATCH_DTYPE)
            k = C.c_size_t(0)
            rc = lib.yume_pdq_cuda_screen_batch_pixels(self.haystack._h, _ptr(frames), pixel_format, n,
                                                       self.threshold, _ptr(hashes), _ptr(quality), _ptr(out),
                                                       cap, C.byref(k), _stream(self.stream))
            if rc == E_CAPACITY:
                if k.value >= COUNT_SATURATED:
                    raise YumePdqCudaError(rc, "screen_batch (more than 4e9 matches: the 32-bit device counter saturated)")
                cap = int(k.value)
                continue
            _check(rc, "screen_batch")
            return out[: k.value].copy()

    def find_device(self, needles_dev, q: int, out_dev, cap: int, count_dev, algo: int = ALGO_AUTO,
                    stream=None, needle_base: int = 0, split_at=None, gap: int = 0):
        """Asynchronous device-side scan (torch CUDA tensors or raw device pointers); ``needle_base`` is
        added to the reported needle indices (chunked callers); with ``split_at`` the batch is the concatenation of two
        ranges of a longer list and needles from ``split_at`` on are reported ``gap`` higher
        (yume_pdq_cuda_matcher_find_device_split)."""
        if split_at is not None:
            _check(lib.yume_pdq_cuda_matcher_find_device_split(self.haystack._h, _ptr(needles_dev), q,
                                                               self.threshold, _ptr(out_dev), cap, _ptr(count_dev),
                                                               algo, needle_base, split_at, gap,
                                                               _stream(stream or self.stream)),
                   "matcher_find_device_split")
            return
        _check(lib.yume_pdq_cuda_matcher_find_device_base(self.haystack._h, _ptr(needles_dev), q,
                                                          self.threshold, _ptr(out_dev), cap, _ptr(count_dev),
                                                          algo, needle_base, _stream(stream or self.stream)),
               "matcher_find_device")


IPC_HANDLE_BYTES = 64


class PeerMailbox:
    """This rank's end of the peer-memory query (yume_pdq_cuda_mailbox_*): `handle` goes to the peers, `connect` takes
    everyone's handles in rank order, `scan` is the collective query.  yume_pdq_b200.sharded wires it to
    torch.distributed."""

    def __init__(self, device: int, rank: int, world: int):
        self._h = _vp()
        buf = (C.c_uint8 * IPC_HANDLE_BYTES)()
        _check(lib.yume_pdq_cuda_mailbox_create(device, rank, world, C.byref(self._h), buf), "mailbox_create")
        self.handle = bytes(buf)
        self.rank, self.world = rank, world

    def connect(self, handles):
        blob = b"".join(handles)
        assert len(blob) == self.world * IPC_HANDLE_BYTES
        _check(lib.yume_pdq_cuda_mailbox_connect(self._h, C.c_char_p(blob)), "mailbox_connect")

    def scan(self, db: "CudaVectorDatabase", needles, q: int, threshold: int, root: int = 0, stream=None) -> int:
        """Bit mask of the ranks whose shard matched (on the root; 0 elsewhere)."""
        if isinstance(needles, np.ndarray):
            needles = np.ascontiguousarray(needles, dtype=np.uint8).reshape(-1, 32)
        m = C.c_uint32(0)
        _check(lib.yume_pdq_cuda_mailbox_scan(self._h, db._h, _ptr(needles), q, threshold, root, C.byref(m), _stream(stream)),
               "mailbox_scan")
        return int(m.value)

    def set_mode(self, mode: int):
        """MAILBOX_AUTO / MAILBOX_SEPARATE / MAILBOX_FUSED (yume_pdq_cuda.h)."""
        _check(lib.yume_pdq_cuda_mailbox_set_mode(self._h, mode), "mailbox_set_mode")

    def last_timing(self):
        """Host-side microsecond stamps of the last query on the root: after the post / wait / scan / publish + collect
        launches, and when the answer arrived."""
        t = (C.c_double * 5)()
        _check(lib.yume_pdq_cuda_mailbox_last_timing(self._h, t), "mailbox_last_timing")
        return list(t)

    def close(self):
        if self._h:
            lib.yume_pdq_cuda_mailbox_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DeviceBytes:
    """__cuda_array_interface__ over raw device memory (uint8)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": (int(nbytes),), "typestr": "|u1", "version": 2}


def device_view(ptr: int, nbytes: int, device: int | None = None):
    """A torch uint8 tensor over `nbytes` of device memory at `ptr` (no copy; the memory must outlive the tensor)."""
    import torch
    return torch.as_tensor(_DeviceBytes(ptr, nbytes), device=torch.device("cuda", torch.cuda.current_device() if device is None else device))


class PeerExchange:
    """This rank's end of the copy-engine needle all-gather (yume_pdq_cuda_exchange_*): `handle` goes to the peers,
    `connect` takes everyone's handles in rank order; `post(slice)` pushes this rank's slice to every peer, `wait()`
    makes a stream wait for the peers' slices and returns the device address of the packed batch of the other ranks'
    slices (rank order).  yume_pdq_b200.sharded.make_exchange wires it to torch.distributed."""

    def __init__(self, device: int, rank: int, world: int, slice_bytes: int):
        self._h = _vp()
        buf = (C.c_uint8 * IPC_HANDLE_BYTES)()
        _check(lib.yume_pdq_cuda_exchange_create(device, rank, world, slice_bytes, C.byref(self._h), buf), "exchange_create")
        self.handle = bytes(buf)
        self.rank, self.world, self.slice_bytes = rank, world, slice_bytes

    def connect(self, handles):
        blob = b"".join(handles)
        assert len(blob) == self.world * IPC_HANDLE_BYTES
        _check(lib.yume_pdq_cuda_exchange_connect(self._h, C.c_char_p(blob)), "exchange_connect")

    def post(self, slice_dev, stream=None, how: int = EXCHANGE_KERNEL, nbytes: int | None = None):
        """``nbytes``: a shorter slice this time (the same on every rank); the peers' buffers are packed at that stride."""
        _check(lib.yume_pdq_cuda_exchange_post_bytes(self._h, _ptr(slice_dev), self.slice_bytes if nbytes is None else nbytes,
                                                     how, _stream(stream)), "exchange_post")

    def wait(self, stream=None) -> int:
        p = _vp()
        _check(lib.yume_pdq_cuda_exchange_wait(self._h, _stream(stream), C.byref(p)), "exchange_wait")
        return int(p.value or 0)

    def failed(self) -> bool:
        f = C.c_int(0)
        _check(lib.yume_pdq_cuda_exchange_failed(self._h, C.byref(f)), "exchange_failed")
        return bool(f.value)

    def close(self):
        if self._h:
            lib.yume_pdq_cuda_exchange_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_frames(frames_dev, n: int, first_frame: int = 0, seed: int = 0x79756D652D706471, kind: int = 1,
                 device: int = 0, stream=None):
    _check(lib.yume_pdq_cuda_synth_frames(device, _ptr(frames_dev), n, first_frame, seed, kind,
                                          _stream(stream)), "synth_frames")


def measure_mxf4_peak(device: int = 0, iters: int = 20000):
    """(tera-ops/s, SM clocks per MMA) of back-to-back tcgen05.mma kind::mxf4 M128 N240 K64: the live roofline
    denominator of the tensor matcher."""
    t, c = C.c_double(0), C.c_double(0)
    _check(lib.yume_pdq_cuda_measure_mxf4_peak(device, iters, C.byref(t), C.byref(c)), "measure_mxf4_peak")
    return float(t.value), float(c.value)


def synth_pixels(frames_dev, n: int, pixel_format: int, first_frame: int = 0, seed: int = 0x79756D652D706471,
                 device: int = 0, stream=None):
    _check(lib.yume_pdq_cuda_synth_pixels(device, _ptr(frames_dev), pixel_format, n, first_frame, seed,
                                          _stream(stream)), "synth_pixels")


def synth_rows(rows_dev, n_rows: int, first_row: int = 0, seed: int = 0x79756D652D706471, device: int = 0,
               stream=None):
    _check(lib.yume_pdq_cuda_synth_rows(device, _ptr(rows_dev), n_rows, first_row, seed, _stream(stream)),
           "synth_rows")
