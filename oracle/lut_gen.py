#!/usr/bin/env python3
"""Regenerate the three numeric look-up tables of the yume-pdq hash path.

TEST/BUILD INFRASTRUCTURE.  The reference produces these tables at build time
with 128-bit MPFR (`rug` 1.27 -> gmp-mpfr-sys 1.6.5, neither vendored in
/root/reference).  This script restates that procedure with mpmath at
mp.prec = 128 and one round-to-nearest conversion to binary32, and writes
`include/yume_pdq_luts.h`, which both the CUDA kernels and the CPU oracle
include.  `tests/test_luts.py` re-runs this generator and checks the committed
header bit-for-bit.

Reference procedure restated here (file:line in /root/reference):
  * DCT matrix      build.rs:316-325 (`d_value`), emitted build.rs:604-615
                    (rows i = 1..=16, cols j = 0..127, n = 127)
  * tent weights    build.rs:334-381 (`tent_filter_weights`: impulse response
                    of two passes of the darwinium box filter on a 7x7 grid),
                    box filter build.rs:193-313 (`mean_box`)
  * sample offsets  build.rs:383-387 (`lerp`), build.rs:758-784
"""
from __future__ import annotations

import os
import struct
import sys

import mpmath as mp

mp.mp.prec = 128

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "include", "yume_pdq_luts.h")


def to_f32_bits(x) -> int:
    """Round an mpf to binary32 (RN-even, subnormals kept) and return its bits."""
    x = mp.mpf(x)
    if x == 0:
        return 0
    sign = 0
    if x < 0:
        sign = 0x80000000
        x = -x
    m, e = mp.frexp(x)  # x = m * 2**e, 0.5 <= m < 1
    # binary32: normal exponent range e-1 in [-126, 127]
    exp = e - 1
    if exp < -126:
        # subnormal: value = k * 2**-149
        k = mp.nint(mp.ldexp(x, 149))  # nint rounds half to even
        k = int(k)
        if k >= (1 << 23):  # rounded up into the normal range
            return sign | (1 << 23)
        return sign | k
    k = int(mp.nint(mp.ldexp(x, 23 - exp)))  # in [2**23, 2**24]
    if k == (1 << 24):
        k >>= 1
        exp += 1
    assert exp <= 127
    return sign | ((exp + 127) << 23) | (k - (1 << 23))


def f32_from_bits(b: int) -> float:
    return struct.unpack("<f", struct.pack("<I", b))[0]


# --------------------------------------------------------------------------
# DCT matrix, build.rs:316-325
def d_value(i: int, j: int, n: int):
    n1, i1, j1 = mp.mpf(n), mp.mpf(i), mp.mpf(j)
    two = mp.mpf(2)
    return mp.sqrt(two / n1) * mp.cos(mp.pi / (two * n1) * i1 * (two * j1 + 1))


def dct_matrix_bits():
    return [[to_f32_bits(d_value(i, j, 127)) for j in range(127)] for i in range(1, 17)]


# --------------------------------------------------------------------------
# darwinium box filter, build.rs:193-313
def box_one_d(invec, in_start, outvec, vector_length, stride, full_window):
    half_window = (full_window + 2) // 2
    phase_1_nreps = half_window - 1
    phase_2_nreps = full_window - half_window + 1
    oi_off = phase_1_nreps * stride
    li_off = phase_2_nreps * stride
    s = mp.mpf(0)
    cur = mp.mpf(0)
    phase_1_end = oi_off + in_start
    for ri in range(in_start, phase_1_end, stride):
        s += invec[ri]
        cur += 1
    phase_2_end = full_window * stride + in_start
    for ri in range(phase_1_end, phase_2_end, stride):
        oi = ri - oi_off
        s += invec[ri]
        cur += 1
        outvec[oi] = s / cur
    phase_3_end = vector_length * stride + in_start
    for ri in range(phase_2_end, phase_3_end, stride):
        oi = ri - oi_off
        li = oi - li_off
        s += invec[ri]
        s -= invec[li]
        outvec[oi] = s / cur
    phase_4_start = (vector_length - half_window + 1) * stride + in_start
    for oi in range(phase_4_start, phase_3_end, stride):
        li = oi - li_off
        s -= invec[li]
        cur -= 1
        outvec[oi] = s / cur


def jarosz_filter(buf, n_rows, n_cols, w_rows, w_cols, nreps):
    tmp = [mp.mpf(0)] * len(buf)
    for _ in range(nreps):
        for i in range(n_rows):
            box_one_d(buf, i * n_cols, tmp, n_cols, 1, w_rows)
        for j in range(n_cols):
            box_one_d(tmp, j, buf, n_rows, n_cols, w_cols)


def tent_filter_weights(old_dim=512, new_dim=127):
    """build.rs:334-381: returns (offset, (rows, cols), weights as mpf list)."""
    window = -(-old_dim // (2 * new_dim))  # div_ceil, build.rs:310-312
    rows = cols = window * 2 + 1
    out = [mp.mpf(0)] * (rows * cols)
    for di in range(rows):
        for dj in range(cols):
            tmp = [mp.mpf(0)] * (rows * cols)
            tmp[di * cols + dj] = mp.mpf(1)
            total = sum(tmp)
            tmp = [x / total for x in tmp]
            jarosz_filter(tmp, rows, cols, window, window, 2)
            middle = tmp[window * cols + window]
            # build.rs:369-371: kept only if it survives as a positive f32
            if f32_from_bits(to_f32_bits(middle)) > 0.0:
                out[di * cols + dj] = middle
    return (cols - 1) // 2, (rows, cols), out


# --------------------------------------------------------------------------
# sampling offsets, build.rs:383-387 and 771-773
def convolution_offsets():
    out = []
    for p in range(127):
        spaces = 127 + 1
        step = (512.0 - 0.0) / spaces
        v = 0.0 + step * (p + 1)
        out.append(int(round(v)))  # f64::round = half away from zero; values are exact ints here
    return out


def generate() -> str:
    dct = dct_matrix_bits()
    offset, (rows, cols), tent = tent_filter_weights()
    assert (offset, rows, cols) == (3, 7, 7), (offset, rows, cols)
    tent_bits = [to_f32_bits(x) for x in tent]
    offs = convolution_offsets()
    L = []
    L.append("/* GENERATED by oracle/lut_gen.py -- do not edit.")
    L.append(" *")
    L.append(" * Bit patterns (binary32) of the yume-pdq look-up tables, regenerated with")
    L.append(" * mpmath at 128-bit precision following /root/reference/build.rs:")
    L.append(" *   DCT matrix 16x127 (rows = frequencies 1..16)   build.rs:316-325, 604-615")
    L.append(" *   7x7 tent-filter impulse response               build.rs:334-381, 193-313")
    L.append(" *   127 sampling offsets (512 -> 127)              build.rs:383-387, 771-773")
    L.append(" */")
    L.append("#ifndef YUME_PDQ_LUTS_H")
    L.append("#define YUME_PDQ_LUTS_H")
    L.append("#include <stdint.h>")
    L.append("")
    L.append("#define YUME_PDQ_DCT_ROWS 16")
    L.append("#define YUME_PDQ_DCT_COLS 127")
    L.append("#define YUME_PDQ_TENT_OFFSET %d" % offset)
    L.append("#define YUME_PDQ_TENT_ROWS %d" % rows)
    L.append("#define YUME_PDQ_TENT_COLS %d" % cols)
    L.append("")
    L.append("/* D[i][j] = sqrt(2/127) * cos(pi/254 * (i+1) * (2j+1)), row-major, as u32 bit patterns */")
    L.append("static const uint32_t YUME_PDQ_DCT_MATRIX_BITS[16 * 127] = {")
    for i in range(16):
        row = dct[i]
        for c in range(0, 127, 8):
            L.append("    " + ", ".join("0x%08xu" % b for b in row[c:c + 8]) + ",")
    L.append("};")
    L.append("")
    L.append("/* W[di][dj]: weight of input pixel (4*o+1+di, 4*o'+1+dj) on output (o, o') */")
    L.append("static const uint32_t YUME_PDQ_TENT_WEIGHT_BITS[7 * 7] = {")
    for r in range(7):
        L.append("    " + ", ".join("0x%08xu" % b for b in tent_bits[r * 7:(r + 1) * 7]) + ",")
    L.append("};")
    L.append("")
    L.append("/* CONVOLUTION_OFFSET_512_TO_127[o] = round(512/128 * (o+1)) = 4*(o+1) */")
    L.append("static const uint16_t YUME_PDQ_CONV_OFFSET[127] = {")
    for c in range(0, 127, 16):
        L.append("    " + ", ".join("%d" % v for v in offs[c:c + 16]) + ",")
    L.append("};")
    L.append("")
    L.append("#endif /* YUME_PDQ_LUTS_H */")
    return "\n".join(L) + "\n"


if __name__ == "__main__":
    text = generate()
    if len(sys.argv) > 1 and sys.argv[1] == "--check":
        with open(OUT) as f:
            sys.exit(0 if f.read() == text else 1)
    with open(OUT, "w") as f:
        f.write(text)
    print("wrote", os.path.normpath(OUT))
