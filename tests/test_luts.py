"""The committed LUT header equals a fresh run of the 128-bit generator, and the
tables have the structure the reference's build.rs asserts (build.rs:649-652)."""
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _arr(text, name):
    m = re.search(name + r"\[[^\]]*\] = \{(.*?)\};", text, re.S)
    return [int(x.rstrip("u"), 16) if x.startswith("0x") else int(x)
            for x in re.findall(r"0x[0-9a-f]+u|\d+", m.group(1))]


def test_header_is_current():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "lut_gen.py"), "--check"])
    assert r.returncode == 0, "include/yume_pdq_luts.h is stale: run oracle/lut_gen.py"


def test_lut_structure():
    t = open(os.path.join(ROOT, "include", "yume_pdq_luts.h")).read()
    d = np.array(_arr(t, "YUME_PDQ_DCT_MATRIX_BITS"), dtype=np.uint32).view(np.float32).reshape(16, 127)
    i = np.arange(1, 17)[:, None]
    j = np.arange(127)[None, :]
    d64 = np.sqrt(2 / 127) * np.cos(np.pi / 254 * i * (2 * j + 1))
    assert np.abs(d - d64).max() < 1e-8
    # rows are orthonormal (DCT-II without the DC row)
    assert np.abs(d.astype(np.float64) @ d.astype(np.float64).T - np.eye(16)).max() < 1e-6
    w = np.array(_arr(t, "YUME_PDQ_TENT_WEIGHT_BITS"), dtype=np.uint32).view(np.float32).reshape(7, 7)
    tent = np.array([0, 1, 2, 3, 2, 1, 0], dtype=np.float64) / 9
    assert np.abs(w - np.outer(tent, tent)).max() < 1e-8
    inner = (np.outer(tent, tent)[1:6, 1:6]).astype(np.float32)
    assert (w[1:6, 1:6] == inner).all()
    # border: zero except the two ~1e-40 residues of the impulse-response procedure
    border = w.copy()
    border[1:6, 1:6] = 0
    assert border.max() < 2e-40 and (border > 0).sum() == 2
    off = _arr(t, "YUME_PDQ_CONV_OFFSET")
    assert off == [4 * (o + 1) for o in range(127)]
