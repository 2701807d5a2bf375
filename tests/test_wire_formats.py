"""Wire formats of `yume-pdq pipe` (src/main.rs:668-778, 1106-1133, 1268-1286) as produced by the
C++ front end (yume_pdq_b200/cli), checked on the CPU through its `format-records` hook against the independent
restatement in helpers.py and against the strings the reference documents."""
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import DIHEDRAL_ORDER, dihedral_pdqf, rust_fixed, wire_record, wire_stream

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "yume_pdq_b200", "yume-pdq-cuda")


@pytest.fixture(scope="module")
def cli():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ypdq_build", os.path.join(ROOT, "yume_pdq_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    assert os.path.exists(CLI)
    return CLI


def _raw_record(h, quality, threshold, star, dihedral, index, pdqf):
    return bytes(bytearray(h)) + struct.pack("<fffII", quality, threshold, star, dihedral, index) + \
        np.asarray(pdqf, dtype="<f4").tobytes()


def _run(cli, fmt, blob, sub="format-records", extra=()):
    r = subprocess.run([cli, sub, "-f", fmt, *extra], input=blob, capture_output=True, timeout=60)
    return r.returncode, r.stdout, r.stderr


BODIES = ["raw", "Raw", "RAW", "q+raw", "q+RAW", "hex", "HEX", "bin", "BIN", "q+hex", "q+HEX", "q+bin", "q+BIN",
          "diagnostic"]
SEPS = ["", "+lf", "+LF", "+crlf", "+cr", "+nul", "+NULL", "+nl"]


def test_every_format_and_separator(cli):
    rng = np.random.default_rng(3)
    n = 5
    hashes = rng.integers(0, 256, size=(n, 8, 32), dtype=np.uint8)
    quality = rng.random((n, 8)).astype(np.float32)
    quality[0, 0], quality[1, 0], quality[2, 0] = 1.0, 0.0, np.float32(0.0005)      # "100.000", "0.000", rounding
    thresholds = (rng.standard_normal((n, 8)) * 20).astype(np.float32)
    pdqf = (rng.standard_normal((n, 256)) * 50).astype(np.float32)
    pdqf[0, 0], pdqf[0, 1], pdqf[0, 2] = 0.0, -0.0, np.float32(1e-7)                # "-0.00000" like Rust
    for dihedrals in (False, True):
        blob = b""
        for k in range(n):
            for d in range(8 if dihedrals else 1):
                p = pdqf[k] if d == 0 else dihedral_pdqf(pdqf[k], d).reshape(-1)
                blob += _raw_record(hashes[k, d], quality[k, d], thresholds[k, d], thresholds[k, 0],
                                    0 if d == 0 else DIHEDRAL_ORDER[d], d, p)
        for body in BODIES:
            for sep in SEPS:
                rc, out, err = _run(cli, body + sep, blob)
                assert rc == 0, err
                assert out == wire_stream(body + sep, hashes, quality, thresholds, pdqf, dihedrals), (body, sep)


def test_documented_strings(cli, kat):
    """`pipe -f q+hex+lf` on aaa-orig prints "100.000:58f8f0ce...\\n" (main.rs:150-154, README.md:164-168)."""
    h = bytes.fromhex(kat["readme_aaa"]["hash"])
    blob = _raw_record(h, 1.0, 11.9068, 11.9068, 0, 0, np.zeros(256))
    assert _run(cli, "q+hex+lf", blob)[1] == b"100.000:" + kat["readme_aaa"]["hash"].encode() + b"\n"
    assert _run(cli, "HEX+crlf", blob)[1] == kat["readme_aaa"]["hash"].upper().encode() + b"\r\n"
    out = _run(cli, "bin", blob)[1]
    assert len(out) == 256 and out[:8] == format(h[0], "08b").encode()
    assert _run(cli, "q+raw", blob)[1] == struct.pack("<f", 1.0) + h
    diag = _run(cli, "diagnostic", blob)[1].decode().split("\n")
    assert diag[0] == '["original"]' and diag[1] == "Index: 0" and diag[2] == "Quality: 100.000"
    assert diag[3] == "Threshold: " + rust_fixed(np.float32(11.9068), 7) and diag[4] == "Hash:" and diag[6] == "PDQF:"
    assert diag[7] == "\t".join(["0.00000"] * 16)
    blob = _raw_record(h, 0.5, 1.0, 0.0, DIHEDRAL_ORDER[3], 3, np.ones(256))
    diag = _run(cli, "diagnostic", blob)[1].decode().split("\n")
    assert diag[0] == "[Dihedrals { packed: %d }]" % 0x010000FF and diag[1] == "Index: 3"
    assert diag[7] == "\t".join(["*1.00000"] * 16)          # star: above the ORIGINAL threshold (0.0 here)


@pytest.mark.parametrize("spec", ["", "lf", "hex+", "hex+tab", "q+", "Hex", "q+Hex", "Bin", "raw+hex", "DIAGNOSTIC", "bin+lf+lf+x"])
def test_invalid_formats_exit_255(cli, spec):
    for sub in ("format-records", "pipe"):           # parsed before any device is touched (main.rs:1261)
        rc, out, err = _run(cli, spec, b"", sub=sub)
        assert rc == 255 and out == b"" and b"invalid output format" in err


def test_format_is_parsed_like_the_reference(cli):
    h = bytes(range(32))
    blob = _raw_record(h, 0.25, 0, 0, 0, 0, np.zeros(256))
    # only the LAST component is tried as a separator (main.rs:1106-1133): "hex+lf+crlf" -> body "hex+lf": invalid
    assert _run(cli, "hex+lf+crlf", blob)[0] == 255
    assert _run(cli, "q+hex+nul", blob)[1] == b"25.000:" + h.hex().encode() + b"\0"
    # invalid input format is the other 255 (main.rs:1288-1295); unknown subcommand too (main.rs:1503)
    rc, _, err = _run(cli, "hex", b"", sub="pipe", extra=("--input-format", "yuv420"))
    assert rc == 255 and b"invalid input format" in err
    assert subprocess.run([cli, "frobnicate"], capture_output=True).returncode == 255


def test_rust_fixed_rounding():
    assert rust_fixed(np.float32(100.0), 3) == "100.000"
    assert rust_fixed(np.float32(0.0005) * np.float32(100), 3) == "0.050"
    assert rust_fixed(np.float32(2.5), 0) == "2" and rust_fixed(np.float32(3.5), 0) == "4"   # ties to even on exact ties
    assert rust_fixed(np.float32(-0.0), 5) == "-0.00000"
    assert wire_record("q+hex", bytes(32), np.float32(0.999994)) .startswith(b"99.999:")


def test_pipe_fails_loudly_without_a_gpu(cli):
    """No CPU fallback: with a valid format but no usable sm_100 device `pipe` exits 128 and says so."""
    import yume_pdq_b200 as yp
    if yp.device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([cli, "pipe", "-f", "hex+lf"], input=bytes(512 * 512), capture_output=True, timeout=60)
    assert r.returncode == 128 and r.stdout == b"" and b"no CPU fallback" in r.stderr
