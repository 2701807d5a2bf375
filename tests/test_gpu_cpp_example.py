"""The C++ mirror of the reference's Kernel / PDQMatcher surface (include/yume_pdq_cuda.hpp) compiled against the
C ABI and run on the GPU: integration/cpp/example.cpp hashes a frame, plants one of its dihedrals, finds it with
PDQMatcher::find semantics, and repeats the search through the luma8 + screen entry points."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_example(cuda, tmp_path):
    assert cuda.device_count() > 0
    exe = tmp_path / "example"
    lib_dir = os.path.join(ROOT, "yume_pdq_b200")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "integration", "cpp", "example.cpp"), "-L" + lib_dir, "-lyume_pdq_cuda",
                    "-Wl,-rpath," + lib_dir, "-o", str(exe)], check=True, timeout=300)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "found at row 123456, needle 5" in r.stdout and "screen: row 123456 needle 5 dist 4" in r.stdout


def test_c_example_of_the_reference_abi(cuda, oracle, tmp_path):
    """integration/c/example.c (plain C, the reference's own two functions): its 8 printed hashes and qualities equal
    the oracle's hash_get_threshold + visit_dihedrals of the same frame, in the reference's callback order."""
    import numpy as np
    exe = tmp_path / "c_example"
    lib_dir = os.path.join(ROOT, "yume_pdq_b200")
    subprocess.run(["gcc", "-std=c11", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "integration", "c", "example.c"),
                    "-L" + lib_dir, "-lyume_pdq_cuda", "-lm", "-Wl,-rpath," + lib_dir, "-o", str(exe)], check=True, timeout=300)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    y, x = np.mgrid[0:512, 0:512]
    frame = ((x * 7 + y * 3 + (x * y) // 64) % 256).astype(np.float32)
    hashes, quality, thr, pdqf = oracle.hash_all_dihedrals(frame[None])
    lines = r.stdout.strip().split("\n")
    assert len(lines) == 9 and lines[8].startswith("version ")
    for d in range(8):
        packed, q, h = lines[d].split()
        assert int(packed, 16) == cuda.DIHEDRAL_ORDER[d]
        assert h == hashes[0, d].tobytes().hex()
        assert abs(float(q) - float(quality[0, d])) < 1e-6
