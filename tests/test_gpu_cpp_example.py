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
