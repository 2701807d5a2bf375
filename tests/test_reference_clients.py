"""The reference's own users of its C ABI against this library.

* In a checkout that has the reference next to it (/root/reference, the build container), the UNMODIFIED
  `integration/hash.py` and `integration/dummy.c` are run / compiled against `libyume_pdq.so` (the alias of
  libyume_pdq_cuda.so), and the reference's `include/yume_pdq.h` is compiled in the same translation unit as ours:
  any declaration that differed would be a "conflicting types" error.  Without a GPU the two functions answer
  NaN / false (no CPU fallback); the clients still load, call and exit cleanly.
* On a GPU box the reference tree does not exist, so `ref_abi_client.py` / `integration/c/ref_abi_client.c` -- clients
  that make the same calls with the same argument types -- are checked against the reference's known answers.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(ROOT, "yume_pdq_b200")
ALIAS = os.path.join(LIB_DIR, "libyume_pdq.so")
REF = "/root/reference"
README_HASH = "58f8f0cee0f4a84f06370a32038f67f0b36e2ed596621e1d33e6b39c4e9c9b22"    # README.md:164-168

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "integration")), reason="no reference checkout")


def test_alias_library_exists(cuda):
    assert os.path.islink(ALIAS) and os.path.realpath(ALIAS) == os.path.realpath(cuda.LIB_PATH)


@needs_ref
def test_reference_header_and_ours_agree(tmp_path):
    src = tmp_path / "both.c"
    # the struct cannot be defined twice in one translation unit; every other declaration (13 extern objects, the
    # callback typedef, the two functions) is re-declared by our header and must be compatible with the reference's
    src.write_text('#include <stdint.h>\n#include "%s/include/yume_pdq.h"\n#define YUME_PDQ_DIHEDRALS_DEFINED\n'
                   '#include "yume_pdq.h"\n#include "yume_pdq_cuda.h"\n'
                   '_Static_assert(sizeof(YumePDQ_Dihedrals) == 4, "packed u32");\n'
                   'int main(void) { return (int)YUME_PDQ_DIHEDRAL_NORMAL.packed & 0; }\n' % REF)
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src), "-L" + LIB_DIR,
                    "-lyume_pdq", "-Wl,-rpath," + LIB_DIR, "-o", str(tmp_path / "both")], check=True, timeout=120)


@needs_ref
def test_reference_dummy_c_links_and_runs(tmp_path, cuda):
    exe = tmp_path / "dummy"
    # the command in integration/dummy.c:27, with -L pointing at this build
    subprocess.run(["gcc", "-o", str(exe), os.path.join(REF, "integration", "dummy.c"), "-L" + LIB_DIR, "-lyume_pdq",
                    "-Wl,-rpath," + LIB_DIR], check=True, timeout=120)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "First hash:" in r.stdout and "Second hash:" in r.stdout
    if cuda.device_count() == 0:
        assert "nan" in r.stdout.lower()          # no CPU fallback


@needs_ref
def test_reference_hash_py_runs_unmodified(cuda):
    pytest.importorskip("PIL")
    env = dict(os.environ, LIBYUME_PDQ_PATH=ALIAS)
    r = subprocess.run([sys.executable, os.path.join(REF, "integration", "hash.py"),
                        os.path.join(REF, "test-data", "aaa-orig.jpg")], capture_output=True, text=True, timeout=600,
                       env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Finished 1000 iterations" in r.stdout
    if cuda.device_count() > 0:
        assert "Hash: " + README_HASH in r.stdout and "Quality: 1.000" in r.stdout


@pytest.mark.gpu
def test_ref_abi_client_py_known_answer(cuda):
    """Same ctypes calls as the reference's integration/hash.py:32-92 (POINTER(c_float) argtypes, ctypes arrays),
    on the committed 512x512 luma of aaa-orig.jpg: README hash, quality 1.0, Torben threshold; with
    YUME_PDQ_CUDA_THRESHOLD=avx2 the README's 12.554."""
    client = os.path.join(ROOT, "integration", "python", "ref_abi_client.py")
    frame = os.path.join(ROOT, "tests", "golden", "aaa_orig_luma512.npz")
    for style, thr in ((None, "11.907"), ("avx2", "12.554")):
        env = dict(os.environ, LIBYUME_PDQ_PATH=ALIAS)
        env.pop("YUME_PDQ_CUDA_THRESHOLD", None)
        if style:
            env["YUME_PDQ_CUDA_THRESHOLD"] = style
        r = subprocess.run([sys.executable, client, frame, "200"], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "Hash: " + README_HASH in r.stdout and "Quality: 1.000" in r.stdout, r.stdout
        assert "Threshold: " + thr in r.stdout, r.stdout


@pytest.mark.gpu
def test_avx2_style_threshold_matches_port(cuda, oracle, monkeypatch):
    """YUME_PDQ_CUDA_THRESHOLD=avx2: the reported float equals the AVX2 kernel port's bisection threshold
    (x86.rs:440-621, oracle/cpu_baseline.cpp) bit for bit; hash, quality, PDQF are unchanged."""
    from oracle import pyoracle
    cpu = pyoracle.CpuBaseline()
    if not cpu.has_avx2:
        pytest.skip("host without AVX2")
    frames = np.concatenate([oracle.synth_frames(6, first_frame=3, kind=1), oracle.synth_frames(2, first_frame=9, kind=0)])
    for f in frames:
        monkeypatch.delenv("YUME_PDQ_CUDA_THRESHOLD", raising=False)
        q0, t0, h0, p0, b0 = cuda.hash_get_threshold(f)
        monkeypatch.setenv("YUME_PDQ_CUDA_THRESHOLD", "avx2")
        q1, t1, h1, p1, b1 = cuda.hash_get_threshold(f)
        assert q0 == q1 and (h0 == h1).all() and np.array_equal(p0, p1) and np.array_equal(b0, b1)
        # the port's AVX2 dct2d rounds differently from the scalar order, so feed it OUR pdqf
        want = cpu.quantize_avx2(p1)[1]
        assert np.float32(t1).view(np.uint32) == np.float32(want).view(np.uint32), (t1, want, t0)
