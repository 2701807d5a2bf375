"""The timed CPU baseline (AVX2 hasher / AVX-512 matcher restatement) agrees with
the scalar oracle within what the reference itself whitelists.  CPU only."""
import numpy as np
import pytest

from conftest import hamming
from helpers import match_tuples, plant
from oracle import pyoracle


@pytest.fixture(scope="module")
def cpu(oracle):
    return pyoracle.CpuBaseline()


def test_avx2_hasher_kats(cpu, oracle, kat, aaa_luma):
    if not cpu.has_avx2:
        pytest.skip("host without AVX2+FMA")
    frames = np.stack([aaa_luma, pyoracle.readme_doctest_frame()])
    h, q, thr, pdqf = cpu.hash_batch(frames, threads=2)
    assert h[0, 0].tobytes().hex() in kat["aaa_orig_expect"]
    assert hamming(h[1, 0], bytes.fromhex(kat["readme_doctest_hash"])) <= 2
    # README-reported AVX2 threshold for aaa-orig.jpg: 12.554 (README.md:164-168)
    assert abs(thr[0, 0] - kat["readme_aaa"]["threshold_avx2"]) < 5e-4
    ho, qo, tho, po = oracle.hash_all_dihedrals(frames)
    for f in range(2):
        for d in range(8):
            assert hamming(h[f, d], ho[f, d]) <= 2
    assert np.abs(pdqf - po).max() <= 1e-4 * np.abs(po).max()
    assert np.allclose(q, qo, rtol=1e-5)


def test_matcher_paths_identical_to_oracle(cpu, oracle):
    rng = np.random.default_rng(5)
    rows = oracle.synth_db(50_000)
    needles = rng.integers(0, 256, size=(40, 32), dtype=np.uint8)   # ragged: 32 + 8
    planted = plant(rows, needles, [0, 30, 31, 32], rng)
    want = match_tuples(oracle.matcher_find(rows, needles, 31))
    assert len(want) == 40 * 3
    for force_scalar in (True, False):
        for threads in (1, 3):
            got = match_tuples(cpu.match(rows, needles, 31, True, threads, force_scalar))
            assert got == want
            assert cpu.match(rows, needles, 31, False, threads, force_scalar) is True
    clean = oracle.synth_db(20_000, first_row=1 << 30)
    assert cpu.match(clean, needles[:8], 31, False, 2, False) is False
    assert cpu.match(clean, needles[:13], 31, False, 2, False) is False
