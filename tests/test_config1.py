"""BASELINE.json configs[0] / SURVEY 8(d) config 1, the reference's own CPU-runnable case: test-data/aaa-orig.jpg as
512 x 512 luma (PIL bilinear, as integration/hash.py:65-69; committed as tests/golden/aaa_orig_luma512.npz), its 8
dihedral hashes matched against a 1,000,000-row synthetic DB (seed "yume-pdq") at threshold 31, with a copy of each
dihedral hash planted at the exact distances {0, 15, 30, 31, 32, 33}.  The expected match set is exactly the plants
at <= 31.  CPU: the oracle alone (and the AVX-512 port); GPU: the product through the C ABI, both call shapes."""
import numpy as np
import pytest

from helpers import match_tuples, plant

SEED = 0x79756D652D706471
N_ROWS = 1_000_000


def _case(oracle, aaa_luma):
    hashes, quality, thr, pdqf = oracle.hash_all_dihedrals(aaa_luma[None])
    needles = np.ascontiguousarray(hashes[0])
    rows = oracle.synth_db(N_ROWS, 0, SEED)
    planted = plant(rows, needles, [0, 15, 30, 31, 32, 33], np.random.default_rng(2024))
    want = sorted(p for p in planted if p[2] <= 31)
    return needles, rows, want, (hashes, quality, thr, pdqf)


def test_config1_oracle(oracle, aaa_luma, kat):
    needles, rows, want, _ = _case(oracle, aaa_luma)
    assert needles[0].tobytes().hex() in kat["aaa_orig_expect"]
    assert len(want) == 8 * 4
    assert match_tuples(oracle.matcher_find(rows, needles, 31)) == want
    assert oracle.matcher_find_min(rows, needles, 31) == want[0][:2]


@pytest.mark.gpu
def test_config1_gpu(cuda, oracle, aaa_luma):
    needles, rows, want, (oh, oq, ot, op) = _case(oracle, aaa_luma)
    got = cuda.CudaKernel(0).hash_batch(aaa_luma[None], want_pdqf=True)
    assert (np.asarray(got["hashes"]) == oh).all() and np.array_equal(np.asarray(got["quality"]), oq)
    assert np.array_equal(np.asarray(got["thresholds"]), ot)
    assert np.array_equal(np.asarray(got["pdqf"]).view(np.uint32), op.view(np.uint32))
    db = cuda.CudaVectorDatabase(rows)
    m = cuda.CudaMatcher(db, 31)
    assert match_tuples(m.find_all(np.asarray(got["hashes"])[0])) == want
    assert m.scan(needles) and m.find_min(needles) == want[0][:2]
    # hash + dihedrals + match in one call, from the f32 frame and from the 8-bit frame the CLI would read
    s = m.screen(aaa_luma[None])
    assert match_tuples(s) == want
    s8 = m.screen(aaa_luma.astype(np.uint8)[None], pixel_format=cuda.PIXEL_LUMA8)
    assert match_tuples(s8) == want
    db.close()
