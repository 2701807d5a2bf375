"""KAT-4: the reference's planted-positive matcher test (matching/mod.rs:460-538)
restated against the oracle, plus the inclusive-boundary plants.  CPU only."""
import numpy as np

from helpers import flip_bits_with_replacement, match_tuples, plant


def test_random_rows_do_not_match(oracle):
    rng = np.random.default_rng(0x79756D65)
    rows = rng.integers(0, 256, size=(2048, 32), dtype=np.uint8)
    for q in (8, 16, 24):
        needles = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
        assert not oracle.matcher_scan(rows, needles, 31)
        assert len(oracle.matcher_find(rows, needles, 31)) == 0
        assert oracle.matcher_find_min(rows, needles, 31) is None


def test_planted_positive_every_slot(oracle):
    rng = np.random.default_rng(7)
    rows = rng.integers(0, 256, size=(2048, 32), dtype=np.uint8)
    needles = rng.integers(0, 256, size=(8, 32), dtype=np.uint8)
    for i in range(0, 2048, 97):
        for j in range(8):
            saved = rows[i].copy()
            rows[i] = flip_bits_with_replacement(needles[j], 31, rng)
            assert oracle.matcher_scan(rows, needles, 31)
            m = oracle.matcher_find(rows, needles, 31)
            assert match_tuples(m)[0][:2] == (i, j)
            assert oracle.matcher_find_min(rows, needles, 31) == (i, j)
            rows[i] = saved


def test_inclusive_boundary_and_order(oracle):
    rng = np.random.default_rng(11)
    rows = oracle.synth_db(100_000)
    needles = rng.integers(0, 256, size=(8, 32), dtype=np.uint8)
    planted = plant(rows, needles, [0, 15, 30, 31, 32, 33], rng)
    m = match_tuples(oracle.matcher_find(rows, needles, 31))
    want = sorted((r, j, d) for r, j, d in planted if d <= 31)
    assert m == want                      # exactly the <=31 plants, (row, needle) order
    assert len(m) == 8 * 4
    m30 = match_tuples(oracle.matcher_find(rows, needles, 30))
    assert m30 == sorted((r, j, d) for r, j, d in planted if d <= 30)


def test_synth_db_is_counter_based(oracle):
    a = oracle.synth_db(1000, first_row=0)
    b = oracle.synth_db(500, first_row=500)
    assert (a[500:] == b).all()
    # balanced bits
    assert abs(np.unpackbits(a).mean() - 0.5) < 0.01
