"""Pin the CPU oracle against every known answer the reference holds for the
hash path: README doctest vector, aaa-orig.jpg whitelist,
re-threshold identity, median semantics, dihedral identities + image-space
cross-check.  CPU only."""
import numpy as np
import pytest

from conftest import hamming
from oracle import pyoracle


def test_kat1_readme_doctest(oracle, kat):
    frame = pyoracle.readme_doctest_frame()
    q, thr, h, pdqf = oracle.hash_get_threshold(frame)
    exp = bytes.fromhex(kat["readme_doctest_hash"])
    assert hamming(h, exp) <= 2          # the reference's own assertion
    assert hamming(h, exp) == 0          # and what this restatement achieves
    table = np.array(kat["readme_doctest_pdqf_1dp"])
    assert np.abs(pdqf - table).max() <= 0.05 + 1e-4   # printed with 1 decimal
    assert q == 1.0


def test_kat2_aaa_orig(oracle, kat, aaa_luma):
    q, thr, h, pdqf = oracle.hash_get_threshold(aaa_luma)
    assert h.tobytes().hex() in kat["aaa_orig_expect"]
    assert h.tobytes().hex() == kat["readme_aaa"]["hash"]
    assert q == kat["readme_aaa"]["quality"] == 1.0
    # the README threshold (12.554) is the AVX2 bisection value, anywhere in
    # [sorted[127], sorted[128]); the scalar kernel returns sorted[127]
    s = np.sort(pdqf.ravel())
    assert thr == s[127]
    assert s[127] <= kat["readme_aaa"]["threshold_avx2"] < s[128]


def test_kat5_rethreshold_identity(oracle, aaa_luma):
    # lib.rs:659-687
    _, thr, h, pdqf = oracle.hash_get_threshold(aaa_luma)
    assert (oracle.threshold_2d_f32(pdqf, thr) == h).all()


def test_torben_is_lower_median(oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        p = rng.standard_normal((16, 16)).astype(np.float32)
        if rng.random() < 0.5:   # heavy ties
            p = np.round(p * 2) / 2
        assert oracle.torben_median(p) == np.sort(p.ravel())[127]


def test_bit_layout(oracle):
    p = np.zeros((16, 16), dtype=np.float32)
    p[0, 0] = 1.0      # flat index 0 -> byte 31, bit 0
    p[15, 15] = 1.0    # flat index 255 -> byte 0, bit 7
    p[1, 3] = 1.0      # flat index 19 -> byte 29, bit 3
    out, thr = oracle.quantize(p)
    assert thr == 0.0
    exp = np.zeros(32, dtype=np.uint8)
    exp[31] = 1
    exp[0] = 0x80
    exp[29] = 8
    assert (out == exp).all()


def test_gradient_quality(oracle):
    p = np.arange(256, dtype=np.float32).reshape(16, 16)
    g = oracle.sum_of_gradients(p)
    assert g == 15 * 16 * 16 + 16 * 15
    assert oracle.adjust_quality(90.0) == 0.5
    assert oracle.adjust_quality(1e9) == 1.0


def test_dihedral_identities(oracle, aaa_luma):
    _, thr, h, pdqf = oracle.hash_get_threshold(aaa_luma)
    ok, seen, final = oracle.visit_dihedrals(pdqf, thr)
    assert ok and len(seen) == 7
    assert [s[0] for s in seen] == oracle.dihedral_order[1:]
    C = lambda p: oracle.negate_alt_cols(p, False)
    R = lambda p: oracle.negate_alt_rows(p, True)
    P = pdqf
    states = [C(P), R(C(P)), R(P), P.T, C(P).T, C(R(P.T)), R(P).T]
    for (dih, t, q, hh, st), want in zip(seen, states):
        assert (st == want).all()
    # (C.P)^T = R.P^T and (R.P)^T = C.P^T
    assert (C(P).T == R(P.T)).all() and (R(P).T == C(P.T)).all()
    # final state documented in SURVEY 8(b): C.P^T
    assert (final == C(P.T)).all()
    # thresholds: 3 fresh medians then reuse
    thr_c, thr_od, thr_r = seen[0][1], seen[1][1], seen[2][1]
    assert [s[1] for s in seen[3:]] == [thr, thr_c, thr_od, thr_r]
    # early stop
    ok, seen2, _ = oracle.visit_dihedrals(pdqf, thr, stop_after=3)
    assert not ok and len(seen2) == 3


def test_dihedrals_vs_image_space(oracle, aaa_luma):
    """No reference KAT covers dihedrals; cross-check each frequency-domain hash
    against the hash of the actually transformed image (scripts/dihedrals.sh:11-18)."""
    hs, qs, ths, _ = oracle.hash_all_dihedrals(aaa_luma[None])
    img = aaa_luma
    flop = lambda a: a[:, ::-1]
    flip = lambda a: a[::-1, :]
    rot90 = lambda a: np.rot90(a, -1)      # clockwise, ImageMagick -rotate 90
    xforms = [img, flop(img), np.rot90(img, 2), flip(img), flop(rot90(img)),
              np.rot90(img, 1), rot90(flop(img)), rot90(img)]
    dists = []
    for i, x in enumerate(xforms):
        _, _, h, _ = oracle.hash_get_threshold(np.ascontiguousarray(x))
        dists.append(hamming(h, hs[0, i]))
    assert dists[0] == 0
    assert max(dists) <= 10, dists
    # the trait-default (odd-row) convention is NOT the contract: it complements ROTATED_180
    _, thr, _, pdqf = oracle.hash_get_threshold(img)
    bad = oracle.negate_alt_rows_trait_default(oracle.negate_alt_cols(pdqf, False), True)
    hb, _ = oracle.quantize(bad)
    assert hamming(hb, hs[0, 2]) > 200


def test_dihedral_constants(oracle):
    # lib.rs:85-134
    assert oracle.dihedral_order == [0x01000001, 0xFF000001, 0xFF0000FF, 0x010000FF,
                                     0x00FFFF00, 0x00010100, 0x00FF0100, 0x0001FF00]
