"""GPU parity of the matcher against the CPU oracle (bit-exact match sets), through the
C ABI: KAT-4 planted positives, inclusive boundary, every kernel variant, ragged sizes."""
import numpy as np
import pytest

from helpers import flip_bits, flip_bits_with_replacement, match_tuples, plant

pytestmark = pytest.mark.gpu


def _find_device(cuda, matcher, needles, algo, cap=1 << 16):
    import torch
    q = needles.shape[0]
    nd = torch.from_numpy(needles).cuda()
    out = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    matcher.find_device(nd, q, out, cap, cnt, algo=algo)
    torch.cuda.synchronize()
    n = int(cnt.item())
    assert n <= cap
    m = out[:n].cpu().numpy().view(cuda.MATCH_DTYPE).reshape(-1)
    return sorted(match_tuples(m))


@pytest.mark.parametrize("n_rows", [1, 511, 4096, 100_003])
@pytest.mark.parametrize("q", [1, 8, 13, 32, 40])
def test_planted_vs_oracle_all_algos(cuda, oracle, n_rows, q):
    rng = np.random.default_rng(n_rows * 131 + q)
    rows = oracle.synth_db(n_rows, first_row=77)
    needles = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
    if n_rows >= 6 * q:
        plant(rows, needles, [0, 15, 30, 31, 32, 33], rng)
    else:
        rows[0] = flip_bits(needles[q - 1], 31, rng)
    want = match_tuples(oracle.matcher_find(rows, needles, 31))
    db = cuda.CudaVectorDatabase(rows)
    m = cuda.CudaMatcher(db, 31)
    assert match_tuples(m.find_all(needles)) == want          # sorted (row, needle) like CpuMatcher::find
    assert m.scan(needles) == (len(want) > 0)
    assert m.find_min(needles) == oracle.matcher_find_min(rows, needles, 31)
    for algo in (cuda.ALGO_EXACT, cuda.ALGO_TILED, cuda.ALGO_TENSOR, cuda.ALGO_TENSOR_I8,
                 cuda.ALGO_TENSOR_FULL, cuda.ALGO_TENSOR_PREFIX, cuda.ALGO_TENSOR_PAIRED) + (
                     (cuda.ALGO_TENSOR_ROWS,) if q <= 64 else ()):
        assert _find_device(cuda, m, needles, algo) == sorted(want), algo
    db.close()


@pytest.mark.parametrize("threshold", [0, 1, 30, 31, 32, 40, 41, 64, 128, 256])
def test_thresholds(cuda, oracle, threshold):
    rng = np.random.default_rng(threshold)
    rows = oracle.synth_db(20_000)
    needles = rng.integers(0, 256, size=(8, 32), dtype=np.uint8)
    dists = sorted({0, max(threshold - 1, 0), threshold, min(threshold + 1, 256)} & set(range(257)))
    plant(rows, needles, dists, rng)
    want = match_tuples(oracle.matcher_find(rows, needles, threshold))
    db = cuda.CudaVectorDatabase(rows)
    m = cuda.CudaMatcher(db, threshold)
    if len(want) < 200_000:
        assert match_tuples(m.find_all(needles, cap=1 << 18)) == want
    else:   # threshold 128/256 match (nearly) everything: compare counts and a prefix
        got = m.find_all(needles, cap=1 << 18)
        assert len(got) == len(want) and match_tuples(got[:1000]) == want[:1000]
    if threshold <= 40:
        for algo in (cuda.ALGO_TILED,):
            assert _find_device(cuda, m, needles, algo, cap=1 << 18) == sorted(want)
    if len(want) < 200_000:     # the tensor-core paths are exact for every threshold they accept
        assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR, cap=1 << 18) == sorted(want)
        assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR_I8, cap=1 << 18) == sorted(want)
        if threshold <= 127:
            assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR_FULL, cap=1 << 18) == sorted(want)
        if threshold <= 63:    # beyond 36 nearly every pair is a candidate: slow, still exact
            assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR_PREFIX, cap=1 << 18) == sorted(want)
        if threshold <= 60:
            assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR_PAIRED, cap=1 << 18) == sorted(want)
        if threshold <= 128:
            assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR_ROWS, cap=1 << 18) == sorted(want)
    db.close()
    # the forced variants reject thresholds they cannot represent instead of answering wrongly
    if threshold > 127:
        with pytest.raises(cuda.YumePdqCudaError):
            _find_device(cuda, cuda.CudaMatcher(cuda.CudaVectorDatabase(rows[:64]), threshold), needles,
                         cuda.ALGO_TENSOR_FULL)
    if threshold > 63:
        with pytest.raises(cuda.YumePdqCudaError):
            _find_device(cuda, cuda.CudaMatcher(cuda.CudaVectorDatabase(rows[:64]), threshold), needles,
                         cuda.ALGO_TENSOR_PREFIX)
        with pytest.raises(cuda.YumePdqCudaError):
            _find_device(cuda, cuda.CudaMatcher(cuda.CudaVectorDatabase(rows[:64]), threshold), needles,
                         cuda.ALGO_TENSOR_PAIRED)


def test_reference_planted_positive_procedure(cuda, oracle):
    """matching/mod.rs:460-538: 2048 random rows, 8/16/24 needles, no match; then plant a copy with
    31 random flips at every slot (subsampled) and require scan + find -> (row, needle)."""
    rng = np.random.default_rng(0x79756D65)
    rows = rng.integers(0, 256, size=(2048, 32), dtype=np.uint8)
    for q in (8, 16, 24):
        needles = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
        db = cuda.CudaVectorDatabase(rows)
        m = cuda.CudaMatcher(db, 31)
        assert not m.scan(needles) and len(m.find_all(needles)) == 0 and m.find_min(needles) is None
        for i in range(0, 2048, 331):
            for j in range(0, q, 3):
                db.write_rows(i, flip_bits_with_replacement(needles[j], 31, rng)[None])
                assert m.scan(needles)
                assert m.find(needles, lambda r, n: (r, n)) == (i, j)
                assert m.find_min(needles) == (i, j)
                db.write_rows(i, rows[i][None])
        db.close()


def test_large_needle_batch_multi_chunk(cuda, oracle):
    """q > 2048 spans several needle chunks of the tiled kernel (the INT-pipe regime)."""
    rng = np.random.default_rng(21)
    q = 5000
    rows = oracle.synth_db(30_011, first_row=10**9)
    needles = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
    sel = rng.choice(q, size=300, replace=False)
    pos = rng.choice(rows.shape[0], size=300, replace=False)
    for j, r in zip(sel, pos):
        rows[r] = flip_bits(needles[j], int(rng.integers(0, 34)), rng)
    want = match_tuples(oracle.matcher_find(rows, needles, 31))
    assert len(want) > 200
    db = cuda.CudaVectorDatabase(rows)
    m = cuda.CudaMatcher(db, 31)
    assert match_tuples(m.find_all(needles)) == want
    assert _find_device(cuda, m, needles, cuda.ALGO_EXACT) == sorted(want)
    assert _find_device(cuda, m, needles, cuda.ALGO_TILED) == sorted(want)
    assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR) == sorted(want)
    assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR_I8) == sorted(want)
    assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR_FULL) == sorted(want)
    assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR_PREFIX) == sorted(want)
    assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR_PAIRED) == sorted(want)
    db.close()


def _flip_in(h, k_lo, k_hi, rng):
    """Copy of hash h with exactly k_lo distinct bits flipped inside bytes 0..15 (the 128 bits the tensor prefix
    filter looks at) and k_hi inside bytes 16..31."""
    bits = np.unpackbits(h)
    pos = np.concatenate([rng.choice(128, size=k_lo, replace=False), 128 + rng.choice(128, size=k_hi, replace=False)])
    bits[pos.astype(np.int64)] ^= 1
    return np.packbits(bits)


@pytest.mark.parametrize("q", [8, 700])
def test_prefix_filter_boundaries(cuda, oracle, q):
    """The 128-bit prefix filter of the tensor matcher: a pair is a candidate iff its distance over bytes 0..15 is
    <= threshold, and a candidate is a match iff its full distance is.  Plants pin both decisions on both sides."""
    rng = np.random.default_rng(1000 + q)
    rows = oracle.synth_db(50_021, first_row=5)
    needles = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
    cases = [(31, 0), (0, 31), (16, 15), (32, 0), (0, 32), (31, 1), (30, 1), (20, 20), (1, 31), (0, 0), (31, 97),
             (64, 0), (0, 128)]
    pos = rng.choice(rows.shape[0], size=len(cases) * 8, replace=False)
    k = 0
    for rep in range(8):
        for lo, hi in cases:
            rows[pos[k]] = _flip_in(needles[(rep * 89 + 7 * k) % q], lo, hi, rng)
            k += 1
    # tile edges: last row of a 240-row tile / 480-row super-tile, first of the next, last row of the DB
    for r in (127, 128, 239, 240, 255, 256, 367, 368, 479, 480, 735, 736, 847, 848, 959, 960, 1199, 1200, rows.shape[0] - 1):
        rows[r] = _flip_in(needles[r % q], 15, 16, rng)
    want = sorted(match_tuples(oracle.matcher_find(rows, needles, 31)))
    assert len(want) >= 8 * 5 + 5
    db = cuda.CudaVectorDatabase(rows)
    m = cuda.CudaMatcher(db, 31)
    for algo in (cuda.ALGO_TENSOR_PREFIX, cuda.ALGO_TENSOR_PAIRED, cuda.ALGO_TENSOR_FULL, cuda.ALGO_TENSOR, cuda.ALGO_AUTO):
        assert _find_device(cuda, m, needles, algo) == want, algo
    if q <= 32:
        assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR_ROWS) == want
        for thr in (0, 15, 36, 37, 50, 63):
            w = sorted(match_tuples(oracle.matcher_find(rows, needles, thr)))
            assert _find_device(cuda, cuda.CudaMatcher(db, thr), needles, cuda.ALGO_TENSOR_ROWS, cap=1 << 18) == w, thr
    for thr in (0, 15, 36, 37, 50, 63):
        mt = cuda.CudaMatcher(db, thr)
        w = sorted(match_tuples(oracle.matcher_find(rows, needles, thr)))
        if q <= 8 or thr <= 37:        # 700 needles at thr 50+: every pair is a candidate, slow but covered at q = 8
            assert _find_device(cuda, mt, needles, cuda.ALGO_TENSOR_PREFIX, cap=1 << 18) == w, thr
            if thr <= 60:
                assert _find_device(cuda, mt, needles, cuda.ALGO_TENSOR_PAIRED, cap=1 << 18) == w, thr
        assert _find_device(cuda, mt, needles, cuda.ALGO_TENSOR, cap=1 << 18) == w, thr
    db.close()


@pytest.mark.parametrize("q", [5, 300])
def test_paired_accumulator_fields(cuda, oracle, q):
    """ALGO_TENSOR_PAIRED packs the prefix dot products of two rows of a 960-row super-tile into one f32 accumulator
    as two s16 fields: rows c and c + 240 of a group of 480 rows (groups start at 0 and 480; the older 128 / 112
    geometry is still planted, it only spreads the columns); the high field does not see hash bits 0, 4, 8, 12, 16,
    20 of word 0 (bits 0 and 4 of bytes 0..2), whose operand positions carry the bias constants.  Plant every
    combination of prefix distances in the two rows of one column -- including the ones that push the packed value
    out of its exponent window, a negative low field that borrows from the high one, and differences confined to
    the six bias positions -- and require the exact match set."""
    rng = np.random.default_rng(77 + q)
    n_rows = 960 * 40 + 123
    rows = oracle.synth_db(n_rows, first_row=11)
    needles = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
    groups = [(0, 128), (256, 112), (480, 128), (736, 112), (0, 240), (480, 240)]
    bias_pos = [(0, 0), (0, 4), (1, 0), (1, 4), (2, 0), (2, 4)]      # (byte, bit) of the six bias positions

    def flip_bias_bits(h, k):        # flip k of the six hash bits the high field does not see
        out = h.copy()
        for b in rng.choice(6, size=k, replace=False):
            byte, bit = bias_pos[int(b)]
            out[byte] ^= np.uint8(1 << bit)
        return out

    lo_cases = [(0, 0), (31, 0), (32, 0), (0, 31), (16, 15), (128, 0), (100, 3), (64, 0), (33, 0), (127, 0)]
    hi_cases = [(0, 0), (31, 0), (32, 0), (0, 31), (16, 15), (128, 0), (97, 0), (96, 0), (30, 1), (1, 0), (127, 128),
                "bias6", "bias3+29", "bias6+26"]
    col = 0
    for a in lo_cases:
        for b in hi_cases:
            st = (col * 7) % (n_rows // 960 - 1)
            g0, w = groups[col % len(groups)]
            c = (col * 53) % w
            needle = needles[(col * 3) % q]
            rows[st * 960 + g0 + c] = _flip_in(needle, a[0], a[1], rng)                 # low field of the column
            if b == "bias6":            # distance 6, all of it invisible to the high field
                hi = flip_bias_bits(needle, 6)
            elif b == "bias3+29":       # distance 32: 3 invisible + 29 visible -> candidate, not a match
                hi = flip_bias_bits(_flip_in(needle, 0, 29, rng), 3)
            elif b == "bias6+26":       # distance 32 again, 26 visible
                hi = flip_bias_bits(_flip_in(needle, 0, 26, rng), 6)
            else:
                hi = _flip_in(needle, b[0], b[1], rng)
            rows[st * 960 + g0 + w + c] = hi                                            # high field of the same column
            col += 1
    for thr in (31, 0, 36, 60):
        want = sorted(match_tuples(oracle.matcher_find(rows, needles, thr)))
        assert len(want) >= 4
        m = cuda.CudaMatcher(cuda.CudaVectorDatabase(rows), thr)
        if q <= 8 or thr <= 36:
            assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR_PAIRED, cap=1 << 18) == want, thr
        assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR_FULL, cap=1 << 18) == want, thr


def _with_prefix_popcount(rng, k, suffix=None):
    """32-byte hash whose first 128 bits hold exactly k ones (random positions)."""
    bits = np.zeros(256, dtype=np.uint8)
    bits[rng.choice(128, size=k, replace=False)] = 1
    bits[128:] = np.unpackbits(suffix)[128:] if suffix is not None else rng.integers(0, 2, size=128, dtype=np.uint8)
    return np.packbits(bits)


@pytest.mark.parametrize("thr", [31, 0, 12, 60])
def test_paired_sparse_and_dense_operands(cuda, oracle, thr):
    """The tensor matcher's rows are 0/1-valued operands (hash bits masked into place), so a prefix dot product is
    popc(needle) - distance and the candidate minimum is per needle: it is NEGATIVE for needles with fewer set prefix
    bits than the threshold (the all-zero hash of a flat frame is the realistic case), and a dense row in the high
    field of the same accumulator column then moves the packed value below its exponent window, where the low field
    reads doubled.  Plant sparse / dense / all-zero / all-one needles against sparse / dense / all-zero / all-one rows
    in both fields of shared columns and require the exact match set from every tensor variant."""
    rng = np.random.default_rng(4242 + thr)
    n_rows = 960 * 56 + 77
    rows = oracle.synth_db(n_rows, first_row=99)
    pops = [0, 1, 5, 10, 20, 30, 31, 32, 33, 64, 96, 100, 127, 128]
    needles = np.stack([_with_prefix_popcount(rng, k) for k in pops] +
                       [np.zeros(32, np.uint8), np.full(32, 255, np.uint8)] +
                       [rng.integers(0, 256, size=32, dtype=np.uint8) for _ in range(6)])
    q = needles.shape[0]
    zeros, ones = np.zeros(32, np.uint8), np.full(32, 255, np.uint8)
    col = 0
    for j in range(q):
        for d in (0, max(thr - 10, 0), max(thr - 1, 0), thr, thr + 1, min(thr + 40, 128)):
            near = _flip_in(needles[j], min(d, 128), 0, rng)            # prefix distance d, suffix equal
            for other in (ones, zeros, _with_prefix_popcount(rng, 120), _with_prefix_popcount(rng, 3), None):
                for near_is_low in (True, False):
                    st, stage, c = col % 56, (col // 56) % 2, (col * 37) % 240     # distinct columns
                    r_lo = st * 960 + stage * 480 + c
                    a, b = (near, other) if near_is_low else (other, near)
                    if a is not None:
                        rows[r_lo] = a
                    if b is not None:
                        rows[r_lo + 240] = b
                    col += 1
    want = sorted(match_tuples(oracle.matcher_find(rows, needles, thr)))
    assert len(want) > 400
    m = cuda.CudaMatcher(cuda.CudaVectorDatabase(rows), thr)
    for algo in (cuda.ALGO_TENSOR_PAIRED, cuda.ALGO_TENSOR_PREFIX, cuda.ALGO_TENSOR_FULL, cuda.ALGO_TENSOR, cuda.ALGO_AUTO,
                 cuda.ALGO_TILED):
        assert _find_device(cuda, m, needles, algo, cap=1 << 18) == want, algo
    # the same needles padded to several tiles (the per-needle minima are indexed by tile)
    many = np.concatenate([rng.integers(0, 256, size=(300, 32), dtype=np.uint8), needles])
    want2 = sorted(match_tuples(oracle.matcher_find(rows, many, thr)))
    assert _find_device(cuda, m, many, cuda.ALGO_TENSOR_PAIRED, cap=1 << 18) == want2


def test_paired_needle_chunks(cuda, oracle):
    """A small shard is cut along the needle axis too (work unit = super-tile x chunk of >= 32 needle tiles), which
    is how an 8-GPU split of a 10M-row DB keeps 148 SMs busy: 12,365 needles (97 tiles -> 3 chunks) against 3,000 and
    961 rows, matches planted in every chunk and at the chunk edges."""
    rng = np.random.default_rng(99)
    q = 3 * 32 * 128 + 77
    needles = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
    for n_rows in (3000, 961):
        rows = oracle.synth_db(n_rows, first_row=3)
        sel = np.concatenate([rng.choice(q, size=150, replace=False),
                              [0, 33 * 128 - 1, 33 * 128, 66 * 128 - 1, 66 * 128, q - 1]])
        pos = rng.choice(n_rows, size=len(sel), replace=False)
        for j, r in zip(sel, pos):
            rows[r] = flip_bits(needles[j], int(rng.integers(0, 34)), rng)
        want = sorted(match_tuples(oracle.matcher_find(rows, needles, 31)))
        assert len(want) > 100
        m = cuda.CudaMatcher(cuda.CudaVectorDatabase(rows), 31)
        assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR_PAIRED) == want
        assert _find_device(cuda, m, needles, cuda.ALGO_AUTO) == want


def test_adversarial_dense_matches(cuda, oracle):
    """Near-duplicate database: every prefix filter survives, the candidate queue of the tiled kernel
    overflows, the output list overflows the first capacity guess.  Correctness must not depend on
    the data."""
    rng = np.random.default_rng(4)
    base = rng.integers(0, 256, size=32, dtype=np.uint8)
    rows = np.stack([flip_bits(base, int(rng.integers(0, 40)), rng) for _ in range(6000)])
    needles = np.stack([flip_bits(base, int(rng.integers(0, 6)), rng) for _ in range(64)])
    want = match_tuples(oracle.matcher_find(rows, needles, 31))
    assert len(want) > 100_000
    db = cuda.CudaVectorDatabase(rows)
    m = cuda.CudaMatcher(db, 31)
    got = m.find_all(needles, cap=1000)        # forces the E_CAPACITY retry path
    assert match_tuples(got) == want
    for algo in (cuda.ALGO_EXACT, cuda.ALGO_TILED, cuda.ALGO_TENSOR, cuda.ALGO_TENSOR_I8, cuda.ALGO_TENSOR_FULL,
                 cuda.ALGO_TENSOR_PREFIX, cuda.ALGO_TENSOR_PAIRED):
        assert _find_device(cuda, m, needles, algo, cap=1 << 19) == sorted(want)
    # capacity error contract of the C ABI: first `cap` entries in sorted order + total
    import ctypes as C
    out = np.zeros(10, dtype=cuda.MATCH_DTYPE)
    n = C.c_size_t(0)
    rc = cuda.lib.yume_pdq_cuda_matcher_find(db._h, cuda._ptr(needles), 64, 31, cuda._ptr(out), 10,
                                             C.byref(n), None)
    assert rc == cuda.E_CAPACITY and n.value == len(want) and match_tuples(out) == want[:10]
    db.close()


def test_global_row_offset_and_empty(cuda, oracle):
    rng = np.random.default_rng(8)
    rows = oracle.synth_db(5000)
    needles = rng.integers(0, 256, size=(8, 32), dtype=np.uint8)
    plant(rows, needles, [3, 31], rng)
    want = match_tuples(oracle.matcher_find(rows, needles, 31))
    off = 3_000_000_000
    db = cuda.CudaVectorDatabase(rows, global_row_offset=off)
    m = cuda.CudaMatcher(db, 31)
    assert match_tuples(m.find_all(needles)) == [(r + off, j, d) for r, j, d in want]
    assert m.find_min(needles) == oracle.matcher_find_min(rows, needles, 31)   # local row, like Vulkan
    db.close()
    empty = cuda.CudaVectorDatabase(np.zeros((0, 32), dtype=np.uint8))
    assert len(empty) == 0 and not cuda.CudaMatcher(empty, 31).scan(needles)
    assert len(cuda.CudaMatcher(empty, 31).find_all(needles)) == 0
    with pytest.raises(cuda.YumePdqCudaError):
        cuda.CudaMatcher(db if False else empty, 31).find_all(needles[:0])    # q == 0 is invalid


def test_headline_config_10m(cuda, oracle):
    """BASELINE config 3: 10,000,384-row synthetic DB (benches/matcher.rs:18-24), 8 needles,
    threshold 31: random needles do not match (benches/matcher.rs:79 asserts !matched); ~100
    planted rows are found exactly -- by every kernel variant."""
    n = (10_000_000 // 2048 + 1) * 2048
    rng = np.random.default_rng(31)
    db = cuda.CudaVectorDatabase(synthetic_rows=n, seed=0x79756D652D706471)
    assert len(db) == n
    # the device generator equals the oracle generator (slice check)
    import torch
    base = torch.empty(0)
    sl = oracle.synth_db(4096, first_row=n - 4096)
    import ctypes as C
    buf = np.zeros((4096, 32), dtype=np.uint8)
    torch.cuda.synchronize()
    t = torch.zeros((4096, 32), dtype=torch.uint8, device="cuda")
    cuda.synth_rows(t, 4096, first_row=n - 4096)
    torch.cuda.synchronize()
    assert np.array_equal(t.cpu().numpy(), sl)
    needles = rng.integers(0, 256, size=(8, 32), dtype=np.uint8)
    m = cuda.CudaMatcher(db, 31)
    assert not m.scan(needles)
    planted = []
    for r in rng.choice(n, size=104, replace=False):
        j = int(rng.integers(0, 8))
        d = int(rng.choice([0, 7, 30, 31, 32, 33, 40]))
        db.write_rows(int(r), flip_bits(needles[j], d, rng)[None])
        planted.append((int(r), j, d))
    want = sorted(p for p in planted if p[2] <= 31)
    assert match_tuples(m.find_all(needles)) == want
    assert m.find_min(needles) == (want[0][0], want[0][1])
    for algo in (cuda.ALGO_EXACT, cuda.ALGO_TILED, cuda.ALGO_TENSOR, cuda.ALGO_TENSOR_I8,
                 cuda.ALGO_TENSOR_FULL, cuda.ALGO_TENSOR_PREFIX, cuda.ALGO_TENSOR_PAIRED, cuda.ALGO_TENSOR_ROWS):
        assert _find_device(cuda, m, needles, algo) == want
    # last rows of the last chunk are planted too (tail handling)
    db.write_rows(n - 1, needles[5][None])
    assert match_tuples(m.find_all(needles))[-1] == (n - 1, 5, 0)
    db.close()


def _prefix_twins(rng, n_rows, q, n_twins, n_matches, oracle):
    """A haystack the prefix filter is bad at: n_twins rows share the first 128 bits (within 0..31 bits) with one of q
    random needles but not the rest -- candidates that are not matches -- plus n_matches true near-duplicates."""
    needles = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
    rows = oracle.synth_db(n_rows, first_row=5)
    where = rng.choice(n_rows, size=n_twins + n_matches, replace=False)
    for k, r in enumerate(where):
        n = needles[int(rng.integers(0, q))]
        if k < n_twins:
            rows[r, :16] = _flip_in(n, int(rng.integers(0, 32)), 0, rng)[:16]        # the suffix stays random
        else:
            rows[r] = flip_bits(n, int(rng.integers(0, 34)), rng)
    return rows, needles


def _clustered(rng, n_rows, q, n_near, oracle):
    """A near-duplicate family: q needles within 6 bits of one base hash (consecutive frames of a video) and n_near
    rows within 0..40 bits of it, scattered through an otherwise random haystack."""
    base = rng.integers(0, 256, size=32, dtype=np.uint8)
    needles = np.stack([flip_bits(base, int(rng.integers(0, 7)), rng) for _ in range(q)])
    rows = oracle.synth_db(n_rows, first_row=5)
    for r in rng.choice(n_rows, size=n_near, replace=False):
        rows[r] = flip_bits(base, int(rng.integers(0, 41)), rng)
    return rows, needles


def _find_device_np(cuda, matcher, needles_dev, q, algo, cap):
    """sorted (row, needle, dist) as one structured array -- for match sets too large for Python tuples"""
    import torch
    out = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    matcher.find_device(needles_dev, q, out, cap, cnt, algo=algo)
    torch.cuda.synchronize()
    n = int(cnt.item())
    assert n <= cap
    return np.sort(out[:n].cpu().numpy().view(cuda.MATCH_DTYPE).reshape(-1), order=["row", "needle"])


def test_auto_falls_back_on_dense_haystack(cuda, oracle):
    """ALGO_AUTO on a haystack full of prefix twins: the first large-batch launch runs the prefix
    filter and reports how many accumulator batches flagged and how many matches they held; from the next launch on the
    unfiltered tensor kernel runs (fallback), probes of the filter on a slice keep it there, and a batch of unrelated
    needles brings the filter back.  Every launch returns the oracle's match set."""
    import torch
    rng = np.random.default_rng(77)
    n_rows, q = 400_000, 1024
    rows, needles = _prefix_twins(rng, n_rows, q, 40_000, 300, oracle)
    want = np.sort(oracle.matcher_find(rows, needles, 31), order=["row", "needle"])
    assert 200 < len(want) < 400
    db = cuda.CudaVectorDatabase(rows)
    m = cuda.CudaMatcher(db, 31)
    cap = 1 << 16
    nd = torch.from_numpy(needles).cuda()
    find = lambda dev, algo=cuda.ALGO_AUTO: _find_device_np(cuda, m, dev, q, algo, cap)
    assert db.auto_state == (False, 0)
    assert np.array_equal(find(nd), want)                   # prefix filter; leaves its statistics
    assert db.auto_state[0] is False
    assert np.array_equal(find(nd), want)                   # sees them: unfiltered from here on
    assert db.auto_state == (True, 8)
    for _ in range(18):                                     # the probes at launches 8 and 16 find dense slices
        assert np.array_equal(find(nd), want)
    fb, interval = db.auto_state
    assert fb and interval >= 16
    sparse = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
    want_sparse = np.sort(oracle.matcher_find(rows, sparse, 31), order=["row", "needle"])
    sd = torch.from_numpy(sparse).cuda()
    for _ in range(2 * interval + 2):
        assert np.array_equal(find(sd), want_sparse)
    assert db.auto_state[0] is False                        # a sparse probe ended the fallback
    for algo in (cuda.ALGO_TENSOR_PAIRED, cuda.ALGO_TENSOR_FULL, cuda.ALGO_TILED):
        assert np.array_equal(find(nd, algo), want)
    db.close()


def test_auto_keeps_the_filter_on_near_duplicate_families(cuda, oracle):
    """A video's 1024 near-identical frames against rows near them: 1e-3 of the pairs are candidates, nearly all of them
    matches.  All needles of a tile flag the same accumulator column, and the unfiltered kernel would pay for every match:
    ALGO_AUTO must stay on the prefix filter (measured 28e12 vs 10.5e12 pairs/s at 10M rows, profiles/r02y)."""
    import torch
    rng = np.random.default_rng(78)
    n_rows, q = 400_000, 1024
    rows, needles = _clustered(rng, n_rows, q, 600, oracle)
    want = np.sort(oracle.matcher_find(rows, needles, 31), order=["row", "needle"])
    assert len(want) > 200_000
    db = cuda.CudaVectorDatabase(rows)
    m = cuda.CudaMatcher(db, 31)
    nd = torch.from_numpy(needles).cuda()
    for _ in range(4):
        assert np.array_equal(_find_device_np(cuda, m, nd, q, cuda.ALGO_AUTO, 1 << 20), want)
        assert db.auto_state[0] is False
    for algo in (cuda.ALGO_TENSOR_PAIRED, cuda.ALGO_TENSOR_FULL):        # candidate-list overflow -> inline verification
        assert np.array_equal(_find_device_np(cuda, m, nd, q, algo, 1 << 20), want)
    db.close()


@pytest.mark.parametrize("q", [1, 8, 9, 16, 17, 32, 33, 64])
@pytest.mark.parametrize("n_rows", [1, 127, 128, 129, 256, 257, 148 * 256 * 3 + 5, 300_001])
def test_rows_in_m_tensor_scan(cuda, oracle, q, n_rows):
    """ALGO_TENSOR_ROWS (matcher_tc1.cu): 128-row tiles, T per work unit, units dealt round-robin to (CTA, worker group);
    plants sit on tile and unit edges, in the last (partial) unit, in every TMEM lane quadrant and against every needle
    slot of the four instantiations (N = 8, 16, 32, 64); needles with no / all prefix bits set move the bias constant to
    its extremes (threshold - popc122 = 31 .. -91); an all-zero tail row (the padding of a partial tile) must not match
    a needle with fewer than `threshold` set prefix bits."""
    rng = np.random.default_rng(n_rows * 7 + q)
    rows = oracle.synth_db(n_rows, first_row=3)
    needles = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
    needles[0, :16] = 0                                   # popc(prefix) = 0: bias + 31
    if q > 1:
        needles[q - 1, :16] = 0xFF                        # popc(prefix) = 128: bias - 91
    if q > 2:
        needles[1, :16] = np.packbits(np.arange(128) % 5 == 0)     # few set bits, some under the six constant positions
    edges = [r for r in (0, 31, 32, 63, 64, 95, 96, 127, 128, 255, 256, 511, 512, 148 * 128 - 1, 148 * 128, 148 * 256 - 1,
                         148 * 256, 3 * 148 * 256 - 1, 3 * 148 * 256, n_rows - 2, n_rows - 1) if 0 <= r < n_rows]
    for k, r in enumerate(edges):
        lo = int(rng.integers(0, 20))
        rows[r] = _flip_in(needles[k % q], lo, int(rng.integers(0, 33 - lo)), rng)     # 0 .. 32 bits away
    if n_rows > 1000:
        for k, r in enumerate(rng.choice(n_rows, size=300, replace=False)):
            rows[r] = _flip_in(needles[k % q], int(rng.integers(0, 34)), int(rng.integers(0, 8)), rng)
        # differences confined to the six hash bits the filter does not see (bits 0, 4 of bytes 0..2): candidates, and
        # matches or not by the full distance alone
        for k, r in enumerate(rng.choice(n_rows, size=24, replace=False)):
            row = _flip_in(needles[k % q], 29, 0, rng)
            row[:3] ^= np.uint8(0x11)
            rows[r] = row
    want = sorted(match_tuples(oracle.matcher_find(rows, needles, 31)))
    db = cuda.CudaVectorDatabase(rows, global_row_offset=10**10)
    m = cuda.CudaMatcher(db, 31)
    want = [(r + 10**10, n, d) for r, n, d in want]
    assert _find_device(cuda, m, needles, cuda.ALGO_TENSOR_ROWS) == want
    assert _find_device(cuda, m, needles, cuda.ALGO_AUTO) == want
    for thr in (0, 30, 40) + ((128,) if n_rows < 1000 else ()):       # 128: half of all pairs match
        w = sorted((r + 10**10, n, d) for r, n, d in match_tuples(oracle.matcher_find(rows, needles, thr)))
        assert _find_device(cuda, cuda.CudaMatcher(db, thr), needles, cuda.ALGO_TENSOR_ROWS, cap=1 << 18) == w, thr
    db.close()


@pytest.mark.parametrize("q", [1, 8, 13, 32])
@pytest.mark.parametrize("n_rows", [1, 129, 148 * 256 + 77, 1_200_003])
def test_mailbox_query_one_rank(cuda, oracle, q, n_rows):
    """The peer-memory query with a world of ONE rank (this GPU is root and only peer): the fused form -- post, wait,
    rows-in-M scan, publish and collect inside one launch (matcher_tc1.cu, MailboxHook) -- and the separate-launch form
    must both answer what the oracle says, for host and device needles, query after query on the same mailbox (the
    two match counters alternate; the finished-CTA counter is reset by the last CTA)."""
    import torch
    rng = np.random.default_rng(n_rows + q)
    rows = oracle.synth_db(n_rows, first_row=11)
    db = cuda.CudaVectorDatabase(rows)
    mb = cuda.PeerMailbox(0, 0, 1)
    mb.connect([mb.handle])
    queries = []
    for k in range(6):
        needles = rng.integers(0, 256, size=(q, 32), dtype=np.uint8)
        r = int(rng.integers(0, n_rows))
        if k == 1:
            needles[q - 1] = flip_bits(rows[r], 31, rng)          # exactly at the threshold, last needle slot
        if k == 2:
            needles[0] = flip_bits(rows[r], 32, rng)              # one bit too far
        if k == 3:
            needles[q // 2] = flip_bits(rows[n_rows - 1], 0, rng)     # last row (partial tile)
        if k == 5:
            needles[0] = flip_bits(rows[0], 7, rng)
        queries.append((needles, int(len(oracle.matcher_find(rows, needles, 31)["row"]) > 0)))
    assert queries[1][1] == 1 and queries[3][1] == 1 and queries[0][1] == 0
    for mode in (cuda.MAILBOX_AUTO, cuda.MAILBOX_FUSED, cuda.MAILBOX_SEPARATE, cuda.MAILBOX_FUSED):
        mb.set_mode(mode)
        for k, (needles, want) in enumerate(queries):
            src = needles if k % 2 == 0 else torch.from_numpy(needles).cuda()
            assert mb.scan(db, src, q, 31, root=0) == want, (mode, k)
    # other thresholds through the fused form (the filter is exact at every threshold; 128 matches half of all pairs)
    mb.set_mode(cuda.MAILBOX_FUSED)
    for thr in (0, 40, 128):
        for needles, _ in queries[:4]:
            want = int(len(oracle.matcher_find(rows, needles, thr)["row"]) > 0)
            assert mb.scan(db, needles, q, thr, root=0) == want, thr
    mb.close()
    db.close()


@pytest.mark.parametrize("algo_name", ["ALGO_AUTO", "ALGO_TILED", "ALGO_EXACT", "ALGO_TENSOR", "ALGO_TENSOR_FULL", "ALGO_TENSOR_ROWS"])
def test_find_device_split_needle_indices(cuda, oracle, algo_name):
    """yume_pdq_cuda_matcher_find_device_split: the batch is a longer needle list with a hole cut out (every rank's
    all-gathered needles but the caller's own); reported indices must be those of the un-cut list, in every kernel."""
    import torch
    algo = getattr(cuda, algo_name)
    rng = np.random.default_rng(99)
    n_rows = 70_003
    q_full, lo, hi = (48, 16, 32) if algo_name == "ALGO_TENSOR_ROWS" else (700, 300, 450)     # the hole: [lo, hi)
    rows = oracle.synth_db(n_rows, first_row=5)
    full = rng.integers(0, 256, size=(q_full, 32), dtype=np.uint8)
    for k, r in enumerate(rng.choice(n_rows, size=200, replace=False)):
        rows[r] = flip_bits(full[(k * 7) % q_full], int(rng.integers(0, 34)), rng)
    want = [(r, n, d) for r, n, d in match_tuples(oracle.matcher_find(rows, full, 31)) if not lo <= n < hi]
    assert len(want) > 100 and any(n >= hi for _, n, _ in want) and any(n < lo for _, n, _ in want)
    cut = np.concatenate([full[:lo], full[hi:]])
    db = cuda.CudaVectorDatabase(rows)
    m = cuda.CudaMatcher(db, 31)
    out = torch.zeros((1 << 12, 16), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    m.find_device(torch.from_numpy(cut).cuda(), cut.shape[0], out, 1 << 12, cnt, algo=algo, needle_base=1000, split_at=lo, gap=hi - lo)
    torch.cuda.synchronize()
    got = out[: int(cnt.item())].cpu().numpy().view(cuda.MATCH_DTYPE).reshape(-1)
    assert sorted(match_tuples(got)) == sorted((r, n + 1000, d) for r, n, d in want)
    db.close()
