"""Shared test helpers: planted-positive DB construction (the procedure of
matching/mod.rs:460-538 restated with our own seeded RNG, plus exact-distance
plants that pin the inclusive threshold boundary)."""
import numpy as np


def flip_bits(h: np.ndarray, k: int, rng) -> np.ndarray:
    """Copy of 32-byte hash `h` with exactly k DISTINCT bit positions flipped."""
    bits = np.unpackbits(h)
    pos = rng.choice(256, size=k, replace=False)
    bits[pos] ^= 1
    return np.packbits(bits)


def flip_bits_with_replacement(h: np.ndarray, k: int, rng) -> np.ndarray:
    """generate_positive_control, matching/mod.rs:479-494: k random flips that may
    collide, so the distance is <= k."""
    out = h.copy()
    for _ in range(k):
        b = int(rng.integers(0, 256))
        out[b // 8] ^= np.uint8(1 << (b % 8))
    return out


def plant(rows: np.ndarray, needles: np.ndarray, dists, rng):
    """Plant, for every needle and every distance in `dists`, one copy of the needle
    at that exact distance at a random distinct row.  Returns the list of
    (row, needle, dist) planted."""
    n = rows.shape[0]
    total = needles.shape[0] * len(dists)
    where = rng.choice(n, size=total, replace=False)
    planted = []
    k = 0
    for j in range(needles.shape[0]):
        for d in dists:
            r = int(where[k])
            k += 1
            rows[r] = flip_bits(needles[j], d, rng)
            planted.append((r, j, d))
    return planted


def match_tuples(m):
    return [(int(r), int(n), int(d)) for r, n, d in zip(m["row"], m["needle"], m["dist"])]
