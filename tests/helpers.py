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


# ---- wire formats of the reference CLI's `pipe` (src/main.rs:668-778, 1106-1133), restated independently of the
# C++ front end: exact decimal arithmetic instead of printf ----
DIHEDRAL_ORDER = [0x01000001, 0xFF000001, 0xFF0000FF, 0x010000FF, 0x00FFFF00, 0x00010100, 0x00FF0100, 0x0001FF00]
SEPARATORS = {"lf": b"\n", "LF": b"\n", "nl": b"\n", "NL": b"\n", "crlf": b"\r\n", "CRLF": b"\r\n", "cr": b"\r",
              "CR": b"\r", "nul": b"\0", "NUL": b"\0", "null": b"\0", "NULL": b"\0"}


def rust_fixed(v, decimals: int) -> str:
    """Rust's `{:.N}` of an f32: the exact binary value, rounded half-to-even at N decimals."""
    from decimal import ROUND_HALF_EVEN, Decimal
    v = np.float32(v)
    if np.isnan(v):
        return "NaN"
    if np.isinf(v):
        return "inf" if v > 0 else "-inf"
    d = Decimal(float(v)).quantize(Decimal(1).scaleb(-decimals), rounding=ROUND_HALF_EVEN)
    s = format(d, "f")
    if np.signbit(v) and not s.startswith("-"):
        s = "-" + s
    return s


def dihedral_pdqf(pdqf, d):
    """PDQF state when dihedral d (index into DIHEDRAL_ORDER) is reported (lib.rs:391-417)."""
    p = np.asarray(pdqf, dtype=np.float32).reshape(16, 16)
    t, c, r = bool(d & 4), bool((0xC6 >> d) & 1), bool((0x6C >> d) & 1)
    out = p.T.copy() if t else p.copy()
    if c:
        out[:, 0::2] = -out[:, 0::2]
    if r:
        out[0::2, :] = -out[0::2, :]
    return out


def wire_record(fmt: str, hash32, quality, threshold=0.0, star_threshold=0.0, dihedral=0, index=0, pdqf=None) -> bytes:
    """One record in the body format `fmt` (no separator)."""
    h = hash32 if isinstance(hash32, (bytes, bytearray)) else bytes(bytearray(np.asarray(hash32, dtype=np.uint8)))
    h = bytes(h)
    q100 = np.float32(quality) * np.float32(100.0)
    prefix = (rust_fixed(q100, 3) + ":").encode()
    bits = "".join(format(b, "08b") for b in h).encode()
    if fmt in ("raw", "Raw", "RAW"):
        return h
    if fmt in ("q+raw", "q+Raw", "q+RAW"):
        return np.float32(quality).astype("<f4").tobytes() + h
    if fmt == "hex":
        return h.hex().encode()
    if fmt == "HEX":
        return h.hex().upper().encode()
    if fmt in ("bin", "BIN"):
        return bits
    if fmt == "q+hex":
        return prefix + h.hex().encode()
    if fmt == "q+HEX":
        return prefix + h.hex().upper().encode()
    if fmt in ("q+bin", "q+BIN"):
        return prefix + bits
    assert fmt == "diagnostic"
    name = '"original"' if dihedral == 0 else "Dihedrals { packed: %d }" % dihedral
    lines = ["[%s]" % name, "Index: %d" % index, "Quality: " + rust_fixed(q100, 3),
             "Threshold: " + rust_fixed(threshold, 7), "Hash:", bits.decode(), "PDQF:"]
    p = np.asarray(pdqf, dtype=np.float32).reshape(16, 16)
    for i in range(16):
        lines.append("\t".join(("*" if p[i, j] > np.float32(star_threshold) else "") + rust_fixed(p[i, j], 5)
                               for j in range(16)))
    return ("\n".join(lines) + "\n").encode()


def wire_stream(spec: str, hashes, quality, thresholds, pdqf=None, dihedrals=False) -> bytes:
    """Expected `pipe -f spec [--dihedrals]` output for hash_all_dihedrals-style arrays (n x 8 x 32, n x 8, n x 8,
    n x 256)."""
    parts = spec.split("+")
    sep = b""
    if parts[-1] in SEPARATORS:
        sep = SEPARATORS[parts[-1]]
        parts = parts[:-1]
    body = "+".join(parts)
    out = bytearray()
    for k in range(len(hashes)):
        for d in range(8 if dihedrals else 1):
            out += wire_record(body, hashes[k][d], quality[k][d], thresholds[k][d], thresholds[k][0],
                               0 if d == 0 else DIHEDRAL_ORDER[d], d,
                               None if pdqf is None else (pdqf[k] if d == 0 else dihedral_pdqf(pdqf[k], d)))
            out += sep
    return bytes(out)
