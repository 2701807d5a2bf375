"""CPU checks of the oracle's input conversion: cvt_rgb8_to_luma8f / cvt_rgba8_to_luma8f
(src/kernel/mod.rs:303-340), the AVX2 rounding variant (x86.rs:686-753), pinned on the reference's own
unit test vector (x86.rs:1510-1556), and the 8-bit synthetic frame generator."""
import numpy as np

F32 = np.float32


def test_reference_rgba_kat(oracle):
    # x86.rs:1510-1556: red, green, blue, white cycling over a 512-pixel RGBA row, alpha = 128;
    # expected coeff * 255 (white: 255) within 1e-4.
    px = np.array([[255, 0, 0, 128], [0, 255, 0, 128], [0, 0, 255, 128], [255, 255, 255, 128]], dtype=np.uint8)
    row = np.tile(px, (128, 1))
    itu = oracle.luma_table_itu
    assert itu.tobytes() == np.array([0.299, 0.587, 0.114], dtype=F32).tobytes()   # constants.rs:23
    expect = np.tile(np.array([itu[0] * F32(255), itu[1] * F32(255), itu[2] * F32(255), F32(255)], dtype=F32), 128)
    for fma in (False, True):
        got = oracle.cvt_to_luma8f(row, 4, fma=fma)
        assert got.shape == (512,)
        assert np.abs(got - expect).max() < 1e-4
    # the RGB8 variant sees the same pixels without alpha
    assert np.array_equal(oracle.cvt_to_luma8f(row[:, :3].copy(), 3), oracle.cvt_to_luma8f(row, 4))


def test_scalar_order_and_fma_variant(oracle):
    rng = np.random.default_rng(11)
    px = rng.integers(0, 256, size=(4096, 3), dtype=np.uint8)
    cr, cg, cb = (F32(c) for c in oracle.luma_table_itu)
    r, g, b = (px[:, i].astype(F32) for i in range(3))
    want = (cr * r + cg * g) + cb * b                     # numpy f32: one rounding per operation
    assert want.dtype == F32
    assert np.array_equal(oracle.cvt_to_luma8f(px, 3), want)
    # AVX2 rounding: r*cr rounded, then two fused multiply-adds (exact product, one rounding each)
    v = (r * cr).astype(np.float64)
    v = (g.astype(np.float64) * np.float64(cg) + v).astype(F32).astype(np.float64)   # 24x24-bit product is exact in f64
    v = (b.astype(np.float64) * np.float64(cb) + v).astype(F32)
    got = oracle.cvt_to_luma8f(px, 3, fma=True)
    # the f64 emulation double-rounds in rare ties; allow those to differ by one ulp, require most to be equal
    assert (got == v).mean() > 0.999 and np.abs(got - v).max() <= np.spacing(F32(255))
    assert (got != want).any()                            # the two roundings really differ somewhere
    # custom coefficients (the trait's const generics)
    half = oracle.cvt_to_luma8f(px, 3, coeff=(0.5, 0.25, 0.25))
    assert np.array_equal(half, (F32(0.5) * r + F32(0.25) * g) + F32(0.25) * b)


def test_luma8_is_a_cast(oracle):
    px = np.arange(256, dtype=np.uint8)
    assert np.array_equal(oracle.cvt_to_luma8f(px, 1), px.astype(F32))


def test_synth_pixels_layout(oracle):
    l8 = oracle.synth_pixels(2, 1, first_frame=7)
    rgb = oracle.synth_pixels(2, 3, first_frame=7)
    rgba = oracle.synth_pixels(2, 4, first_frame=7)
    assert l8.shape == (2, 512, 512) and rgb.shape == (2, 512, 512, 3) and rgba.shape == (2, 512, 512, 4)
    assert np.array_equal(rgb[..., 0], l8) and np.array_equal(rgba[..., :3], rgb)
    assert l8.max() <= 243 and l8.std() > 20                 # smooth content, not a constant
    assert len(np.unique(rgba[..., 3])) == 256               # alpha is noise (and must be ignored)
    # frame f of a batch equals frame 0 of a batch starting at f (counter based)
    assert np.array_equal(oracle.synth_pixels(1, 3, first_frame=8)[0], rgb[1])
    # hashing the converted frames gives usable quality
    q = oracle.hash_get_threshold(oracle.cvt_to_luma8f(rgb[0], 3))[0]
    assert q > 0.5
