"""GPU parity of the fused input conversion through the C ABI: LUMA8 / RGB8 / RGBA8 frames hashed
directly must equal -- bit for bit -- the oracle's cvt (mod.rs:303-340, main.rs:508-510) followed by the oracle's
f32 hash path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FORMATS = [("LUMA8", 1, False), ("RGB8", 3, False), ("RGB8", 3, True), ("RGBA8", 4, False), ("RGBA8", 4, True)]


@pytest.fixture(scope="module")
def kernel(cuda):
    assert cuda.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return cuda.CudaKernel(0)


def _fmt(cuda, name, fma):
    return getattr(cuda, "PIXEL_" + name) | (cuda.PIXEL_FMA if fma else 0)


def _same(got, want):
    want_h, want_q, want_t, want_p = want
    assert (np.asarray(got["hashes"]) == want_h).all()
    assert np.array_equal(np.asarray(got["pdqf"]).view(np.uint32), want_p.view(np.uint32))
    assert np.array_equal(np.asarray(got["quality"]), want_q)
    assert np.array_equal(np.asarray(got["thresholds"]), want_t)


@pytest.mark.parametrize("name,channels,fma", FORMATS)
def test_fused_conversion_bit_exact(cuda, kernel, oracle, name, channels, fma):
    import torch
    n = 24
    fmt = _fmt(cuda, name, fma)
    shape = (n, 512, 512) if channels == 1 else (n, 512, 512, channels)
    dev = torch.empty(shape, dtype=torch.uint8, device="cuda:0")
    cuda.synth_pixels(dev, n, fmt, first_frame=300)
    torch.cuda.synchronize()
    host = oracle.synth_pixels(n, channels, first_frame=300)
    assert np.array_equal(dev.cpu().numpy(), host)                 # generators agree exactly
    luma = oracle.cvt_to_luma8f(host, channels, fma=fma)
    want = oracle.hash_all_dihedrals(luma)
    got = kernel.hash_batch(dev, want_pdqf=True, pixel_format=fmt)
    torch.cuda.synchronize()
    _same({k: v.cpu().numpy() for k, v in got.items()}, want)
    # host frames (staged in chunks) give the same
    _same(kernel.hash_batch(host, want_pdqf=True, pixel_format=fmt), want)
    # and so does the two-step path: standalone conversion, then the f32 hasher
    cvt = {1: kernel.cvt_luma8_to_luma8f, 3: kernel.cvt_rgb8_to_luma8f, 4: kernel.cvt_rgba8_to_luma8f}[channels]
    luma_gpu = cvt(host) if channels == 1 else cvt(host, fma=fma)
    assert np.array_equal(luma_gpu.view(np.uint32), luma.view(np.uint32))
    _same(kernel.hash_batch(luma_gpu, want_pdqf=True), want)


def test_random_bytes_and_alpha_ignored(cuda, kernel, oracle):
    rng = np.random.default_rng(5)
    rgba = rng.integers(0, 256, size=(3, 512, 512, 4), dtype=np.uint8)
    want = oracle.hash_all_dihedrals(oracle.cvt_to_luma8f(rgba, 4))
    _same(kernel.hash_batch(rgba, want_pdqf=True, pixel_format=cuda.PIXEL_RGBA8), want)
    other_alpha = rgba.copy()
    other_alpha[..., 3] = rng.integers(0, 256, size=other_alpha.shape[:-1], dtype=np.uint8)
    _same(kernel.hash_batch(other_alpha, want_pdqf=True, pixel_format=cuda.PIXEL_RGBA8), want)
    _same(kernel.hash_batch(np.ascontiguousarray(rgba[..., :3]), want_pdqf=True, pixel_format=cuda.PIXEL_RGB8), want)
    # extremes: all 0 and all 255 (luma of white must be what the scalar order gives, not exactly 255)
    flat = np.zeros((2, 512, 512, 3), dtype=np.uint8)
    flat[1] = 255
    _same(kernel.hash_batch(flat, want_pdqf=True, pixel_format=cuda.PIXEL_RGB8),
          oracle.hash_all_dihedrals(oracle.cvt_to_luma8f(flat, 3)))


def test_reference_rgba_kat_and_cvt_edges(cuda, kernel, oracle):
    # x86.rs:1510-1556
    px = np.array([[255, 0, 0, 128], [0, 255, 0, 128], [0, 0, 255, 128], [255, 255, 255, 128]], dtype=np.uint8)
    row = np.tile(px, (128, 1))
    itu = np.array(cuda.RGB8_TO_LUMA8_TABLE_ITU, dtype=np.float32)
    expect = np.tile(np.array([itu[0] * 255, itu[1] * 255, itu[2] * 255, 255], dtype=np.float32), 128)
    for fma in (False, True):
        got = kernel.cvt_rgba8_to_luma8f(row, fma=fma)
        assert np.abs(got - expect).max() < 1e-4
        assert np.array_equal(got, oracle.cvt_to_luma8f(row, 4, fma=fma))
    # ragged sizes (tails), unaligned device views, custom coefficients
    import torch
    rng = np.random.default_rng(2)
    for n_px in (1, 3, 5, 513, 4099):
        rgb = rng.integers(0, 256, size=(n_px, 3), dtype=np.uint8)
        assert np.array_equal(kernel.cvt_rgb8_to_luma8f(rgb), oracle.cvt_to_luma8f(rgb, 3))
        assert np.array_equal(kernel.cvt_rgb8_to_luma8f(rgb, coeff=(0.2126, 0.7152, 0.0722), fma=True),
                              oracle.cvt_to_luma8f(rgb, 3, coeff=(0.2126, 0.7152, 0.0722), fma=True))
        l8 = rng.integers(0, 256, size=(n_px,), dtype=np.uint8)
        assert np.array_equal(kernel.cvt_luma8_to_luma8f(l8), l8.astype(np.float32))
    t = torch.from_numpy(rng.integers(0, 256, size=(1001, 3), dtype=np.uint8)).cuda()
    view = t[1:]                                                   # 3-byte offset: unaligned device pointer
    got = kernel.cvt_rgb8_to_luma8f(view.contiguous() if not view.is_contiguous() else view)
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy(), oracle.cvt_to_luma8f(view.cpu().numpy(), 3))
    assert kernel.cvt_rgb8_to_luma8f(np.zeros((0, 3), dtype=np.uint8)).shape == (0,)


def test_bad_formats_rejected(cuda, kernel):
    frames = np.zeros((1, 512, 512), dtype=np.uint8)
    with pytest.raises(ValueError):
        kernel.hash_batch(frames, pixel_format=7)
    with pytest.raises(ValueError):
        kernel.hash_batch(frames, pixel_format=cuda.PIXEL_LUMA8 | cuda.PIXEL_FMA)   # FMA has no meaning for LUMA8
    with pytest.raises(ValueError):
        kernel.hash_batch(frames, pixel_format=cuda.PIXEL_RGB8)                      # byte size mismatch
    hashes = np.zeros((1, 8, 32), dtype=np.uint8)
    rc = cuda.lib.yume_pdq_cuda_hash_batch_pixels(0, frames.ctypes.data, 9, 1, hashes.ctypes.data, None, None, None,
                                                  None)
    assert rc == cuda.E_INVALID


@pytest.mark.parametrize("name,channels", [("LUMA8", 1), ("RGBA8", 4)])
def test_screen_pixels(cuda, kernel, oracle, name, channels):
    """screen_batch_pixels: 8-bit host frames -> hashes + match list, against oracle cvt + hash + match."""
    fmt = _fmt(cuda, name, False)
    n = 150                                                        # > 128: two pipeline chunks
    frames = oracle.synth_pixels(n, channels, first_frame=40)
    want_h = oracle.hash_all_dihedrals(oracle.cvt_to_luma8f(frames, channels))[0]
    needles = want_h.reshape(-1, 32)
    rows = oracle.synth_db(60_000)
    rng = np.random.default_rng(9)
    where = rng.choice(rows.shape[0], size=40, replace=False)
    for k, r in enumerate(where):
        rows[r] = needles[(37 * k) % needles.shape[0]]
        rows[r, k % 32] ^= np.uint8(0x0F)                          # 4 bits away
    db = cuda.CudaVectorDatabase(rows)
    m = cuda.CudaMatcher(db, 31)
    hashes = np.zeros((n, 8, 32), dtype=np.uint8)
    got = m.screen(frames, hashes=hashes, pixel_format=fmt)
    assert (hashes == want_h).all()
    want = oracle.matcher_find(rows, needles, 31)
    assert len(got) == len(want) >= 40 and (got == want).all()
    db.close()
