"""GPU parity of the pipelined hash+match call (yume_pdq_cuda_screen_batch) against the CPU oracle:
the flow of examples/end_to_end.rs:224-256 for a batch -- hash every frame, emit its 8 dihedral
hashes, find them in the haystack.  Bit-exact hashes, identical match sets, host and device frames,
chunk boundaries (n not a multiple of the 128-frame pipeline step), list-capacity growth."""
import numpy as np
import pytest

from helpers import flip_bits, match_tuples

pytestmark = pytest.mark.gpu


def _case(oracle, n_frames, n_rows, seed, n_plant):
    rng = np.random.default_rng(seed)
    frames = oracle.synth_frames(n_frames, first_frame=1000 + seed, kind=1)
    wh, wq, _, _ = oracle.hash_all_dihedrals(frames)
    needles = wh.reshape(-1, 32)
    rows = oracle.synth_db(n_rows, first_row=5 * seed)
    for r in rng.choice(n_rows, size=n_plant, replace=False):
        j = int(rng.integers(0, needles.shape[0]))
        rows[r] = flip_bits(needles[j], int(rng.integers(0, 36)), rng)
    want = match_tuples(oracle.matcher_find(rows, needles, 31))
    return frames, wh, wq, rows, want


@pytest.mark.parametrize("n_frames", [1, 5, 130, 300])
def test_screen_host_frames(cuda, oracle, n_frames):
    frames, wh, wq, rows, want = _case(oracle, n_frames, 50_021, n_frames, 40)
    assert len(want) >= 20
    db = cuda.CudaVectorDatabase(rows)
    m = cuda.CudaMatcher(db, 31)
    hashes = np.zeros((n_frames, 8, 32), dtype=np.uint8)
    quality = np.zeros((n_frames, 8), dtype=np.float32)
    got = m.screen(frames, hashes=hashes, quality=quality)
    assert np.array_equal(hashes, wh) and np.array_equal(quality, wq)
    assert match_tuples(got) == want          # sorted by (row, needle = 8 * frame + dihedral)
    # no outputs wanted besides the matches
    assert match_tuples(m.screen(frames)) == want
    db.close()


def test_screen_device_frames_and_capacity(cuda, oracle):
    import torch
    frames, wh, wq, rows, want = _case(oracle, 64, 20_000, 7, 300)
    db = cuda.CudaVectorDatabase(rows, global_row_offset=1 << 33)
    m = cuda.CudaMatcher(db, 31)
    fd = torch.from_numpy(frames).cuda()
    hd = torch.zeros((64, 8, 32), dtype=torch.uint8, device="cuda")
    got = m.screen(fd, hashes=hd, cap=16)        # forces the E_CAPACITY retry in the Python mirror
    assert np.array_equal(hd.cpu().numpy(), wh)
    assert match_tuples(got) == [(r + (1 << 33), j, d) for r, j, d in want]
    # C ABI capacity contract: first `cap` entries in sorted order + total
    import ctypes as C
    out = np.zeros(5, dtype=cuda.MATCH_DTYPE)
    k = C.c_size_t(0)
    rc = cuda.lib.yume_pdq_cuda_screen_batch(db._h, cuda._ptr(frames), 64, 31, None, None, cuda._ptr(out), 5,
                                             C.byref(k), None)
    assert rc == cuda.E_CAPACITY and k.value == len(want)
    assert match_tuples(out) == [(r + (1 << 33), j, d) for r, j, d in want[:5]]
    # empty batch
    rc = cuda.lib.yume_pdq_cuda_screen_batch(db._h, None, 0, 31, None, None, None, 0, C.byref(k), None)
    assert rc == 0 and k.value == 0
    db.close()


def test_screen_equals_hash_then_find(cuda, oracle):
    """screen == hash_batch + matcher_find on the same inputs (the two-call path the reference's
    example takes), at a size where the needle batch goes through the tensor-core kernel."""
    import torch
    n = 256
    fd = torch.empty((n, 512, 512), dtype=torch.float32, device="cuda")
    cuda.synth_frames(fd, n, first_frame=77, kind=1)
    k = cuda.CudaKernel(0)
    h = k.hash_batch(fd)["hashes"]
    h = h.cpu().numpy() if hasattr(h, "cpu") else h
    rng = np.random.default_rng(3)
    rows = oracle.synth_db(300_000)
    for r in rng.choice(rows.shape[0], size=200, replace=False):
        rows[r] = flip_bits(h.reshape(-1, 32)[int(rng.integers(0, 8 * n))], int(rng.integers(0, 34)), rng)
    db = cuda.CudaVectorDatabase(rows)
    m = cuda.CudaMatcher(db, 31)
    two_call = match_tuples(m.find_all(h.reshape(-1, 32)))
    assert len(two_call) > 100
    assert match_tuples(m.screen(fd)) == two_call
    assert match_tuples(m.screen(fd.cpu().numpy())) == two_call
    assert two_call == match_tuples(oracle.matcher_find(rows, h.reshape(-1, 32), 31))
    db.close()
