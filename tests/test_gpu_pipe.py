"""`yume-pdq-cuda pipe` end to end on the GPU: raw luma8 / rgb8 / rgba8 frames in, the reference CLI's
wire formats out (src/main.rs:448-823, 1090-1300), compared byte for byte with the oracle's hashes rendered by the
independent format restatement in helpers.py; EOF semantics and exit codes of the reference."""
import os
import subprocess
import time

import numpy as np
import pytest

from helpers import wire_stream

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "yume_pdq_b200", "yume-pdq-cuda")


def _pipe(args, data, timeout=120):
    r = subprocess.run([CLI, "pipe", *args], input=data, capture_output=True, timeout=timeout)
    return r.returncode, r.stdout, r.stderr


@pytest.fixture(scope="module")
def streams(cuda, oracle):
    assert cuda.device_count() > 0 and os.path.exists(CLI)
    out = {}
    for name, ch in (("luma8", 1), ("rgb8", 3), ("rgba8", 4)):
        px = oracle.synth_pixels(19, ch, first_frame=900)
        want = oracle.hash_all_dihedrals(oracle.cvt_to_luma8f(px, ch))
        out[name] = (px.tobytes(), want)
    return out


@pytest.mark.parametrize("fmt", ["luma8", "rgb8", "rgba8"])
def test_pipe_matches_oracle(streams, fmt):
    data, (h, q, t, p) = streams[fmt]
    for spec, dihedrals in (("hex+lf", False), ("q+hex+lf", True), ("raw", True), ("q+raw", False), ("bin+crlf", False),
                            ("q+BIN+nul", True), ("diagnostic", True), ("diagnostic+lf", False)):
        args = ["--input-format", fmt, "-f", spec, "--batch", "8"] + (["--dihedrals"] if dihedrals else [])
        rc, out, err = _pipe(args, data)
        assert rc == 0, err
        assert out == wire_stream(spec, h, q, t, p, dihedrals), (fmt, spec, dihedrals)
    # the default format is `bin` without separator, upper-case input format names are accepted (main.rs:1298-1300)
    rc, out, _ = _pipe(["--input-format", fmt.upper()], data)
    assert rc == 0 and out == wire_stream("bin", h, q, t, p, False)


def test_batch_size_does_not_change_the_stream(streams):
    data, (h, q, t, p) = streams["luma8"]
    want = wire_stream("q+hex+lf", h, q, t, p, True)
    for batch in (1, 3, 19, 256):
        assert _pipe(["-f", "q+hex+lf", "--dihedrals", "--batch", str(batch)], data)[1] == want


def test_readme_example(aaa_luma, kat):
    """ffmpeg ... -pix_fmt gray8 -f rawvideo - | yume-pdq pipe -f q+hex+lf  ->  100.000:58f8f0ce... (main.rs:150-154)"""
    assert np.array_equal(aaa_luma, np.round(aaa_luma)) and aaa_luma.min() >= 0 and aaa_luma.max() <= 255
    rc, out, err = _pipe(["-f", "q+hex+lf"], aaa_luma.astype(np.uint8).tobytes())
    assert rc == 0, err
    assert out == b"100.000:" + kat["readme_aaa"]["hash"].encode() + b"\n"


def test_eof_semantics(streams):
    """EOF between frames or on a row boundary ends the stream silently; inside a row it is exit code 5 after the
    complete frames were written (main.rs:494-506, 1036-1042)."""
    data, (h, q, t, p) = streams["rgb8"]
    frame = 512 * 512 * 3
    want3 = wire_stream("hex+lf", h[:3], q[:3], t[:3])
    rc, out, err = _pipe(["--input-format", "rgb8", "-f", "hex+lf"], data[:3 * frame + 10 * 512 * 3])
    assert rc == 0 and out == want3
    rc, out, err = _pipe(["--input-format", "rgb8", "-f", "hex+lf"], data[:3 * frame + 10 * 512 * 3 + 7])
    assert rc == 5 and out == want3 and b"Unexpected EOF" in err
    rc, out, err = _pipe(["-f", "hex+lf"], b"")
    assert rc == 0 and out == b""


def test_files_and_stats(streams, tmp_path):
    data, (h, q, t, p) = streams["rgba8"]
    src, dst = tmp_path / "in.rgba", tmp_path / "out.txt"
    src.write_bytes(data)
    r = subprocess.run([CLI, "pipe", "-i", str(src), "-o", str(dst), "--input-format", "rgba8", "-f", "HEX+lf", "--stats"],
                       capture_output=True, timeout=120)
    assert r.returncode == 0 and r.stdout == b""
    assert dst.read_bytes() == wire_stream("HEX+lf", h, q, t)


def test_slow_producer_gets_each_hash_immediately(streams):
    """A batch closes as soon as the input has no more bytes ready, so a frame-at-a-time producer sees every hash
    before it sends the next frame (the reference flushes per frame, main.rs:812)."""
    data, (h, q, t, p) = streams["luma8"]
    frame = 512 * 512
    proc = subprocess.Popen([CLI, "pipe", "-f", "hex+lf"], stdin=subprocess.PIPE, stdout=subprocess.PIPE)
    try:
        for k in range(4):
            proc.stdin.write(data[k * frame:(k + 1) * frame])
            proc.stdin.flush()
            t0 = time.time()
            line = proc.stdout.readline()
            assert line == wire_stream("hex+lf", h[k:k + 1], q[k:k + 1], t[k:k + 1])
            assert time.time() - t0 < 30
        proc.stdin.close()
        assert proc.wait(timeout=60) == 0
    finally:
        proc.kill()
