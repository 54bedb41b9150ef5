"""The C++ host driver (reve_b200/host): PNG codec on CPU, segment contract on the GPU."""
import os
import subprocess

import numpy as np
import pytest

from oracle import srvgg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "reve_b200", "host", "reve-upscale")


@pytest.fixture(scope="module")
def exe(lib):
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(os.path.join(ROOT, "reve_b200", "host", "reve_upscale.cpp")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "reve_b200", "host")])
    return EXE


def test_png_codec_round_trip(exe, tmp_path):
    import cv2
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    gray = rng.integers(0, 256, (20, 31), dtype=np.uint8)
    rgba = rng.integers(0, 256, (16, 9, 4), dtype=np.uint8)
    cases = {"rgb": rgb[:, :, ::-1], "gray": gray, "rgba": rgba}
    for name, arr in cases.items():
        src, dst = str(tmp_path / f"{name}.png"), str(tmp_path / f"{name}_out.png")
        assert cv2.imwrite(src, arr, [cv2.IMWRITE_PNG_COMPRESSION, 6])
        subprocess.check_call([exe, "--png-roundtrip", src, dst])
        got = cv2.imread(dst, cv2.IMREAD_COLOR)
        want = cv2.imread(src, cv2.IMREAD_COLOR)           # gray -> RGB, alpha dropped: what upstream loads
        assert got is not None and np.array_equal(got, want), name
    bad = str(tmp_path / "bad.png")
    open(bad, "wb").write(b"not a png")
    assert subprocess.call([exe, "--png-roundtrip", bad, str(tmp_path / "x.png")], stderr=subprocess.DEVNULL) == 1


def test_cli_argument_errors(exe, tmp_path):
    assert subprocess.call([exe], stderr=subprocess.DEVNULL) == 2
    assert subprocess.call([exe, "-i", str(tmp_path), "-o", str(tmp_path / "o"), "-s", "5"], stderr=subprocess.DEVNULL) == 2
    assert subprocess.call([exe, "-i", str(tmp_path / "missing"), "-o", str(tmp_path / "o")], stderr=subprocess.DEVNULL) == 1
    # an empty segment directory is not an error (nothing to do)
    (tmp_path / "empty").mkdir()
    assert subprocess.call([exe, "-i", str(tmp_path / "empty"), "-o", str(tmp_path / "o2")]) == 0


@pytest.mark.gpu
def test_cli_is_a_drop_in_for_the_spawned_upscaler(exe, tmp_path):
    """Same argv as reve-shared/src/lib.rs:134-147; progress = stderr lines containing 'done'."""
    import cv2
    indir, outdir = tmp_path / "tmp_frames" / "3", tmp_path / "out_frames" / "3"
    indir.mkdir(parents=True)
    frames = [srvgg.synthetic_frame(96, 54, 20 + i, "edges") for i in range(4)]
    for i, f in enumerate(frames):
        assert cv2.imwrite(str(indir / f"frame{i + 1:08d}.png"), f[:, :, ::-1])
    r = subprocess.run([exe, "-i", str(indir), "-o", str(outdir), "-n", "realesr-animevideov3-x2", "-s", "3",
                        "-f", "png", "-v", "-m", str(tmp_path / "no_models")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert sum("done" in l for l in r.stderr.splitlines()) == 4
    wts = srvgg.make_weights(3, 1234)          # the driver's random-init fallback seed
    for i, f in enumerate(frames):
        got = cv2.imread(str(outdir / f"frame{i + 1:08d}.png"), cv2.IMREAD_COLOR)[:, :, ::-1]
        par = srvgg.parity(got, srvgg.upscale(f, wts, tile=200, prepad=10))
        assert par["within1"] >= 0.999 and par["psnr"] >= 50, par


@pytest.mark.gpu
def test_raw_rgb24_stream_mode(exe, tmp_path):
    """Raw-frame staging: rgb24 rawvideo in -> rgb24 out, frames in order."""
    w, h, n, s = 160, 90, 11, 2
    frames = [srvgg.synthetic_frame(w, h, 50 + i, "random" if i % 2 else "edges") for i in range(n)]
    src, dst = tmp_path / "in.rgb", tmp_path / "out.rgb"
    src.write_bytes(b"".join(f.tobytes() for f in frames))
    r = subprocess.run([exe, "--raw", f"{w}x{h}", "-i", str(src), "-o", str(dst), "-s", str(s), "-v",
                        "-m", str(tmp_path / "no_models")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert sum("done" in l for l in r.stderr.splitlines()) == n
    out = np.frombuffer(dst.read_bytes(), np.uint8).reshape(n, h * s, w * s, 3)
    wts = srvgg.make_weights(s, 1234)
    for i in (0, 3, 10):
        par = srvgg.parity(out[i], srvgg.upscale(frames[i], wts, tile=200, prepad=10))
        assert par["within1"] >= 0.999 and par["psnr"] >= 50, (i, par)
    # stdin / stdout pipes
    p = subprocess.run([exe, "--raw", f"{w}x{h}", "-i", "-", "-o", "-", "-s", str(s), "-m", str(tmp_path / "no_models")],
                       input=src.read_bytes(), capture_output=True)
    assert p.returncode == 0 and p.stdout == dst.read_bytes()
    # a truncated stream is an error, not silently dropped
    (tmp_path / "bad.rgb").write_bytes(src.read_bytes()[:-5])
    assert subprocess.call([exe, "--raw", f"{w}x{h}", "-i", str(tmp_path / "bad.rgb"), "-o", str(tmp_path / "o.rgb"),
                            "-m", str(tmp_path / "no_models")], stderr=subprocess.DEVNULL) == 1


def test_pix_fmt_needs_raw_mode(exe, tmp_path):
    (tmp_path / "d").mkdir()
    assert subprocess.call([exe, "-i", str(tmp_path / "d"), "-o", str(tmp_path / "o"), "--pix-fmt", "yuv420p10le"],
                           stderr=subprocess.DEVNULL) == 2
    assert subprocess.call([exe, "--raw", "8x8", "-i", "-", "-o", "-", "--pix-fmt", "nv12"], stderr=subprocess.DEVNULL) == 2


@pytest.mark.gpu
def test_raw_stream_yuv420p10le_output(exe, tmp_path):
    """rgb24 in, planar yuv420p10le out (what `ffmpeg -f rawvideo -pix_fmt yuv420p10le`
    reads), bit-exact against the colour oracle applied to the RGB stream of the same run."""
    from oracle import colour
    w, h, n, s = 160, 90, 5, 2
    frames = [srvgg.synthetic_frame(w, h, 70 + i, "random" if i % 2 else "edges") for i in range(n)]
    src, rgb, yuv = tmp_path / "in.rgb", tmp_path / "out.rgb", tmp_path / "out.yuv"
    src.write_bytes(b"".join(f.tobytes() for f in frames))
    base = [exe, "--raw", f"{w}x{h}", "-i", str(src), "-s", str(s), "-m", str(tmp_path / "no_models")]
    assert subprocess.call(base + ["-o", str(rgb)]) == 0
    assert subprocess.call(base + ["-o", str(yuv), "--pix-fmt", "yuv420p10le"]) == 0
    W, H = w * s, h * s
    out_rgb = np.frombuffer(rgb.read_bytes(), np.uint8).reshape(n, H, W, 3)
    raw = np.frombuffer(yuv.read_bytes(), "<u2")
    assert raw.size == n * (W * H * 3 // 2)
    per = raw.reshape(n, W * H * 3 // 2)
    for i in range(n):
        y, u, v = colour.rgb_to_yuv420p10(out_rgb[i], 601)
        assert np.array_equal(per[i, :W * H].reshape(H, W), y)
        assert np.array_equal(per[i, W * H:W * H * 5 // 4].reshape(H // 2, W // 2), u)
        assert np.array_equal(per[i, W * H * 5 // 4:].reshape(H // 2, W // 2), v)
