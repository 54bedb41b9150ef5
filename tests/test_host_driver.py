"""The C++ host driver (reve_b200/host): PNG codec on CPU, segment contract on the GPU."""
import os
import subprocess

import numpy as np
import pytest

import reve_b200
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


def test_missing_model_files_fail_the_run(exe, tmp_path):
    """the driver used to print one warning, upscale with random weights and exit 0.  Upstream fails when
    models/<name>.param|.bin are missing; so does the drop-in, with a non-zero exit code (the reference's caller never
    looks at it, reve-shared/src/lib.rs:148-154, but no frame and no 'done' line is produced either)."""
    import cv2
    d = tmp_path / "seg"
    d.mkdir()
    assert cv2.imwrite(str(d / "frame00000001.png"), np.zeros((16, 16, 3), np.uint8))
    r = subprocess.run([exe, "-i", str(d), "-o", str(tmp_path / "out"), "-s", "2", "-v", "-m", str(tmp_path / "no_models")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "not found" in r.stderr and "done" not in r.stderr
    assert not any((tmp_path / "out").glob("*.png"))


def _video_temp(tmp, frame_count=4500, segment_size=1000, ratio=2, drop=()):
    st = reve_b200.VideoState.new("in.mkv", "out.mkv", frame_count, 23.976, segment_size, ratio)
    for i in drop:
        st.mark_done(i)
    (tmp / "video.temp").write_text(st.to_json())
    return st


def test_segment_scheduler_runs_every_segment_and_persists_a_set(exe, tmp_path):
    """SURVEY.md 8(f) rank 2, host logic on CPU (--schedule-only skips the GPU stage): the reference's loop
    (reve-cli/src/main.rs:172-350) for G workers over one queue; video.temp holds the SET of unfinished segments."""
    st = _video_temp(tmp_path)                      # 5 segments: 4 x 1000 + 499 (lib.rs:282-289)
    assert [n for _, n in st.segments] == [1000, 1000, 1000, 1000, 499]
    log = tmp_path / "export.log"
    r = subprocess.run([exe, "--segments", str(tmp_path), "--schedule-only", "-g", "0,1,2",
                        "--export-cmd", f"echo {{index}} {{size}} {{seek}} >> {log} && touch {{in_dir}}/frame00000001.png",
                        "--encode-cmd", "test -d {out_dir} && echo {fps} > {part}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # every segment exported with ITS OWN size (the reference sizes the export by queue position, lib.rs:99,117:
    # with two segments left the first would be exported with the last one's 499 frames) and its own seek (lib.rs:94-98)
    rows = sorted(tuple(l.split()) for l in log.read_text().splitlines())
    assert [(int(a), int(b)) for a, b, _ in rows] == [(0, 1000), (1, 1000), (2, 1000), (3, 1000), (4, 499)]
    assert float(rows[0][2]) == 0.0 and abs(float(rows[3][2]) - (3 * 1000 - 1) / 23.976) < 1e-3
    for i in range(5):
        assert (tmp_path / "video_parts" / f"{i}.mp4").read_text().strip() == "23.976/1"      # main.rs:302
        assert not (tmp_path / "tmp_frames" / str(i)).exists() and not (tmp_path / "out_frames" / str(i)).exists()
    done = reve_b200.VideoState.from_json((tmp_path / "video.temp").read_text())
    assert done.segments == [] and done.segment_count == 5 and done.frame_count == 4500
    assert sum("done on gpu" in l for l in r.stderr.splitlines()) == 5


def test_segment_scheduler_checks_exit_codes_and_resumes(exe, tmp_path):
    """The reference never looks at a child's exit status (lib.rs:120-126,163-170) and pops a segment from the queue
    when it is HANDED to the encoder (main.rs:340-343), repairing that by re-encoding one more segment on resume
    (main.rs:142-159).  Here a segment leaves video.temp only after its encode exited 0, a failure ends the run with
    exit code 1, and the resumed run redoes exactly what is still listed (stale part files deleted first)."""
    _video_temp(tmp_path, drop=(0,))                # segment 0 finished in an earlier run
    (tmp_path / "video_parts").mkdir()
    (tmp_path / "video_parts" / "0.mp4").write_text("finished earlier")
    (tmp_path / "video_parts" / "3.mp4").write_text("cut by the crash")
    bad = subprocess.run([exe, "--segments", str(tmp_path), "--schedule-only", "-g", "0,1",
                          "--encode-cmd", "test {index} -ne 2 && echo new > {part}"], capture_output=True, text=True)
    assert bad.returncode == 1 and "encode of segment 2 failed" in bad.stderr
    left = reve_b200.VideoState.from_json((tmp_path / "video.temp").read_text()).remaining()
    assert 2 in left and 0 not in left
    assert (tmp_path / "video_parts" / "0.mp4").read_text() == "finished earlier"       # finished parts are never touched
    for i in left:                                  # whatever is listed has no part file worth keeping
        p = tmp_path / "video_parts" / f"{i}.mp4"
        assert not p.exists() or p.read_text() != "cut by the crash"
    good = subprocess.run([exe, "--segments", str(tmp_path), "--schedule-only", "-g", "0",
                           "--encode-cmd", "echo resumed > {part}"], capture_output=True, text=True)
    assert good.returncode == 0, good.stderr
    assert reve_b200.VideoState.from_json((tmp_path / "video.temp").read_text()).segments == []
    for i in left:
        assert (tmp_path / "video_parts" / f"{i}.mp4").read_text().strip() == "resumed"
    # a resume file that is not the reference's schema is refused, not guessed at
    (tmp_path / "video.temp").write_text('{"path": "x"}')
    assert subprocess.call([exe, "--segments", str(tmp_path), "--schedule-only"], stderr=subprocess.DEVNULL) == 1


def test_video_temp_round_trips_between_python_and_the_driver(exe, tmp_path):
    """Same JSON schema as serde's `Video` (reve-shared/src/lib.rs:9-25) on both sides."""
    st = reve_b200.VideoState.new('dir with "quotes"\\and back\\slash.mkv', "o.mkv", 1440, 23.976, 1000, 3)
    (tmp_path / "video.temp").write_text(st.to_json())
    assert subprocess.call([exe, "--segments", str(tmp_path), "--schedule-only", "-g", "0"]) == 0
    back = reve_b200.VideoState.from_json((tmp_path / "video.temp").read_text())
    assert back.path == st.path and back.segments == [] and back.upscale_ratio == 3 and abs(back.frame_rate - 23.976) < 1e-6


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
                        "-f", "png", "-v", "-m", str(tmp_path / "no_models"), "--random-weights"], capture_output=True, text=True)
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
                        "-m", str(tmp_path / "no_models"), "--random-weights"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert sum("done" in l for l in r.stderr.splitlines()) == n
    out = np.frombuffer(dst.read_bytes(), np.uint8).reshape(n, h * s, w * s, 3)
    wts = srvgg.make_weights(s, 1234)
    for i in (0, 3, 10):
        par = srvgg.parity(out[i], srvgg.upscale(frames[i], wts, tile=200, prepad=10))
        assert par["within1"] >= 0.999 and par["psnr"] >= 50, (i, par)
    # stdin / stdout pipes
    p = subprocess.run([exe, "--raw", f"{w}x{h}", "-i", "-", "-o", "-", "-s", str(s), "-m", str(tmp_path / "no_models"), "--random-weights"],
                       input=src.read_bytes(), capture_output=True)
    assert p.returncode == 0 and p.stdout == dst.read_bytes()
    # a truncated stream is an error, not silently dropped
    (tmp_path / "bad.rgb").write_bytes(src.read_bytes()[:-5])
    assert subprocess.call([exe, "--raw", f"{w}x{h}", "-i", str(tmp_path / "bad.rgb"), "-o", str(tmp_path / "o.rgb"),
                            "-m", str(tmp_path / "no_models"), "--random-weights"], stderr=subprocess.DEVNULL) == 1


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
    base = [exe, "--raw", f"{w}x{h}", "-i", str(src), "-s", str(s), "-m", str(tmp_path / "no_models"), "--random-weights"]
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


@pytest.mark.gpu
def test_segment_scheduler_on_the_gpu_two_lanes(exe, tmp_path):
    """The running scheduler: 5 segments, two worker lanes (-g 0,0: two contexts on one
    device, each launching its chained kernels cooperatively), export / encode stages as shell templates, every output
    frame checked against the oracle, video.temp empty at the end."""
    import cv2
    w, h, s, per = 320, 96, 2, 3
    st = reve_b200.VideoState.new("in.mkv", "out.mkv", 5 * per, 24.0, per, s)
    st.segments = [(i, per) for i in range(5)]                  # plain sizes: the frames are written below, not by ffmpeg
    (tmp_path / "video.temp").write_text(st.to_json())
    frames = {}
    for i in range(5):
        d = tmp_path / "tmp_frames" / str(i)
        d.mkdir(parents=True)
        for k in range(per):
            f = srvgg.synthetic_frame(w, h, 100 * i + k, "edges" if k % 2 else "random")
            frames[(i, k)] = f
            assert cv2.imwrite(str(d / f"frame{k + 1:08d}.png"), f[:, :, ::-1])
    keep = tmp_path / "kept"
    keep.mkdir()
    import torch
    lanes = "0,1" if torch.cuda.device_count() >= 2 else "0,0"     # two GPUs when the box has them
    r = subprocess.run([exe, "--segments", str(tmp_path), "-g", lanes, "-v", "--random-weights", "-m", str(tmp_path / "none"),
                        "--encode-cmd", f"mkdir -p {keep}/{{index}} && cp {{out_dir}}/*.png {keep}/{{index}}/ && echo ok > {{part}}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert sum(" done" in l and "->" in l for l in r.stderr.splitlines()) == 5 * per      # what main.rs:269 counts
    assert reve_b200.VideoState.from_json((tmp_path / "video.temp").read_text()).segments == []
    wts = srvgg.make_weights(s, 1234)
    for (i, k), f in frames.items():
        got = cv2.imread(str(keep / str(i) / f"frame{k + 1:08d}.png"), cv2.IMREAD_COLOR)[:, :, ::-1]
        par = srvgg.parity(got, srvgg.upscale(f, wts, tile=200, prepad=10))
        assert par["within1"] >= 0.999 and par["psnr"] >= 50, (i, k, par)


def test_segment_scheduler_skips_the_empty_last_segment(exe, tmp_path):
    """frame_count % segment_size == 1 makes the reference's last segment ZERO frames long (lib.rs:282-289: remainder - 1;
    `reve_b200.last_segment_size(1001, 1000) == 0`).  There is nothing to export, upscale or encode for it: the
    scheduler drops it instead of running ffmpeg with `-vframes 0` and waiting for a part file that cannot exist."""
    st = _video_temp(tmp_path, frame_count=2001, segment_size=1000)
    assert [n for _, n in st.segments] == [1000, 1000, 0]
    r = subprocess.run([exe, "--segments", str(tmp_path), "--schedule-only", "-g", "0,1",
                        "--export-cmd", "test {size} -gt 0", "--encode-cmd", "test {size} -gt 0 && echo ok > {part}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "segment 2 holds no frames" in r.stderr
    assert sorted(p.name for p in (tmp_path / "video_parts").iterdir()) == ["0.mp4", "1.mp4"]
    assert reve_b200.VideoState.from_json((tmp_path / "video.temp").read_text()).segments == []
