"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle.

Acceptance bar (BASELINE.json north_star): >= 99.9 % of u8 pixels within +-1 LSB and PSNR >= 50 dB
against the fp32 oracle evaluated with identical tile / pre-pad semantics.  Activations are stored
as fp16 on the device, so intermediate features are compared with a tolerance of 8e-3 + 8e-3*|ref| (twice the largest
error measured, tests/helpers.py).
"""
import ctypes as C
import glob
import os
import threading

import numpy as np
import pytest

import reve_b200
from helpers import feature_report, needed_rows, oracle_canvas
from oracle import srvgg

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
WITHIN1, PSNR = 0.999, 50.0


def check(out, ref):
    par = srvgg.parity(out, ref)
    assert par["within1"] >= WITHIN1 and par["psnr"] >= PSNR, par
    return par


def test_native_library_is_loaded_and_sees_the_gpu(lib):
    n = C.c_int()
    assert lib.reve_device_count(C.byref(n)) == 0 and n.value >= 1
    loaded = open("/proc/self/maps").read()
    assert "libreve_cuda.so" in loaded


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_vectors(path):
    g = np.load(path)
    scale, seed, tile, prepad = int(g["scale"]), int(g["seed"]), int(g["tile"]), int(g["prepad"])
    frame = g["frame"]
    model = reve_b200.Model.random(scale, seed)
    with reve_b200.Upscaler(model, frame.shape[1], frame.shape[0], tile=tile, prepad=prepad) as up:
        out = up.upscale(frame)
    assert out.shape == g["out"].shape
    check(out, g["out"])


@pytest.mark.parametrize("w,h,scale,tile,grid,pairs", [
    (100, 30, 2, 0, "4", "chain4"),   # chained body layers: one chain of 4 CTAs (layer j -> j+1 through the L2 scratch rings)
    (300, 40, 2, 0, "8", "chain4"),   # two chains, streams spanning several strips of 120 columns
    (300, 200, 2, 0, None, "chain4"), # 37 chains
    (200, 150, 2, 64, None, "chain4"),# tiles: gap rows/columns inside the extended rows of the inner layers
    (137, 91, 3, 50, None, "chain2"), # chains of two layers (strips of 124 columns), x3
    (500, 300, 2, 200, "6", "chain2"),# 3 chains of 2, needed-row lists with several segments per stream
    (100, 30, 2, 0, "1", False),      # one CTA, two streams walking the whole strip: every bank rotation, many ring laps
    (300, 40, 2, 0, "2", False),      # streams spanning two strips (several segments per stream)
    (300, 200, 2, 0, None, False),    # 148 CTAs, 2-3 rows per stream
    (200, 150, 2, 64, None, False),   # several tiles: gap rows/columns, reflect at every border
    (137, 91, 3, 50, None, False),    # ragged sizes, x3 (N padded to 32)
    (150, 90, 4, 0, "3", False),      # x4 (N = 48)
    (300, 40, 2, 0, "2", True),       # one CTA pair (tcgen05 cta_group::2), 4 lock-step streams over 3 strips
    (200, 150, 2, 64, None, True),    # 74 pairs, tiles
    (500, 300, 2, 200, "6", True),    # 3 pairs, streams of unequal segment counts (padding steps)
])
def test_per_layer_features_and_output(w, h, scale, tile, grid, pairs):
    opts = dict(debug_grid=int(grid) if grid else 0,
                debug_flags=reve_b200.DBG_CTA_PAIRS if pairs is True else 0,
                layers_per_launch=int(pairs[5:]) if isinstance(pairs, str) else 1)
    wts = srvgg.make_weights(scale, 1234)
    frame = srvgg.synthetic_frame(w, h, 5, "random")
    model = reve_b200.Model.random(scale, 1234)
    with reve_b200.Upscaler(model, w, h, tile=tile, prepad=10, ring_depth=2, **opts) as up:
        for layer in (1, 2, 3, 8, 9, 10, 13, 17):
            dev = up.debug_features(frame, layer)
            ref = oracle_canvas(frame, wts, tile, 10, layer)
            assert dev.shape == ref.shape
            rows = needed_rows(h, scale, tile, 10, layer)   # late layers skip rows no kept pixel depends on
            assert rows.all() == (layer <= 8)
            rep = feature_report(dev, ref, rows)
            assert rep["bad_frac"] == 0.0, (layer, rep)
        out = up.upscale(frame)
    check(out, srvgg.upscale(frame, wts, tile=tile, prepad=10))


def test_one_hot_weights_pin_tap_and_channel_layout():
    """A single non-zero weight per layer: any mix-up of (ky, kx, ci, co) or of the PixelShuffle
    order moves the response somewhere else."""
    scale = 2
    wts = srvgg.make_weights(scale, 1)
    for k in range(18):
        wts.conv_w[k][...] = 0
        wts.conv_b[k][...] = 0
    rng = np.random.default_rng(3)
    wts.conv_w[0][5, 1, 0, 2] = 1.0
    prev = 5
    for k in range(1, 17):
        co, ky, kx = int(rng.integers(0, 64)), int(rng.integers(0, 3)), int(rng.integers(0, 3))
        wts.conv_w[k][co, prev, ky, kx] = 1.0
        prev = co
    wts.conv_w[17][7, prev, 2, 0] = 0.5
    import tempfile
    d = tempfile.mkdtemp()
    srvgg.write_ncnn(wts, d + "/m.param", d + "/m.bin", fp16=True)
    model = reve_b200.Model.load_ncnn(d + "/m.param", d + "/m.bin")
    frame = srvgg.synthetic_frame(90, 70, 8, "random")
    with reve_b200.Upscaler(model, 90, 70, tile=0, prepad=10) as up:
        out = up.upscale(frame)
        feat = up.debug_features(frame, 17)
    ref = srvgg.upscale(frame, wts, tile=0, prepad=10)
    par = srvgg.parity(out, ref)
    assert par["within1"] == 1.0 and par["exact"] > 0.95, par   # x/255 passes through fp16 storage
    ref_feat = oracle_canvas(frame, wts, 0, 10, 17)
    assert feature_report(feat, ref_feat, needed_rows(70, 2, 0, 10, 17))["bad_frac"] == 0.0
    assert np.abs(ref_feat).max() > 0.1      # the probe actually lights something up


def test_edge_frames_and_strides():
    wts = srvgg.make_weights(2, 21)
    model = reve_b200.Model.random(2, 21)
    for (w, h, tile, prepad) in ((11, 11, 200, 10), (1, 1, 0, 0), (127, 1, 0, 0), (1, 130, 64, 0), (253, 12, 126, 10)):
        frame = srvgg.synthetic_frame(w, h, w + h, "random")
        with reve_b200.Upscaler(model, w, h, tile=tile, prepad=prepad) as up:
            out = up.upscale(frame)
            # strided (non-packed rows) input and output buffers
            big_in = np.zeros((h, w + 5, 3), np.uint8)
            big_in[:, :w] = frame
            big_out = np.full((h * 2, w * 2 + 7, 3), 77, np.uint8)
            up.submit(big_in[:, :w], big_out[:, :w * 2], 5)
            assert up.wait() == 5
        ref = srvgg.upscale(frame, wts, tile=tile, prepad=prepad)
        check(out, ref)
        assert np.array_equal(big_out[:, :w * 2], out)
        assert (big_out[:, w * 2:] == 77).all()          # bytes beyond the row are untouched


def test_ring_fifo_order_busy_and_empty():
    model = reve_b200.Model.random(2, 3)
    w, h = 96, 64
    wts = srvgg.make_weights(2, 3)
    frames = [srvgg.synthetic_frame(w, h, i, "edges" if i % 2 else "random") for i in range(7)]
    with reve_b200.Upscaler(model, w, h, tile=200, prepad=10, ring_depth=3) as up:
        with pytest.raises(reve_b200.ReveError) as e:
            up.wait()
        assert e.value.status == -8                       # REVE_E_EMPTY
        ins = [up.pinned((h, w, 3)) for _ in range(3)]
        outs = [up.pinned((h * 2, w * 2, 3)) for _ in range(3)]
        for i in range(3):
            ins[i][...] = frames[i]
            up.submit(ins[i], outs[i], 100 + i)
        with pytest.raises(reve_b200.ReveError) as e:
            up.submit(ins[0], outs[0], 999)
        assert e.value.status == -7                       # REVE_E_BUSY
        got = []
        for i in range(3):
            tag = up.wait()
            got.append((tag, outs[i].copy()))
        assert [t for t, _ in got] == [100, 101, 102]
        for i, (_, o) in enumerate(got):
            check(o, srvgg.upscale(frames[i], wts, tile=200, prepad=10))
        # pipelined helper over more frames than ring slots
        res = [np.empty((h * 2, w * 2, 3), np.uint8) for _ in frames]
        order = []
        assert up.upscale_many(frames, res, on_done=order.append) == len(frames)
        assert order == list(range(len(frames)))
        for f, o in zip(frames, res):
            check(o, srvgg.upscale(f, wts, tile=200, prepad=10))


def test_device_resident_path_matches_submit_path():
    import torch
    model = reve_b200.Model.random(2, 4)
    w, h, n = 160, 90, 3
    frames = np.stack([srvgg.synthetic_frame(w, h, 40 + i, "random") for i in range(n)])
    with reve_b200.Upscaler(model, w, h, tile=64, prepad=10) as up:
        d_in = torch.from_numpy(frames).cuda()
        d_out = torch.zeros((n, h * 2, w * 2, 3), dtype=torch.uint8, device="cuda")
        up.upscale_device(d_in.data_ptr(), d_out.data_ptr(), n)
        up.sync()
        dev = d_out.cpu().numpy()
        for i in range(n):
            assert np.array_equal(dev[i], up.upscale(frames[i]))
        prof = up.profile()
        # 16 body layers per frame whatever the launch structure (chained launches run 2 or 4 layers each)
        assert prof["frames"] == 2 * n and prof["body_layer_frames"] == 16 * 2 * n
        assert prof["launches_conv0"] == prof["launches_tail"]
        assert prof["launches_body"] in (16 * prof["launches_conv0"], 8 * prof["launches_conv0"], 4 * prof["launches_conv0"])


def test_results_are_deterministic_and_contexts_are_independent():
    model = reve_b200.Model.random(3, 6)
    w, h = 120, 80
    frame = srvgg.synthetic_frame(w, h, 9, "edges")
    results = {}

    def run(key):
        with reve_b200.Upscaler(model, w, h, tile=50, prepad=10) as up:
            a = up.upscale(frame)
            b = up.upscale(frame)
            results[key] = (a, b)

    ts = [threading.Thread(target=run, args=(k,)) for k in range(2)]   # one context per host thread
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    a0, b0 = results[0]
    a1, b1 = results[1]
    assert np.array_equal(a0, b0) and np.array_equal(a0, a1) and np.array_equal(a1, b1)
    check(a0, srvgg.upscale(frame, srvgg.make_weights(3, 6), tile=50, prepad=10))


def test_full_size_1080p_properties_and_oracle():
    """BASELINE.json configs[1] at full size: size-independent properties plus one oracle frame."""
    w, h, scale = 1920, 1080, 2
    wts = srvgg.make_weights(scale, 1234)
    model = reve_b200.Model.random(scale, 1234)
    frame = srvgg.synthetic_frame(w, h, 77, "edges")
    frame[::3, ::5] = srvgg.synthetic_frame(w, h, 78, "random")[::3, ::5]     # add texture
    with reve_b200.Upscaler(model, w, h, tile=200, prepad=10) as up:
        out = up.upscale(frame)
        assert np.array_equal(out, up.upscale(frame))                          # idempotent / deterministic
    assert out.shape == (h * scale, w * scale, 3)
    # tile locality: tile (0,0) only sees frame[0:210, 0:210]; a 210x210 frame has the same first tile
    crop = np.ascontiguousarray(frame[:210, :210])
    with reve_b200.Upscaler(model, 210, 210, tile=200, prepad=10) as up:
        small = up.upscale(crop)
    assert np.array_equal(out[:400, :400], small[:400, :400])
    # an interior tile equals the same window run as a stand-alone frame whose borders are real pixels
    # (tile (2,3): x in [600,800), y in [400,600)); compare through the oracle on that window only
    win = np.ascontiguousarray(frame[390:610, 590:810])                        # tile + 10 px of real neighbours
    x = (win.astype(np.float32) * np.float32(1 / 255.0)).transpose(2, 0, 1)
    ref_win = srvgg.quantise(srvgg.forward(x, wts)[:, 20:-20, 20:-20].transpose(1, 2, 0))
    check(out[800:1200, 1200:1600], ref_win)
    # whole frame against the oracle (about 4 s of CPU)
    check(out, srvgg.upscale(frame, wts, tile=200, prepad=10))


def test_whole_frame_mode_720p_x4_against_oracle_window():
    """BASELINE.json configs[2] geometry (1280x720 x4), whole-frame mode, checked on windows."""
    w, h, scale = 1280, 720, 4
    wts = srvgg.make_weights(scale, 5)
    model = reve_b200.Model.random(scale, 5)
    frame = srvgg.synthetic_frame(w, h, 31, "random")
    with reve_b200.Upscaler(model, w, h, tile=0, prepad=10) as up:
        out = up.upscale(frame)
    assert out.shape == (h * 4, w * 4, 3)
    # receptive-field radius is 18 px: a window with a 20 px margin reproduces the interior exactly
    for (x0, y0) in ((300, 200), (1100, 560)):
        win = np.ascontiguousarray(frame[y0 - 20:y0 + 84, x0 - 20:x0 + 84])
        x = (win.astype(np.float32) * np.float32(1 / 255.0)).transpose(2, 0, 1)
        ref = srvgg.quantise(srvgg.forward(x, wts)[:, 80:-80, 80:-80].transpose(1, 2, 0))
        check(out[y0 * 4:(y0 + 64) * 4, x0 * 4:(x0 + 64) * 4], ref)
    # top-left corner: reflect-101 pre-pad of 10 px
    corner = srvgg.upscale(np.ascontiguousarray(frame[:120, :120]), wts, tile=0, prepad=10)
    check(out[:80 * 4, :80 * 4], corner[:80 * 4, :80 * 4])


def test_upscale_segment_directory_contract(tmp_path):
    """Mirror of Video::upscale_segment: frames in, same names out, one 'done' line per frame."""
    import io
    indir, outdir = tmp_path / "tmp_frames" / "0", tmp_path / "out_frames" / "0"
    indir.mkdir(parents=True)
    frames = [srvgg.synthetic_frame(80, 60, i, "edges") for i in range(5)]
    import cv2
    for i, f in enumerate(frames):
        assert cv2.imwrite(str(indir / f"frame{i + 1:08d}.png"), f[:, :, ::-1])
    log = io.StringIO()
    model = reve_b200.Model.random(2, 8)
    n = reve_b200.upscale_segment(str(indir), str(outdir), 2, model=model, progress=log)
    assert n == 5
    lines = [l for l in log.getvalue().splitlines() if "done" in l]     # what reve-cli/src/main.rs:269 counts
    assert len(lines) == 5
    wts = srvgg.make_weights(2, 8)
    for i, f in enumerate(frames):
        got = cv2.imread(str(outdir / f"frame{i + 1:08d}.png"), cv2.IMREAD_COLOR)[:, :, ::-1]
        check(got, srvgg.upscale(f, wts, tile=200, prepad=10))
    with pytest.raises(ValueError):
        reve_b200.upscale_segment(str(indir), str(tmp_path / "o2"), 3, model=model)   # scale / model mismatch


@pytest.mark.parametrize("w,h,scale,fmt,matrix", [
    (96, 64, 2, reve_b200.FMT_YUV420P10LE_BT601, 601),     # aligned fast path
    (96, 64, 4, reve_b200.FMT_YUV420P10LE_BT709, 709),
    (137, 91, 3, reve_b200.FMT_YUV420P10LE_BT601, 601),    # odd output size (411 x 273): edge replication, scalar path
    (51, 33, 2, reve_b200.FMT_YUV420P10LE_BT709, 709),     # even height, width not a multiple of 4
])
def test_yuv420p10le_output_is_bit_exact(w, h, scale, fmt, matrix):
    """the frame leaves the GPU as yuv420p10le; integer work, so bit-exact
    against oracle/colour.py applied to the RGB output of the same context."""
    from oracle import colour
    model = reve_b200.Model.random(scale, 9)
    frame = srvgg.synthetic_frame(w, h, 4, "edges")
    frame[::2, ::3] = srvgg.synthetic_frame(w, h, 5, "random")[::2, ::3]
    with reve_b200.Upscaler(model, w, h, tile=64, prepad=10, ring_depth=4) as up:
        rgb = up.upscale(frame)
        up.set_output_format(fmt)
        stride, nbytes = up.output_layout()
        cw, ch = (up.out_w + 1) // 2, (up.out_h + 1) // 2
        assert stride == 4 * cw and nbytes == stride * up.out_h + stride * ch
        y, u, v = up.upscale_yuv(frame)
        ry, ru, rv = colour.rgb_to_yuv420p10(rgb, matrix)
        assert np.array_equal(y, ry) and np.array_equal(u, ru) and np.array_equal(v, rv)
        # a wider pitch: bytes between the rows are untouched
        y2, u2, v2 = up.upscale_yuv(frame, out_stride=stride + 64)
        assert np.array_equal(y2, ry) and np.array_equal(u2, ru) and np.array_equal(v2, rv)
        raw = up._last_raw
        rows = raw[:(stride + 64) * up.out_h].reshape(up.out_h, stride + 64)
        assert (rows[:, stride:] == 0xAB).all()
        # several frames in flight (one batch), then back to RGB
        bufs = [np.zeros(nbytes, np.uint8) for _ in range(3)]
        for i, b in enumerate(bufs):
            up.submit_raw(frame, b, stride, i)
        assert [up.wait() for _ in range(3)] == [0, 1, 2]
        for b in bufs:
            assert np.array_equal(b[:stride * up.out_h].reshape(up.out_h, stride)[:, :2 * up.out_w].copy().view("<u2"), ry)
        assert up.profile()["launches_yuv"] >= 5
        up.set_output_format(reve_b200.FMT_RGB24)
        assert np.array_equal(up.upscale(frame), rgb)


def test_yuv_format_error_paths():
    model = reve_b200.Model.random(2, 9)
    with reve_b200.Upscaler(model, 40, 20, ring_depth=2) as up:
        with pytest.raises(reve_b200.ReveError) as e:
            up.set_output_format(7)
        assert e.value.status == -1
        up.set_output_format(reve_b200.FMT_YUV420P10LE_BT601)
        stride, nbytes = up.output_layout()
        buf = np.zeros(nbytes, np.uint8)
        frame = srvgg.synthetic_frame(40, 20, 1, "random")
        with pytest.raises(reve_b200.ReveError) as e:
            up.submit_raw(frame, buf, stride - 4)              # pitch too small
        assert e.value.status == -1
        with pytest.raises(reve_b200.ReveError) as e:
            up.submit_raw(frame, np.zeros(2 * nbytes, np.uint8), stride + 2)   # pitch not a multiple of 4
        assert e.value.status == -1
        up.submit_raw(frame, buf, stride, 1)
        with pytest.raises(reve_b200.ReveError) as e:
            up.set_output_format(reve_b200.FMT_RGB24)          # frames in flight
        assert e.value.status == -7
        assert up.wait() == 1


def _random_cases(n, seed):
    rng = np.random.default_rng(seed)
    cases = []
    for i in range(n):
        scale = int(rng.choice([2, 3, 4]))
        prepad = int(rng.choice([0, 1, 3, 10, 10]))
        w = int(rng.integers(max(2, prepad + 1), 260))
        h = int(rng.integers(max(2, prepad + 1), 200))
        tile = int(rng.choice([0, 32, 50, 64, 100, 200]))
        grid = rng.choice(["", "1", "2", "5", "37"])
        pairs = bool(rng.integers(0, 2))
        chain = str(rng.choice(["0", "2", "4", "4"]))        # chained body layers (ignored with pairs / grids below 4)
        cases.append((w, h, scale, tile, prepad, str(grid), pairs, chain, int(rng.integers(0, 1 << 30))))
    return cases


@pytest.mark.parametrize("w,h,scale,tile,prepad,grid,pairs,chain,seed", _random_cases(36, 2026),
                         ids=lambda v: str(v))
def test_random_geometries_against_the_oracle(w, h, scale, tile, prepad, grid, pairs, chain, seed):
    """Ragged sizes x tile sizes x pre-pads x scales x grid sizes x CTA pairs: every combination changes the
    stream / segment / needed-row structure the kernels walk (one row per stream, streams spanning strips,
    pre-pads shorter than the receptive field, a single CTA doing everything)."""
    opts = dict(debug_grid=int(grid) if grid else 0, debug_flags=reve_b200.DBG_CTA_PAIRS if pairs else 0,
                layers_per_launch=int(chain) if chain != "0" else 1)
    wts = srvgg.make_weights(scale, seed % 1000)
    model = reve_b200.Model.random(scale, seed % 1000)
    frames = [srvgg.synthetic_frame(w, h, seed + i, "random" if i % 2 else "edges") for i in range(3)]
    with reve_b200.Upscaler(model, w, h, tile=tile, prepad=prepad, ring_depth=3, **opts) as up:
        outs = [np.empty((h * scale, w * scale, 3), np.uint8) for _ in frames]
        for i, (f, o) in enumerate(zip(frames, outs)):      # three frames in flight = one stacked batch
            up.submit(f, o, i)
        assert [up.wait() for _ in frames] == [0, 1, 2]
        single = up.upscale(frames[1])                       # and a batch of one
    assert np.array_equal(single, outs[1])
    for f, o in zip(frames[:2], outs[:2]):
        check(o, srvgg.upscale(f, wts, tile=tile, prepad=prepad))


def test_device_resident_path_rgb_and_yuv():
    """reve_upscale_device (frames already in HBM, what bench.py's `value` times): same bytes as the staged
    path, in both output formats, for a frame count that is not a multiple of the launch batch."""
    import torch
    from oracle import colour
    w, h, s, n = 112, 72, 2, 6
    model = reve_b200.Model.random(s, 4)
    frames = np.stack([srvgg.synthetic_frame(w, h, 30 + i, "random" if i % 2 else "edges") for i in range(n)])
    with reve_b200.Upscaler(model, w, h, tile=50, prepad=10, ring_depth=4) as up:
        ref = [up.upscale(f) for f in frames]
        d_in = torch.from_numpy(frames).cuda()
        d_out = torch.zeros((n, h * s, w * s, 3), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        up.upscale_device(d_in.data_ptr(), d_out.data_ptr(), n)
        up.sync()
        got = d_out.cpu().numpy()
        for i in range(n):
            assert np.array_equal(got[i], ref[i]), i
        up.set_output_format(reve_b200.FMT_YUV420P10LE_BT709)
        stride, nbytes = up.output_layout()
        d_yuv = torch.zeros((n, nbytes), dtype=torch.uint8, device="cuda")
        up.upscale_device(d_in.data_ptr(), d_yuv.data_ptr(), n)
        up.sync()
        raw = d_yuv.cpu().numpy()
        W, H = w * s, h * s
        for i in range(n):
            y, u, v = colour.rgb_to_yuv420p10(ref[i], 709)
            planes = raw[i].view("<u2")
            assert np.array_equal(planes[:W * H].reshape(H, W), y)
            assert np.array_equal(planes[W * H:W * H * 5 // 4].reshape(H // 2, W // 2), u)
            assert np.array_equal(planes[W * H * 5 // 4:].reshape(H // 2, W // 2), v)


def test_chained_launches_are_race_free_under_load():
    """The chained body layers hand rows from SM to SM through global-memory rings guarded by flags (no cluster, no
    grid sync): a lost or early flag would show up as a frame that differs from the same frame computed a moment
    before.  400 full-size frames back to back (the GPU stays saturated, the power cap moves the clocks), every output
    compared on the device with the first pass, which itself is checked against the oracle elsewhere.  (This test caught
    a real one: rows announced with a relaxed store after cp.async.bulk.wait_group were read partly stale about once in
    1000 frames; tools/race_hunt.py is the long-running version.)"""
    import torch
    w, h, s, n = 1920, 1080, 2, 8
    model = reve_b200.Model.random(s, 11)
    assert reve_b200.launch_plan(w, h, s)["layers_per_launch"] == 4
    frames = np.stack([srvgg.synthetic_frame(w, h, 60 + i, "random" if i % 2 else "edges") for i in range(n)])
    with reve_b200.Upscaler(model, w, h, tile=200, prepad=10, ring_depth=8) as up:
        d_in = torch.from_numpy(frames).cuda()
        ref = torch.zeros((n, h * s, w * s, 3), dtype=torch.uint8, device="cuda")
        out = torch.zeros_like(ref)
        up.upscale_device(d_in.data_ptr(), ref.data_ptr(), n)
        up.sync()
        assert int(ref.max()) > 0
        for it in range(50):
            out.fill_(0)
            torch.cuda.synchronize()
            up.upscale_device(d_in.data_ptr(), out.data_ptr(), n)
            up.sync()
            assert torch.equal(out, ref), f"pass {it} differs from the first pass"
        prof = up.profile()
        assert prof["launches_body"] == 4 * prof["launches_conv0"]          # chains of 4 were what ran


# ----------------------------------------------------------------------------------------------------------------
# Round 2: every BASELINE.json geometry at full size under upstream's tile 200 / pre-pad 10, the fp16 range, two
# contexts on one device, the staged path under load, and what a caller sees after a kernel-side fault.
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w,h,scale,seed", [
    (1280, 720, 4, 5),     # BASELINE.json configs[2]: 7 x 4 tiles (6 full columns + 80 px, 3 full rows + 120 px), N = 48
    (960, 540, 3, 6),      # configs[3]: 5 x 3 tiles, x3 (N padded to 32, 9-byte output pixels)
])
def test_full_size_tile200_against_the_oracle(w, h, scale, seed):
    """Whole frames under upstream's own tile 200 / pre-pad 10 semantics, not the whole-frame
    variant: the complete oracle frame (a few seconds of CPU) plus the tile-locality property."""
    wts = srvgg.make_weights(scale, seed)
    model = reve_b200.Model.random(scale, seed)
    frame = srvgg.synthetic_frame(w, h, 300 + scale, "edges")
    frame[::2, ::3] = srvgg.synthetic_frame(w, h, 400 + scale, "random")[::2, ::3]
    with reve_b200.Upscaler(model, w, h, tile=200, prepad=10, ring_depth=4) as up:
        out = up.upscale(frame)
        outs = [np.empty_like(out) for _ in range(4)]          # and as one stacked batch of four
        for i, o in enumerate(outs):
            up.submit(frame, o, i)
        assert [up.wait() for _ in outs] == [0, 1, 2, 3]
        info = up.launch_info()
    assert all(np.array_equal(o, out) for o in outs)
    assert info["layers_per_launch"] == 4 and info["cooperative"]
    check(out, srvgg.upscale(frame, wts, tile=200, prepad=10))
    # tile (0,0) only sees frame[0:210, 0:210]
    with reve_b200.Upscaler(model, 210, 210, tile=200, prepad=10) as up:
        small = up.upscale(np.ascontiguousarray(frame[:210, :210]))
    assert np.array_equal(out[:200 * scale, :200 * scale], small[:200 * scale, :200 * scale])


def test_full_480p_real_frame_golden():
    """BASELINE.json configs[0] geometry: a whole 640x480 frame decoded from the reference's demo asset
    (reve-cli/assets/onepiece_demo.mp4, frame 60; stored by oracle/make_golden.py because /root/reference does not
    exist on the GPU box), x2, upstream tile 200 / pre-pad 10: 4 x 3 tiles, chains of 2 (launch plan for 480p)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "x2_tile200_real_onepiece_480p.npz"))
    frame = g["frame"]
    assert frame.shape == (480, 640, 3) and int(g["tile"]) == 200 and int(g["prepad"]) == 10
    model = reve_b200.Model.random(int(g["scale"]), int(g["seed"]))
    with reve_b200.Upscaler(model, 640, 480, tile=200, prepad=10, ring_depth=4) as up:
        out = up.upscale(frame)
        assert up.launch_info()["layers_per_launch"] == 2
    check(out, g["out"])
    with reve_b200.Upscaler(model, 640, 480, tile=200, prepad=10, shared_device=True) as up:   # single-layer launches
        assert up.launch_info() == {"layers_per_launch": 1, "batch": 3, "grid": 148, "cooperative": False}
        assert np.array_equal(up.upscale(frame), out)


def _scaled_weights(scale, seed, gain):
    """He-init weights with conv0 scaled up by `gain` and the last conv scaled down by it: every intermediate
    activation is `gain` times larger, the output branch keeps its size."""
    wts = srvgg.make_weights(scale, seed)
    wts.conv_w[0] = (wts.conv_w[0] * np.float32(gain)).astype(np.float16).astype(np.float32)
    wts.conv_b[0] = wts.conv_b[0] * np.float32(gain)
    for k in range(1, 17):
        wts.conv_b[k] = wts.conv_b[k] * np.float32(gain)
    wts.conv_w[17] = (wts.conv_w[17] / np.float32(gain)).astype(np.float16).astype(np.float32)
    return wts


def _model_of(wts):
    return reve_b200.Model.from_arrays(wts.scale, wts.conv_w, wts.conv_b, wts.slopes)


def test_fp16_range_stress_large_activations():
    """Trained weights are not available offline and the He-init family keeps |activation| < 4.  Here the same
    family runs with activations 1000 x larger (absmax 10^3..10^4, the upper decades of fp16): relative precision is
    scale-free, so the features must match the oracle's fp16-storage path to the same relative tolerance and the
    frame must meet the same bar against the fp32 oracle."""
    scale, gain = 2, 1000.0
    wts = _scaled_weights(scale, 77, gain)
    frame = srvgg.synthetic_frame(300, 200, 9, "random")
    with reve_b200.Upscaler(_model_of(wts), 300, 200, tile=0, prepad=10) as up:
        out = up.upscale(frame)
        feats = {layer: up.debug_features(frame, layer) for layer in (1, 4, 9, 17)}
    x = (srvgg.padded_tile(frame, 0, 0, 300, 200, 10).astype(np.float32) * np.float32(1 / 255.0)).transpose(2, 0, 1)
    _, ref16, _ = srvgg.forward(x, wts, taps=True, fp16_storage=True)
    peak = 0.0
    for layer, dev in feats.items():
        ref = ref16[layer - 1].transpose(1, 2, 0)
        rows = needed_rows(200, scale, 0, 10, layer)
        d, r = dev[rows], ref[rows]
        assert np.isfinite(d).all()
        peak = max(peak, float(np.abs(r).max()))
        bad = np.abs(d - r) > gain * 8e-3 + 8e-3 * np.abs(r)       # the tolerance of tests/helpers.py, scaled with the activations
        assert bad.mean() == 0.0, (layer, float(np.abs(d - r).max()), float(np.abs(r).max()))
    assert 1e3 < peak < 6e4, peak                    # the stress actually reached the upper fp16 decades, without overflow
    check(out, srvgg.upscale(frame, wts, tile=0, prepad=10))                       # fp32 oracle
    par16 = srvgg.parity(out, srvgg.upscale(frame, wts, tile=0, prepad=10, fp16_storage=True))
    assert par16["within1"] >= WITHIN1, par16


def test_fp16_overflow_saturates_like_the_oracle():
    """Convolution results beyond +-65504 become +-inf when rounded to fp16, and inf - inf in the next convolution is
    NaN.  Pinned here against the oracle's fp16-arithmetic variant (oracle/srvgg.py:forward, fp16_storage="arith": round
    the convolution result to fp16, then PReLU on fp16 operands -- ncnn's fp16 path and the device's): the CLASS of every
    feature value (finite / +inf / -inf / NaN) after layers 1 and 2 agrees, PReLU keeps NaN a NaN (ncnn:
    `x < 0 ? x * slope : x`), the final quantiser maps NaN to 0 and +-inf to 255 / 0 (oracle/srvgg.py:quantise), and
    nothing traps.  (The fp16-STORAGE variant -- PReLU in fp32, then round -- keeps values in (-65504/slope, -65504)
    finite; upstream's CPU and fp16-Vulkan paths differ in exactly this way, it only matters once a model overflows.)"""
    scale = 2
    wts = _scaled_weights(scale, 78, 4.0e4)      # conv0 weights stay below 65504, its outputs (std ~3e4) do not
    frame = srvgg.synthetic_frame(200, 120, 10, "random")
    with reve_b200.Upscaler(_model_of(wts), 200, 120, tile=0, prepad=10) as up:
        out = up.upscale(frame)
        f1, f2 = up.debug_features(frame, 1), up.debug_features(frame, 2)
        again = up.upscale(frame)
    assert np.array_equal(out, again)
    x = (srvgg.padded_tile(frame, 0, 0, 200, 120, 10).astype(np.float32) * np.float32(1 / 255.0)).transpose(2, 0, 1)
    y16, ref16, _ = srvgg.forward(x, wts, taps=True, fp16_storage="arith")

    def classes(a):
        return np.where(np.isnan(a), 3, np.where(np.isposinf(a), 1, np.where(np.isneginf(a), 2, 0)))

    c1, r1 = classes(f1), classes(ref16[0].transpose(1, 2, 0))
    c2, r2 = classes(f2), classes(ref16[1].transpose(1, 2, 0))
    assert (r1 == 1).mean() > 0.01 and (r1 == 2).mean() > 0.005 and (r2 == 3).mean() > 0.01   # overflows both ways, and NaN
    # values within one fp16 ulp of the overflow threshold may fall on either side (fp32 summation order)
    assert (c1 == r1).mean() >= 0.999, float((c1 == r1).mean())
    assert (c2 == r2).mean() >= 0.995, float((c2 == r2).mean())
    ref = srvgg.quantise(y16[:, 20:-20, 20:-20].transpose(1, 2, 0))
    agree = (np.abs(out.astype(int) - ref.astype(int)) <= 1).mean()
    assert agree >= 0.98, agree


def test_two_contexts_share_one_device_at_full_size():
    """Two 1080p contexts driven concurrently from two host threads on ONE device, one of them producing yuv420p10le,
    both through reve_submit / reve_wait with pinned buffers, 208 frames each.  Each context's chained kernel waits
    on flags written by its own CTAs only, and is launched cooperatively, so the two grids can never be half resident
    next to each other: every frame must be bit-identical to the context's first pass."""
    w, h, s, n, total = 1920, 1080, 2, 8, 208
    model = reve_b200.Model.random(s, 11)
    frames = [srvgg.synthetic_frame(w, h, 60 + i, "random" if i % 2 else "edges") for i in range(n)]
    errors, infos = [], []

    def run(yuv):
        try:
            with reve_b200.Upscaler(model, w, h, tile=200, prepad=10, ring_depth=8) as up:
                if yuv:
                    up.set_output_format(reve_b200.FMT_YUV420P10LE_BT601)
                stride, nbytes = up.output_layout()
                infos.append(up.launch_info())
                hin = [up.pinned((h, w, 3)) for _ in range(n)]
                hout = [up.pinned((nbytes,)) for _ in range(n)]
                for i in range(n):
                    hin[i][...] = frames[i]
                ref = []
                for i in range(n):                               # first pass, one frame at a time
                    up.submit_raw(hin[i], hout[i], stride, i)
                    up.wait()
                    ref.append(hout[i].copy())
                assert int(ref[0].max()) > 0
                inflight = 0
                for k in range(total):
                    if inflight == n:
                        t = up.wait()
                        inflight -= 1
                        if not np.array_equal(hout[t % n], ref[t % n]):
                            errors.append((yuv, t))
                    up.submit_raw(hin[k % n], hout[k % n], stride, k)
                    inflight += 1
                while inflight:
                    t = up.wait()
                    inflight -= 1
                    if not np.array_equal(hout[t % n], ref[t % n]):
                        errors.append((yuv, t))
        except Exception as e:      # noqa: BLE001
            errors.append((yuv, repr(e)))

    ts = [threading.Thread(target=run, args=(y,)) for y in (False, True)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:5]
    assert all(i["layers_per_launch"] == 4 and i["cooperative"] for i in infos), infos


def test_staged_path_is_race_free_over_1000_frames():
    """tools/race_hunt.py as a test, on the STAGED path: 1000 full-size frames through reve_submit / reve_wait with the
    H2D and D2H streams busy next to the chained kernels; every output compared with the first pass."""
    w, h, s, n, total = 1920, 1080, 2, 8, 1000
    model = reve_b200.Model.random(s, 12)
    frames = [srvgg.synthetic_frame(w, h, 80 + i, "random" if i % 2 else "edges") for i in range(n)]
    with reve_b200.Upscaler(model, w, h, tile=200, prepad=10, ring_depth=8) as up:
        hin = [up.pinned((h, w, 3)) for _ in range(n)]
        hout = [up.pinned((h * s, w * s, 3)) for _ in range(n)]
        for i in range(n):
            hin[i][...] = frames[i]
        ref = [up.upscale(frames[i]) for i in range(n)]
        bad, inflight = [], 0
        for k in range(total):
            if inflight == n:
                t = up.wait()
                inflight -= 1
                if not np.array_equal(hout[t % n], ref[t % n]):
                    bad.append(t)
                hout[t % n][::64] = 0                            # a stale buffer cannot pass for a fresh result
            up.submit(hin[k % n], hout[k % n], k)
            inflight += 1
        while inflight:
            t = up.wait()
            inflight -= 1
            if not np.array_equal(hout[t % n], ref[t % n]):
                bad.append(t)
        assert not bad, bad[:10]
        prof = up.profile()
        assert prof["launches_body"] == 4 * prof["launches_conv0"]          # chains of 4 were what ran


FAULT_SCRIPT = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
import reve_b200
from reve_b200 import _lib
w, h = 1920, 270
frame = np.random.default_rng(0).integers(0, 256, (h, w, 3), dtype=np.uint8)
model = reve_b200.Model.random(2, 1)
lib = _lib.load()
other = reve_b200.Upscaler(model, 96, 64)                       # a second, healthy context on the same device
good = other.upscale(frame[:64, :96].copy())
up = reve_b200.Upscaler(model, w, h, layers_per_launch=4, debug_flags=reve_b200.DBG_FAULT)
try:
    up.upscale(frame)
    print("NOFAULT")
    sys.exit(3)
except reve_b200.ReveError as e:
    print("STATUS", e.status)
    print("MSG", e)
try:                                                            # the error is sticky for every context of the device
    other.upscale(frame[:64, :96].copy())
    print("OTHER alive")
except reve_b200.ReveError as e:
    print("OTHER dead", e.status)
up._pinned, other._pinned = [], []                              # pinned buffers die with the device context: drop, do not free
up.close(); other.close()
rc = lib.reve_device_recover(0)
print("RECOVER", rc)
if rc == 0:
    with reve_b200.Upscaler(model, 96, 64) as fresh:
        again = fresh.upscale(frame[:64, :96].copy())
    print("SAME", bool(np.array_equal(again, good)))
else:
    print("REFUSED", lib.reve_last_error(None).decode())
'''


def test_failure_semantics_after_a_kernel_watchdog_trap():
    """include/reve_cuda.h 'Failure semantics': a kernel-side wait that exceeds its watchdog traps; the caller gets
    REVE_E_CUDA with the watchdog diagnostic, every context of that device in that PROCESS is dead (sticky CUDA error),
    reve_ctx_destroy stays safe, and other processes are unaffected.  reve_device_recover either brings the device back
    in-process (then a fresh context must reproduce the earlier result) or reports that the driver refuses (the B200 /
    driver 580 pool: cudaErrorDevicesUnavailable) -- the documented answer to which is a new worker process.
    Runs in a child process: the fault poisons the CUDA context of whoever provokes it."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", FAULT_SCRIPT % root], capture_output=True, text=True, timeout=300)
    out = r.stdout
    assert "STATUS -3" in out, (out, r.stderr[-2000:])
    assert "kernel watchdog: wait tag" in out, out      # whichever of the starved waits of chain 0 expired first
    assert "OTHER dead -3" in out, out
    assert ("RECOVER 0" in out and "SAME True" in out) or ("RECOVER -3" in out and "REFUSED" in out), (out, r.stderr[-2000:])
    # this process (another one, as far as the driver is concerned) still has a working device
    frame = srvgg.synthetic_frame(96, 64, 1, "random")
    with reve_b200.Upscaler(reve_b200.Model.random(2, 1), 96, 64) as up:
        check(up.upscale(frame), srvgg.upscale(frame, srvgg.make_weights(2, 1), tile=200, prepad=10))


@pytest.mark.parametrize("w,h,scale,tile", [(300, 200, 2, 64), (137, 91, 3, 50), (150, 90, 4, 0), (1920, 1080, 2, 200)])
def test_standalone_unpack_and_pack_kernels_are_bit_exact(w, h, scale, tile):
    """the u8 -> fp16 unpack (reflect-101 pre-pad, gaps, x/255) and the fp16 -> u8 pack (crop,
    clamp(floor(v*255 + 0.5))) as stand-alone kernels (reve_b200/csrc/pack.cu; the product path fuses both).  Byte /
    rounding work, so bit-exact against the oracle's restatement."""
    frame = srvgg.synthetic_frame(w, h, 7, "random")
    model = reve_b200.Model.random(scale, 2)
    with reve_b200.Upscaler(model, w, h, tile=tile, prepad=10) as up:
        cw, ch, src_x, out_x, src_y, out_y = reve_b200.geometry(w, h, scale, tile, 10)
        dev, _ = up.debug_unpack(frame)
        ref = np.zeros((ch, cw, 3), np.float32)
        yy, xx = np.where(src_y >= 0)[0], np.where(src_x >= 0)[0]
        ref[np.ix_(yy, xx)] = (frame[np.ix_(src_y[yy], src_x[xx])].astype(np.float32) * np.float32(1 / 255.0)
                               ).astype(np.float16).astype(np.float32)
        assert np.array_equal(dev, ref)
        # pack: a synthetic network output on the canvas (values beyond [0, 1], NaN and +-inf included)
        rng = np.random.default_rng(3)
        y = rng.normal(0.5, 0.6, (ch * scale, cw * scale, 3)).astype(np.float32)
        y[::7, ::5] = np.float32("nan")
        y[1::9, 2::11] = np.float32("inf")
        y[3::13, ::3] = -np.float32("inf")
        got, _ = up.debug_pack(y)
        keep_y, keep_x = np.where(out_y >= 0)[0], np.where(out_x >= 0)[0]
        rows = (keep_y[:, None] * scale + np.arange(scale)[None, :]).reshape(-1)
        cols = (keep_x[:, None] * scale + np.arange(scale)[None, :]).reshape(-1)
        want = srvgg.quantise(y.astype(np.float16).astype(np.float32)[np.ix_(rows, cols)])
        assert got.shape == want.shape == (h * scale, w * scale, 3)
        assert np.array_equal(got, want)


# ----------------------------------------------------------------------------------------------------------------
# The two first-conv kernels: the row-streaming one (conv0_rows.cu, default) and the K = 27 im2col one (conv0.cu,
# DBG_CONV0_IM2COL).  Same function, different accumulation order: features agree to an fp16 ulp, frames to 1 LSB, both
# meet the oracle.  Cases: one CTA (every ring lap), ragged widths, tiles (gap rows / columns), frames stacked four to a
# canvas (gap rows between frames), and a canvas whose geometry tables do not fit shared memory (global-table path).
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("w,h,scale,tile,grid", [
    (100, 30, 2, 0, 1), (129, 17, 2, 0, 0), (300, 200, 2, 64, 0), (137, 91, 3, 50, 0), (150, 90, 4, 0, 3), (11, 11, 4, 0, 0),
    (640, 480, 2, 200, 0),
])
def test_first_conv_kernels_agree_and_meet_the_oracle(w, h, scale, tile, grid):
    wts = srvgg.make_weights(scale, 77)
    model = reve_b200.Model.random(scale, 77)
    frame = srvgg.synthetic_frame(w, h, 12, "random")
    ref = oracle_canvas(frame, wts, tile, 10, 1)
    got = {}
    for name, flags in (("rows", 0), ("im2col", reve_b200.DBG_CONV0_IM2COL)):
        with reve_b200.Upscaler(model, w, h, tile=tile, prepad=10, ring_depth=2, debug_flags=flags, debug_grid=grid) as up:
            feat = up.debug_features(frame, 1)
            assert feat.shape == ref.shape
            assert feature_report(feat, ref, np.ones(ref.shape[0], bool))["bad_frac"] == 0.0, name
            got[name] = (feat, up.upscale(frame))
    d = np.abs(got["rows"][0] - got["im2col"][0])
    assert d.max() <= 2e-3 and (d == 0).mean() > 0.999, (float(d.max()), float((d == 0).mean()))
    assert np.abs(got["rows"][1].astype(int) - got["im2col"][1].astype(int)).max() <= 1
    check(got["rows"][1], srvgg.upscale(frame, wts, tile=tile, prepad=10))


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,tile", [(300, 200, 64), (48, 3000, 0)])
def test_first_conv_on_stacked_frames_and_global_tables(w, h, tile):
    """Four different frames per launch set (one canvas, gap rows between the frames).  48 x 3000 makes the stacked
    canvas 12 083 rows tall: its row tables (2 ints per row) exceed the 48 KB the kernels cache in shared memory, so the
    global-memory table path runs."""
    import torch
    scale, n = 2, 4
    wts = srvgg.make_weights(scale, 5)
    model = reve_b200.Model.random(scale, 5)
    frames = np.stack([srvgg.synthetic_frame(w, h, 30 + i, "random" if i % 2 else "edges") for i in range(n)])
    outs = {}
    for name, flags in (("rows", 0), ("im2col", reve_b200.DBG_CONV0_IM2COL)):
        with reve_b200.Upscaler(model, w, h, tile=tile, prepad=10, ring_depth=4, debug_flags=flags) as up:
            assert up.launch_info()["batch"] == n
            d_in = torch.from_numpy(frames).cuda()
            d_out = torch.zeros((n, h * scale, w * scale, 3), dtype=torch.uint8, device="cuda")
            up.upscale_device(d_in.data_ptr(), d_out.data_ptr(), n)
            up.sync()
            outs[name] = d_out.cpu().numpy()
    assert np.abs(outs["rows"].astype(int) - outs["im2col"].astype(int)).max() <= 1
    for i in (0, n - 1):
        check(outs["rows"][i], srvgg.upscale(frames[i], wts, tile=tile, prepad=10))


@pytest.mark.gpu
def test_4k_input_tile_locality_and_window_oracle():
    """The largest realistic input (3840x2160 -> 7680x4320; the canvas is 4 x a 1080p one, 209 tiles): no oracle run of
    the whole frame (a minute of CPU), but properties that do not depend on the size.  Deterministic; a tile's output
    depends only on the tile and its 10 px of real neighbours, so (a) the first 6 x 10 tiles are bit-identical to the
    same tiles of a 2010 x 1210 crop -- a different canvas, different strips, different CTA ranges -- and (b) an interior
    tile far from the origin meets the oracle run on its 220 x 220 window."""
    w, h, scale = 3840, 2160, 2
    wts = srvgg.make_weights(scale, 1234)
    model = reve_b200.Model.random(scale, 1234)
    frame = srvgg.synthetic_frame(w, h, 91, "edges")
    frame[::2, ::3] = srvgg.synthetic_frame(w, h, 92, "random")[::2, ::3]
    with reve_b200.Upscaler(model, w, h, tile=200, prepad=10, ring_depth=2) as up:
        out = up.upscale(frame)
        assert np.array_equal(out, up.upscale(frame))
    assert out.shape == (h * scale, w * scale, 3)
    crop = np.ascontiguousarray(frame[:1210, :2010])
    with reve_b200.Upscaler(model, 2010, 1210, tile=200, prepad=10, ring_depth=2) as up:
        small = up.upscale(crop)
    assert np.array_equal(out[:2400, :4000], small[:2400, :4000])
    win = np.ascontiguousarray(frame[1390:1610, 2990:3210])                    # tile (7, 15) + 10 px of real neighbours
    x = (win.astype(np.float32) * np.float32(1 / 255.0)).transpose(2, 0, 1)
    ref_win = srvgg.quantise(srvgg.forward(x, wts)[:, 20:-20, 20:-20].transpose(1, 2, 0))
    check(out[2800:3200, 6000:6400], ref_win)
