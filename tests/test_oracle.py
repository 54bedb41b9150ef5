"""CPU tests of the oracle itself: golden vectors, the independent C restatement, border rules.
The reference has no numeric test for this path; these pin the oracle."""
import glob
import os

import numpy as np
import pytest

from oracle import cref, srvgg

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_golden_files_present():
    assert len(GOLDEN) >= 8


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    w = srvgg.make_weights(int(g["scale"]), int(g["seed"]))
    out = srvgg.upscale(g["frame"], w, tile=int(g["tile"]), prepad=int(g["prepad"]))
    assert out.shape == g["out"].shape
    par = srvgg.parity(out, g["out"])
    # same code, same machine class: allow only fp32 summation-order noise of the conv backend
    assert par["within1"] == 1.0 and par["exact"] >= 0.999, par


SMALL = [p for p in GOLDEN if os.path.getsize(p) < 100_000][:4]   # the scalar C restatement needs ~1 us per MAC


@pytest.mark.parametrize("path", SMALL, ids=[os.path.basename(p)[:-4] for p in SMALL])
def test_c_restatement_matches_golden(path):
    g = np.load(path)
    w = srvgg.make_weights(int(g["scale"]), int(g["seed"]))
    out = cref.upscale(g["frame"], w, tile=int(g["tile"]), prepad=int(g["prepad"]))
    par = srvgg.parity(out, g["out"])
    assert par["within1"] == 1.0 and par["exact"] >= 0.999, par


def test_float64_agrees_with_float32():
    w = srvgg.make_weights(2, 5)
    f = srvgg.synthetic_frame(40, 30, 3, "edges")
    a = srvgg.upscale(f, w, tile=0)
    b = srvgg.upscale(f, w, tile=0, double=True)
    par = srvgg.parity(a, b)
    assert par["within1"] == 1.0 and par["exact"] > 0.999


def test_prng_is_deterministic_and_sane():
    a = srvgg.make_weights(3, 42)
    b = srvgg.make_weights(3, 42)
    c = srvgg.make_weights(3, 43)
    assert all(np.array_equal(x, y) for x, y in zip(a.conv_w, b.conv_w))
    assert not np.array_equal(a.conv_w[1], c.conv_w[1])
    assert [x.shape for x in a.conv_w] == [(64, 3, 3, 3)] + [(64, 64, 3, 3)] * 16 + [(27, 64, 3, 3)]
    # He-normal with PReLU gain: std = sqrt(2 / (1.0625 * fan_in))
    assert abs(a.conv_w[5].std() - np.sqrt(2 / (1.0625 * 576))) < 2e-3
    assert all(np.array_equal(x.astype(np.float16).astype(np.float32), x) for x in a.conv_w)
    assert all((s > 0.19).all() and (s < 0.31).all() for s in a.slopes)
    # splitmix64 known answer (seed 0, stream offset folded in): first outputs are stable
    z = srvgg._splitmix64(0, 0, 2)
    assert z.dtype == np.uint64 and int(z[0]) != int(z[1])


def test_reflect101_matches_upstream_rule():
    n = 7
    idx = np.arange(-6, 13)
    ref = []
    for i in idx:
        x = abs(int(i))
        x = (n - 1) - abs(x - (n - 1))
        ref.append(x)
    assert srvgg.reflect101(idx, n).tolist() == ref
    assert srvgg.reflect101(np.array([-1, 0, n - 1, n]), n).tolist() == [1, 0, n - 1, n - 2]


def test_tile_not_smaller_than_frame_equals_whole_frame():
    w = srvgg.make_weights(2, 9)
    f = srvgg.synthetic_frame(30, 22, 1, "random")
    assert np.array_equal(srvgg.upscale(f, w, tile=0), srvgg.upscale(f, w, tile=64))


def test_tiles_are_independent_functions_of_their_padded_window():
    """Upstream semantics: a tile's output depends only on the tile + 10 px of real neighbours."""
    w = srvgg.make_weights(2, 9)
    f = srvgg.synthetic_frame(70, 40, 2, "random")
    full = srvgg.upscale(f, w, tile=32, prepad=10)
    g = f.copy()
    g[:, 60:] = 255 - g[:, 60:]          # change pixels more than 10 px right of tile column 0 (x < 32)
    other = srvgg.upscale(g, w, tile=32, prepad=10)
    assert np.array_equal(full[:, :32 * 2], other[:, :32 * 2])
    assert not np.array_equal(full[:, 64 * 2:], other[:, 64 * 2:])


def test_pixelshuffle_channel_order_one_hot():
    """y[c, Y*s+i, X*s+j] = r[c*s*s + i*s + j, Y, X] (ncnn PixelShuffle mode 0)."""
    for s in (2, 3, 4):
        w = srvgg.make_weights(s, 1)
        for k in range(18):
            w.conv_w[k][...] = 0
            w.conv_b[k][...] = 0
        # identity-ish chain: conv0 copies input channel 0 to feature 0; body passes feature 0 through
        w.conv_w[0][0, 0, 1, 1] = 1.0
        for k in range(1, 17):
            w.conv_w[k][0, 0, 1, 1] = 1.0
        c, i, j = 1, s - 1, 0
        w.conv_w[17][c * s * s + i * s + j, 0, 1, 1] = 1.0
        x = np.zeros((3, 5, 6), np.float32)
        x[0, 2, 3] = 0.5
        y = srvgg.forward(x, w)
        learned = y - np.repeat(np.repeat(x, s, 1), s, 2)
        nz = np.argwhere(np.abs(learned) > 1e-6)
        assert nz.tolist() == [[c, 2 * s + i, 3 * s + j]]
        assert abs(learned[c, 2 * s + i, 3 * s + j] - 0.5) < 1e-6


def test_quantise_rule():
    v = np.array([-0.2, 0.0, 0.5 / 255, 0.49 / 255, 1.0, 1.3, 254.5 / 255], np.float32)
    assert srvgg.quantise(v).tolist() == [0, 0, 1, 0, 255, 255, 255]


def test_ncnn_python_round_trip(tmp_path):
    w = srvgg.make_weights(4, 77)
    for fp16 in (True, False):
        p, b = str(tmp_path / "m.param"), str(tmp_path / "m.bin")
        srvgg.write_ncnn(w, p, b, fp16=fp16)
        r = srvgg.read_ncnn(p, b)
        assert r.scale == 4
        assert all(np.array_equal(x, y) for x, y in zip(w.conv_w, r.conv_w))
        assert all(np.array_equal(x, y) for x, y in zip(w.conv_b, r.conv_b))
        assert all(np.array_equal(x, y) for x, y in zip(w.slopes, r.slopes))


def test_rejects_bad_arguments():
    w = srvgg.make_weights(2, 1)
    with pytest.raises(ValueError):
        srvgg.upscale(np.zeros((8, 8, 3), np.uint8), w, prepad=10)   # reflect-101 undefined
    with pytest.raises(ValueError):
        srvgg.upscale(np.zeros((8, 8), np.uint8), w)
    with pytest.raises(ValueError):
        srvgg.make_weights(5, 1)


# ---------------------------------------------------------------------------------------------
# colour conversion oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("matrix", [601, 709])
def test_yuv_oracle_known_answers_and_float_form(matrix):
    from oracle import colour
    # known answers of the limited-range 10-bit encoding: black, white, mid grey carry no chroma
    for val, y10 in ((0, 64), (255, 940), (128, 504)):
        f = np.full((4, 4, 3), val, np.uint8)
        y, u, v = colour.rgb_to_yuv420p10(f, matrix)
        assert (y == y10).all() and (u == 512).all() and (v == 512).all()
    # primaries: red drives Cr to its maximum (960), blue drives Cb to its maximum
    red = np.zeros((2, 2, 3), np.uint8); red[..., 0] = 255
    blue = np.zeros((2, 2, 3), np.uint8); blue[..., 2] = 255
    kr, kb = colour.KR_KB[matrix]
    y, u, v = colour.rgb_to_yuv420p10(red, matrix)
    assert v[0, 0] == 960 and y[0, 0] == round(64 + 876 * kr)
    y, u, v = colour.rgb_to_yuv420p10(blue, matrix)
    assert u[0, 0] == 960 and y[0, 0] == round(64 + 876 * kb)
    # the integer definition stays within one code value of the fp64 textbook equations (odd sizes too)
    rng = np.random.default_rng(5)
    f = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    yi, ui, vi = colour.rgb_to_yuv420p10(f, matrix)
    yf, uf, vf = colour.rgb_to_yuv420p10_float(f, matrix)
    assert yi.shape == (37, 53) and ui.shape == (19, 27) and vi.shape == (19, 27)
    assert np.abs(yi - yf).max() <= 0.51 and np.abs(ui - uf).max() <= 0.51 and np.abs(vi - vf).max() <= 0.51
    assert yi.min() >= 64 and yi.max() <= 940 and ui.min() >= 64 and ui.max() <= 960
