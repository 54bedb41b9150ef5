"""CPU tests of the host side of libreve_cuda: exported ABI, model loader, geometry, error paths.
No compute call is made here (there is no GPU on the CPU box and no CPU fallback in the library)."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

import reve_b200
from oracle import srvgg
from reve_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    hdr = open(os.path.join(ROOT, "include", "reve_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(reve_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_functions()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/reve_cuda.h but not exported"
    # and the ctypes binding covers exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_strerror(lib):
    assert lib.reve_version() == 200
    assert lib.reve_strerror(0) == b"ok"
    assert b"sm_100" in lib.reve_strerror(-6)
    assert lib.reve_strerror(-999) == b"unknown status"


def test_random_model_is_bit_identical_to_oracle(lib, tmp_path):
    for scale, seed in ((2, 7), (3, 123456789), (4, 2 ** 63 + 5)):
        m = reve_b200.Model.random(scale, seed)
        assert m.scale == scale
        p, b = str(tmp_path / f"m{scale}.param"), str(tmp_path / f"m{scale}.bin")
        m.save_ncnn(p, b, fp16=True)
        got = srvgg.read_ncnn(p, b)
        want = srvgg.make_weights(scale, seed)
        assert all(np.array_equal(x, y) for x, y in zip(got.conv_w, want.conv_w))
        assert all(np.array_equal(x, y) for x, y in zip(got.conv_b, want.conv_b))
        assert all(np.array_equal(x, y) for x, y in zip(got.slopes, want.slopes))


@pytest.mark.parametrize("fp16", [True, False])
def test_loader_round_trip_against_python_writer(lib, tmp_path, fp16):
    w = srvgg.make_weights(3, 99, fp16_weights=fp16)
    p, b = str(tmp_path / "a.param"), str(tmp_path / "a.bin")
    srvgg.write_ncnn(w, p, b, fp16=fp16)
    m = reve_b200.Model.load_ncnn(p, b)
    assert m.scale == 3
    p2, b2 = str(tmp_path / "b.param"), str(tmp_path / "b.bin")
    m.save_ncnn(p2, b2, fp16=fp16)
    assert open(b, "rb").read() == open(b2, "rb").read()
    r = srvgg.read_ncnn(p2, b2)       # the C++ writer's .param parses with the independent reader
    assert all(np.array_equal(x, y) for x, y in zip(w.conv_w, r.conv_w))


def _write(tmp_path, w=None, scale=2):
    w = w or srvgg.make_weights(scale, 1)
    p, b = str(tmp_path / "m.param"), str(tmp_path / "m.bin")
    srvgg.write_ncnn(w, p, b)
    return p, b


def _load_err(p, b):
    with pytest.raises(reve_b200.ReveError) as e:
        reve_b200.Model.load_ncnn(p, b)
    return e.value


def test_loader_rejects_malformed_files(lib, tmp_path):
    p, b = _write(tmp_path)
    assert _load_err(str(tmp_path / "missing.param"), b).status == -4
    assert _load_err(p, str(tmp_path / "missing.bin")).status == -4
    txt = open(p).read()
    bad = str(tmp_path / "bad.param")
    open(bad, "w").write(txt.replace("7767517", "7767518"))
    assert _load_err(bad, b).status == -5                       # magic
    open(bad, "w").write(txt.replace("PReLU            PRelu_1 ", "ReLU             Relu_1 ", 1))
    assert _load_err(bad, b).status == -5                       # foreign layer type
    open(bad, "w").write(txt.replace(" 0=64 1=3 11=3", " 0=64 1=5 11=5", 1))
    assert _load_err(bad, b).status == -5                       # not 3x3
    open(bad, "w").write(txt.replace("PixelShuffle     DepthToSpace_36          1 1 conv17 ps 0=2", "PixelShuffle     DepthToSpace_36          1 1 conv17 ps 0=3"))
    assert _load_err(bad, b).status == -5                       # shuffle factor != last conv / Interp
    data = open(b, "rb").read()
    tb = str(tmp_path / "t.bin")
    open(tb, "wb").write(data[:-8])
    assert _load_err(p, tb).status == -5                        # truncated
    open(tb, "wb").write(data + b"\0\0\0\0")
    assert _load_err(p, tb).status == -5                        # trailing bytes
    open(tb, "wb").write(struct.pack("<I", 0x000D4B38) + data[4:])
    assert _load_err(p, tb).status == -5                        # int8 tag unsupported
    # drop one conv+prelu pair: wrong depth
    lines = txt.split("\n")
    lines = [l for l in lines if "Conv_10 " not in l and "PRelu_11 " not in l]
    open(bad, "w").write("\n".join(lines).replace("\n40 44\n", "\n38 42\n").replace("conv5 prelu5", "conv5 prelu5"))
    assert _load_err(bad, b).status == -5


def test_model_random_rejects_bad_scale(lib):
    for s in (0, 1, 5):
        with pytest.raises(reve_b200.ReveError) as e:
            reve_b200.Model.random(s, 1)
        assert e.value.status == -1


def _py_axis(n, tile, prepad):
    t = tile if tile > 0 else n
    src, out = [], []
    for t0 in range(0, n, t):
        tn = min(t, n - t0)
        if t0 > 0:
            src.append(-1); out.append(-1)
        for i in range(-prepad, tn + prepad):
            src.append(int(srvgg.reflect101(np.array([t0 + i]), n)[0]))
            out.append(t0 + i if 0 <= i < tn else -1)
    return src, out


@pytest.mark.parametrize("w,h,tile,prepad", [(1920, 1080, 200, 10), (1920, 1080, 0, 10), (450, 230, 200, 10),
                                             (11, 11, 200, 10), (5, 3, 2, 0), (1280, 720, 200, 10), (960, 540, 100, 10)])
def test_geometry_tables(lib, w, h, tile, prepad):
    cw, ch, sx, ox, sy, oy = reve_b200.geometry(w, h, 2, tile, prepad)
    psx, pox = _py_axis(w, tile, prepad)
    psy, poy = _py_axis(h, tile, prepad)
    assert (cw, ch) == (len(psx), len(psy))
    assert sx.tolist() == psx and ox.tolist() == pox and sy.tolist() == psy and oy.tolist() == poy
    # every output coordinate is produced exactly once
    assert sorted(v for v in pox if v >= 0) == list(range(w))
    assert sorted(v for v in poy if v >= 0) == list(range(h))
    if tile == 200 and (w, h) == (1920, 1080):
        assert (cw, ch) == (2129, 1205)        # 10 x 6 padded tiles + 9 x 5 gaps


def test_geometry_rejects_bad_arguments(lib):
    for args in ((0, 10, 2, 0, 0), (10, 10, 5, 0, 0), (10, 10, 2, -1, 0), (10, 10, 2, 0, 10), (10, 10, 2, 0, -1),
                 (20000, 10, 2, 0, 0)):
        with pytest.raises(reve_b200.ReveError) as e:
            reve_b200.geometry(*args)
        assert e.value.status == -1


def test_null_and_invalid_arguments_are_reported_not_crashed(lib):
    assert lib.reve_model_load_ncnn(None, None, None) == -1
    h = C.c_void_p()
    assert lib.reve_ctx_create(0, None, 64, 64, 0, 10, 2, C.byref(h)) == -1
    assert lib.reve_submit(None, None, 0, None, 0, 0) == -1
    assert lib.reve_wait(None, None) == -1
    assert lib.reve_device_count(None) == -1
    lib.reve_model_free(None)
    lib.reve_ctx_destroy(None)
    lib.reve_host_free(None)
    m = reve_b200.Model.random(2, 1)
    assert lib.reve_ctx_create(0, m._h, 64, 64, 0, 10, 0, C.byref(h)) == -1      # ring depth
    assert lib.reve_ctx_create(0, m._h, 64, 64, 0, 10, 17, C.byref(h)) == -1
    assert b"ring_depth" in lib.reve_last_error(None)


def test_no_cpu_fallback_without_device(lib):
    """On a box without an sm_100 device context creation must fail loudly (CUDA / ARCH error)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    m = reve_b200.Model.random(2, 1)
    with pytest.raises(reve_b200.ReveError) as e:
        reve_b200.Upscaler(m, 64, 64)
    assert e.value.status in (-3, -6)
    n = C.c_int(-1)
    rc = lib.reve_device_count(C.byref(n))
    assert rc in (0, -3) and n.value <= 0


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under reve_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "reve_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "srvgg_ref" not in txt, f


def _srvgg_state_dict(w):
    """State dict of Real-ESRGAN's SRVGGNetCompact(num_conv=16) holding the oracle weights `w`: nn.Sequential
    `body` of conv, PReLU, conv, PReLU, ..., conv (indices 0..34)."""
    import torch
    sd = {}
    for k in range(18):
        sd[f"body.{2 * k}.weight"] = torch.from_numpy(w.conv_w[k].copy())
        sd[f"body.{2 * k}.bias"] = torch.from_numpy(w.conv_b[k].copy())
        if k < 17:
            sd[f"body.{2 * k + 1}.weight"] = torch.from_numpy(w.slopes[k].copy())
    return sd


@pytest.mark.parametrize("scale", [2, 3, 4])
def test_pth_ingestion_matches_ncnn_route(lib, tmp_path, scale):
    """a .pth checkpoint (params / params_ema wrapper) gives the same model as
    the .param/.bin pair of the same weights."""
    import torch
    w = srvgg.make_weights(scale, 11)
    pth = str(tmp_path / "realesr-animevideov3.pth")
    torch.save({"params_ema": _srvgg_state_dict(w)}, pth)
    m = reve_b200.Model.from_pth(pth)
    assert m.scale == scale
    p, b = str(tmp_path / "m.param"), str(tmp_path / "m.bin")
    m.save_ncnn(p, b, fp16=False)
    got = srvgg.read_ncnn(p, b)
    assert all(np.array_equal(x, y) for x, y in zip(got.conv_w, w.conv_w))
    assert all(np.array_equal(x, y) for x, y in zip(got.conv_b, w.conv_b))
    assert all(np.array_equal(x, y) for x, y in zip(got.slopes, w.slopes))
    # for_scale picks the checkpoint up from a models directory
    os.makedirs(tmp_path / "models")
    torch.save({"params": _srvgg_state_dict(w)}, str(tmp_path / "models" / f"realesr-animevideov3-x{scale}.pth"))
    m2 = reve_b200.Model.for_scale(scale, str(tmp_path / "models"))
    m2.save_ncnn(p, b, fp16=False)
    assert all(np.array_equal(x, y) for x, y in zip(srvgg.read_ncnn(p, b).conv_w, w.conv_w))


def test_pth_ingestion_rejects_foreign_checkpoints(lib):
    w = srvgg.make_weights(2, 3)
    sd = _srvgg_state_dict(w)
    with pytest.raises(ValueError):
        reve_b200.Model.from_state_dict({k: v for k, v in sd.items() if not k.startswith("body.34")})
    with pytest.raises(ValueError):
        reve_b200.Model.from_state_dict(sd, scale=3)              # a x2 checkpoint asked for as x3
    bad = dict(sd)
    bad["body.4.weight"] = sd["body.4.weight"][:, :32]
    with pytest.raises(ValueError):
        reve_b200.Model.from_state_dict(bad)
    nan = dict(sd)
    t = sd["body.6.weight"].clone()
    t[0, 0, 0, 0] = float("nan")
    nan["body.6.weight"] = t
    with pytest.raises(reve_b200.ReveError) as e:
        reve_b200.Model.from_state_dict(nan)
    assert e.value.status == -5


def test_launch_plan_for_the_baseline_geometries():
    """The body layers run as chains of 4, 2 or 1 layers per launch, chosen from the canvas width (include/reve_cuda.h,
    reve_launch_plan): BASELINE.json's four frame sizes, tile 200 / pre-pad 10, and the whole-frame variant."""
    plan = reve_b200.launch_plan
    # 1080p: canvas 2129 columns -> 18 strips of 120 for chains of 4 (17 of 126 for single layers: the chain still wins)
    assert plan(1920, 1080, 2) == {"layers_per_launch": 4, "strip_px": 120, "n_strips": 18, "launches_per_batch": 6}
    assert plan(1280, 720, 4)["layers_per_launch"] == 4 and plan(1280, 720, 4)["n_strips"] == 12
    assert plan(960, 540, 3) == {"layers_per_launch": 4, "strip_px": 120, "n_strips": 9, "launches_per_batch": 6}
    # 480p: canvas 723 columns -> a chain of 4 would need 7 strips instead of 6: chains of 2 (strips of 124)
    assert plan(640, 480, 2) == {"layers_per_launch": 2, "strip_px": 124, "n_strips": 6, "launches_per_batch": 10}
    # whole-frame 1080p: canvas 1940 -> 16 strips of 124, 17 of 120
    assert plan(1920, 1080, 2, tile=0)["layers_per_launch"] == 2
    # a canvas narrower than one strip: nothing to lose, chains of 4
    p1 = plan(40, 30, 2, tile=0)
    assert p1["n_strips"] == 1 and p1["layers_per_launch"] == 4
    with pytest.raises(reve_b200.ReveError):
        plan(0, 10, 2)


def test_missing_model_is_a_hard_error(lib, tmp_path):
    """a misplaced models directory must not produce a segment of noise frames.  The reference's spawned
    upscaler fails when models/realesr-animevideov3-x{s}.param|.bin are absent; random weights are an explicit opt-in."""
    with pytest.raises(reve_b200.ReveError) as e:
        reve_b200.Model.for_scale(2, str(tmp_path / "nowhere"))
    assert e.value.status == -4 and "realesr-animevideov3-x2" in str(e.value)
    m = reve_b200.Model.for_scale(2, str(tmp_path / "nowhere"), allow_random=True, seed=5)
    assert m.scale == 2
    # the segment-level mirror fails the same way, before it touches the GPU
    np.save(str(tmp_path / "frame00000001.npy"), np.zeros((8, 8, 3), np.uint8))
    with pytest.raises(reve_b200.ReveError) as e:
        reve_b200.upscale_segment(str(tmp_path), str(tmp_path / "o"), 3, model_dir=str(tmp_path / "nowhere"))
    assert e.value.status == -4


def test_fp32_payload_outside_the_fp16_range_is_rejected(lib, tmp_path):
    """an fp32-tagged .bin with |w| > 65504 would pack to +-inf on the device; same rule as
    reve_model_from_arrays."""
    w = srvgg.make_weights(2, 4, fp16_weights=False)
    w.conv_w[5][3, 2, 1, 0] = 70000.0
    p, b = str(tmp_path / "big.param"), str(tmp_path / "big.bin")
    srvgg.write_ncnn(w, p, b, fp16=False)
    with pytest.raises(reve_b200.ReveError) as e:
        reve_b200.Model.load_ncnn(p, b)
    assert e.value.status == -5 and "fp16" in str(e.value)
    w.conv_w[5][3, 2, 1, 0] = 1.0
    w.conv_b[7][0] = float("inf")
    srvgg.write_ncnn(w, p, b, fp16=False)
    with pytest.raises(reve_b200.ReveError) as e:
        reve_b200.Model.load_ncnn(p, b)
    assert e.value.status == -5


def test_ctx_options_are_validated_before_any_device_work(lib):
    """reve_ctx_create_ex: the options struct replaces every environment knob of round 1 (the library reads no
    environment variables); bad values are REVE_E_INVAL, not silently ignored."""
    m = reve_b200.Model.random(2, 1)
    opt = _lib.reve_ctx_options()
    h = C.c_void_p()
    assert lib.reve_ctx_create_ex(0, m._h, 64, 64, 200, 10, 2, C.byref(opt), C.byref(h)) == -1    # struct_size unset
    opt.struct_size = C.sizeof(_lib.reve_ctx_options)
    opt.layers_per_launch = 3
    assert lib.reve_ctx_create_ex(0, m._h, 64, 64, 200, 10, 2, C.byref(opt), C.byref(h)) == -1
    assert b"layers_per_launch" in lib.reve_last_error(None)
    opt.layers_per_launch, opt.max_batch = 0, 9
    assert lib.reve_ctx_create_ex(0, m._h, 64, 64, 200, 10, 2, C.byref(opt), C.byref(h)) == -1
    src = open(os.path.join(ROOT, "reve_b200", "csrc", "api.cu")).read()
    for f in os.listdir(os.path.join(ROOT, "reve_b200", "csrc")):
        if f.endswith((".cu", ".cpp", ".cuh", ".h")):
            assert "getenv" not in open(os.path.join(ROOT, "reve_b200", "csrc", f)).read(), f
    assert "cudaLaunchAttributeCooperative" in open(os.path.join(ROOT, "reve_b200", "csrc", "conv_umma.cu")).read()
    assert src.count("launch_conv_chain(") == 1
