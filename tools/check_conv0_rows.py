"""The two first-conv kernels side by side (GPU box): layer-1 features and final frames of the row-streaming kernel
(conv0_rows.cu) against the im2col kernel (conv0.cu) and against the oracle, then their device times at 1080p x2."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import reve_b200
from helpers import feature_report, oracle_canvas
from oracle import srvgg

IM2COL = reve_b200.DBG_CONV0_IM2COL


def compare(w, h, scale, tile, grid=0, batch=0):
    wts = srvgg.make_weights(scale, 1234)
    model = reve_b200.Model.random(scale, 1234)
    frame = srvgg.synthetic_frame(w, h, 5, "random")
    res = {}
    for name, flags in (("im2col", IM2COL), ("rows", 0)):
        with reve_b200.Upscaler(model, w, h, tile=tile, prepad=10, ring_depth=2, debug_flags=flags, debug_grid=grid, max_batch=batch) as up:
            res[name] = (up.debug_features(frame, 1), up.upscale(frame))
    ref = oracle_canvas(frame, wts, tile, 10, 1)
    rep = feature_report(res["rows"][0], ref, np.ones(ref.shape[0], bool))
    d = np.abs(res["rows"][0] - res["im2col"][0])
    out = {"case": [w, h, scale, tile, grid], "oracle_bad_frac": rep["bad_frac"], "max_abs_vs_im2col": float(d.max()),
           "features_identical_frac": float((d == 0).mean()), "frames_identical": bool(np.array_equal(res["rows"][1], res["im2col"][1])),
           "frames_max_diff": int(np.abs(res["rows"][1].astype(int) - res["im2col"][1].astype(int)).max())}
    print(json.dumps(out), flush=True)
    return out


def timing(w, h, scale, n=24):
    model = reve_b200.Model.random(scale, 1234)
    frames = [srvgg.synthetic_frame(w, h, i, "random") for i in range(4)]
    for name, flags in (("im2col", IM2COL), ("rows", 0), ("im2col", IM2COL), ("rows", 0)):
        with reve_b200.Upscaler(model, w, h, tile=200, prepad=10, debug_flags=flags, ring_depth=8) as up:
            hin = [up.pinned((h, w, 3)) for _ in range(4)]
            hout = [up.pinned((h * scale, w * scale, 3)) for _ in range(4)]
            for i in range(4):
                hin[i][...] = frames[i]
            for rep in range(3):
                if rep == 2:
                    up.set_profiling(True)
                    up.profile(reset=True)
                for k in range(n):
                    up.submit(hin[k % 4], hout[k % 4], k % 4)
                    if k >= 3:
                        up.wait()
                for _ in range(3):
                    up.wait()
            p = up.profile()
            print(json.dumps({"kernel": name, "size": [w, h, scale], "conv0_ms_per_frame": p["ms_conv0"] / max(p["frames"], 1),
                              "launches_conv0": p["launches_conv0"], "frames": p["frames"], "ms_conv0": p["ms_conv0"]}), flush=True)


if __name__ == "__main__":
    bad = 0
    for case in [(100, 30, 2, 0, 1), (100, 30, 2, 0, 0), (300, 40, 2, 0, 2), (300, 200, 2, 0, 0), (200, 150, 2, 64, 0), (137, 91, 3, 50, 0),
                 (150, 90, 4, 0, 3), (500, 300, 2, 200, 6), (640, 480, 2, 200, 0), (129, 17, 2, 0, 0), (16, 16, 2, 0, 0), (11, 11, 4, 0, 0)]:
        r = compare(*case)
        bad += (r["oracle_bad_frac"] != 0.0) or r["frames_max_diff"] > 1
    print("PARITY", "OK" if not bad else f"FAILED ({bad})", flush=True)
    timing(1920, 1080, 2)
    timing(1280, 720, 4)
    # the write-only ceiling: the kernel writes 128 B per canvas pixel and reads 3
    import torch
    buf = torch.empty(2129 * 1205 * 128 * 4, dtype=torch.uint8, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        buf.zero_()
    ev[0].record()
    for _ in range(10):
        buf.zero_()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 10
    print(json.dumps({"memset_bytes": buf.numel(), "ms": ms, "GBps": buf.numel() / ms / 1e6, "ms_per_1080p_canvas": ms / 4}), flush=True)
