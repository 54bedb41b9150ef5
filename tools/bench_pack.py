"""Stand-alone u8 <-> fp16 unpack / pack kernels (reve_b200/csrc/pack.cu) against the measured HBM bandwidth.
Algorithmic bytes: unpack = 3 B/px in + 8 B per canvas px out; pack = 8 B per canvas px * s^2 in + 3 B per output px out.
GPU box only:  python tools/bench_pack.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import reve_b200  # noqa: E402

peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
for (w, h, s) in ((1920, 1080, 2), (1280, 720, 4), (960, 540, 3)):
    with reve_b200.Upscaler(reve_b200.Model.random(s, 1), w, h, tile=200, prepad=10) as up:
        cw, ch, *_ = reve_b200.geometry(w, h, s, 200, 10)
        frame = np.random.default_rng(0).integers(0, 256, (h, w, 3), dtype=np.uint8)
        _, ms_u = up.debug_unpack(frame, reps=50)
        y = np.random.default_rng(1).random((ch * s, cw * s, 3), dtype=np.float32)
        _, ms_p = up.debug_pack(y, reps=50)
    bu = w * h * 3 + cw * ch * 8
    bp = cw * ch * s * s * 8 + w * h * s * s * 3
    print(json.dumps({"frame": [w, h], "scale": s, "unpack_ms": round(ms_u, 4), "unpack_gbs": round(bu / ms_u / 1e6, 1),
                      "unpack_frac_hbm": round(bu / ms_u / 1e6 / peaks["hbm_gbs"], 3), "pack_ms": round(ms_p, 4),
                      "pack_gbs": round(bp / ms_p / 1e6, 1), "pack_frac_hbm": round(bp / ms_p / 1e6 / peaks["hbm_gbs"], 3)}), flush=True)
