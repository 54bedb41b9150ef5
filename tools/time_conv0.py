"""Device time of the first conv ALONE (GPU otherwise idle, so every build runs at the same, maximum, clock): the
test hook reve_debug_features(layer = 1) launches only that kernel; the context's profiling brackets it with CUDA events.
usage: REVE_LIB=... python tools/time_conv0.py [WxHxS] [flags]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import reve_b200

w, h, s = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1920x1080x2").split("x"))
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
frame = np.random.default_rng(0).integers(0, 256, (h, w, 3), dtype=np.uint8)
with reve_b200.Upscaler(reve_b200.Model.random(s, 1), w, h, tile=200, prepad=10, debug_flags=flags) as up:
    for _ in range(3):
        up.debug_features(frame, 1)
    up.set_profiling(True)
    up.profile(reset=True)
    ts = []
    for _ in range(10):
        up.debug_features(frame, 1)
        p = up.profile(reset=True)
        ts.append(p["ms_conv0"] / max(p["launches_conv0"], 1))
    print(json.dumps({"lib": os.path.basename(os.environ.get("REVE_LIB", "default")), "flags": flags, "size": [w, h, s],
                      "conv0_ms_min": min(ts), "conv0_ms_median": sorted(ts)[len(ts) // 2]}), flush=True)
