"""How far the device's fp16 features are from the fp32 oracle, per layer (test infrastructure: imports oracle/).
Used to set the tolerance of tests/helpers.py:feature_report from data.  GPU box only."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import needed_rows, oracle_canvas  # noqa: E402
from oracle import srvgg  # noqa: E402

import reve_b200  # noqa: E402

worst = 0.0
for (w, h, scale, tile, seed, opts) in [(300, 200, 2, 0, 1234, {}), (200, 150, 2, 64, 1234, {}), (137, 91, 3, 50, 7, {"layers_per_launch": 2}),
                                        (150, 90, 4, 0, 99, {"layers_per_launch": 1}), (500, 300, 2, 200, 5, {})]:
    wts = srvgg.make_weights(scale, seed)
    model = reve_b200.Model.random(scale, seed)
    for kind in ("random", "edges"):
        frame = srvgg.synthetic_frame(w, h, 5, kind)
        with reve_b200.Upscaler(model, w, h, tile=tile, prepad=10, ring_depth=2, **opts) as up:
            for layer in (1, 2, 5, 9, 13, 17):
                dev = up.debug_features(frame, layer)
                ref = oracle_canvas(frame, wts, tile, 10, layer)
                rows = needed_rows(h, scale, tile, 10, layer)
                d, r = dev[rows], ref[rows]
                err = np.abs(d - r)
                # smallest a such that err <= a * (1 + |ref|) everywhere
                a = float((err / (1.0 + np.abs(r))).max())
                worst = max(worst, a)
                print(json.dumps({"case": [w, h, scale, tile, kind], "layer": layer, "max_abs_err": float(err.max()),
                                  "ref_absmax": float(np.abs(r).max()), "min_tol_a": a}), flush=True)
print(json.dumps({"worst_min_tol_a": worst, "meaning": "err <= a + a*|ref| holds everywhere with a = this"}))
