"""First-contact GPU diagnostics: per-layer parity of the CUDA path against the oracle on small
frames, with error localisation.  Run on the GPU box:  python tools/gpu_diag.py > gpurun_out/diag.log
(test infrastructure: imports oracle/)."""
import json
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import feature_report, needed_rows, oracle_canvas  # noqa: E402
from oracle import srvgg  # noqa: E402

from _opts import opts_from_env  # noqa: E402
import reve_b200  # noqa: E402


def run_case(name, w_px, h_px, scale, tile, prepad, layers, env):
    for k in ("REVE_DEBUG_GRID",):
        os.environ.pop(k, None)
    os.environ.update(env)
    print(f"=== case {name}: {w_px}x{h_px} x{scale} tile={tile} prepad={prepad} env={env}", flush=True)
    wts = srvgg.make_weights(scale, 1234)
    frame = srvgg.synthetic_frame(w_px, h_px, 5, "random")
    model = reve_b200.Model.random(scale, 1234)
    ok = True
    with reve_b200.Upscaler(model, w_px, h_px, tile=tile, prepad=prepad, ring_depth=2, **opts_from_env()) as up:
        for layer in layers:
            t0 = time.time()
            dev = up.debug_features(frame, layer)
            ref = oracle_canvas(frame, wts, tile, prepad, layer)
            rep = feature_report(dev, ref, needed_rows(h_px, scale, tile, prepad, layer))
            good = rep["bad_frac"] == 0.0
            ok = ok and good
            print(f"  layer {layer:2d}: {'OK ' if good else 'BAD'} {json.dumps(rep)} ({time.time() - t0:.2f}s)", flush=True)
            if not good and layer >= 2:
                break
        out = up.upscale(frame)
        ref = srvgg.upscale(frame, wts, tile=tile, prepad=prepad)
        par = srvgg.parity(out, ref)
        good = par["within1"] >= 0.999 and par["psnr"] >= 50
        ok = ok and good
        print(f"  output : {'OK ' if good else 'BAD'} {json.dumps(par)}", flush=True)
        if not good:
            d = np.abs(out.astype(int) - ref.astype(int)).max(axis=2)
            rows = np.where((d > 1).any(axis=1))[0]
            cols = np.where((d > 1).any(axis=0))[0]
            print(f"    bad out rows {len(rows)} first {rows[:10].tolist()}  cols {len(cols)} first {cols[:10].tolist()}")
    return ok


def main():
    cases = [
        ("A1 single CTA", 100, 30, 2, 0, 10, [1, 2, 3, 17], {"REVE_DEBUG_GRID": "1"}),
        ("A2 two CTAs 3 strips", 300, 40, 2, 0, 10, [2, 17], {"REVE_DEBUG_GRID": "2"}),
        ("B one row per CTA", 100, 30, 2, 0, 10, [2, 17], {}),
        ("C 3 strips", 300, 200, 2, 0, 10, [2, 17], {}),
        ("D tiles", 200, 150, 2, 64, 10, [1, 2, 17], {}),
        ("E x3", 150, 90, 3, 0, 10, [17], {}),
        ("F x4 tiles", 150, 90, 4, 50, 10, [17], {}),
    ]
    results = {}
    for c in cases:
        try:
            results[c[0]] = run_case(*c)
        except Exception as e:  # keep going: a trapped kernel poisons the context, so stop then
            traceback.print_exc()
            results[c[0]] = f"EXC {e}"
            if "CUDA" in str(e) or "cuda" in str(e):
                print("CUDA failure: stopping (context is unusable after a kernel trap)")
                break
    print("SUMMARY", json.dumps(results))


if __name__ == "__main__":
    main()
