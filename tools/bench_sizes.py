"""Throughput / latency at the other BASELINE.json configs (parity-test geometries, not bench lines):
720p x4, 540p x3 (latency too), 480p x2, 1080p x2 — device-resident frames, batch sizes 1 and 4."""
import json
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(w, h, scale, tile, batch, frames=64):
    import torch
    import reve_b200
    from _opts import opts_from_env
    os.environ["REVE_DEBUG_BATCH"] = str(batch)
    model = reve_b200.Model.random(scale, 1)
    up = reve_b200.Upscaler(model, w, h, tile=tile, prepad=10, ring_depth=8, **opts_from_env())
    n = 8
    d_in = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
    d_out = torch.empty((n, h * scale, w * scale, 3), dtype=torch.uint8, device="cuda")
    st = torch.cuda.ExternalStream(up.stream)
    for _ in range(3):
        up.upscale_device(d_in.data_ptr(), d_out.data_ptr(), n)
    up.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(frames // n):
        up.upscale_device(d_in.data_ptr(), d_out.data_ptr(), n)
    e1.record(st)
    up.sync()
    fps = frames / (e0.elapsed_time(e1) / 1e3)
    # single-frame latency through submit/wait (H2D + kernels + D2H), ring depth 1 semantics
    hin, hout = up.pinned((h, w, 3)), up.pinned((h * scale, w * scale, 3))
    lat = []
    for _ in range(10):
        t0 = time.perf_counter()
        up.submit(hin, hout, 0)
        up.wait()
        lat.append((time.perf_counter() - t0) * 1e3)
    up.close()
    return fps, float(np.median(lat))


if __name__ == "__main__":
    cfgs = [(1280, 720, 4, 200), (960, 540, 3, 200), (640, 480, 2, 200), (1920, 1080, 2, 200), (1920, 1080, 2, 0)]
    chains = sys.argv[1].split(",") if len(sys.argv) > 1 else [os.environ.get("REVE_CHAIN", "")]
    for (w, h, s, t) in cfgs:
      for chain in chains:
        if chain != "":
            os.environ["REVE_CHAIN"] = chain
        for b in (1, 4):
            fps, lat = run(w, h, s, t, b, frames=256)
            print(json.dumps({"frame": [w, h], "scale": s, "tile": t, "chain": chain, "batch": b, "fps": round(fps, 1),
                              "out_mpix_s": round(fps * w * h * s * s / 1e6, 1), "latency_ms_1frame": round(lat, 3)}), flush=True)
