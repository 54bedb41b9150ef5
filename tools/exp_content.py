"""How much the frame CONTENT moves the power-capped throughput (1080p x2, tile 200, device-resident, ~1.7 s per case):
uniform noise (bench.py's frames: every operand bit toggles), anime-like flat regions + hard edges, a constant frame."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(frames_np, frames=640):
    import torch
    import reve_b200
    from _opts import opts_from_env
    n, h, w, _ = frames_np.shape
    up = reve_b200.Upscaler(reve_b200.Model.random(2, 1), w, h, tile=200, prepad=10, ring_depth=8, **opts_from_env())
    d_in = torch.from_numpy(frames_np).cuda()
    d_out = torch.empty((n, h * 2, w * 2, 3), dtype=torch.uint8, device="cuda")
    st = torch.cuda.ExternalStream(up.stream)
    for _ in range(10):
        up.upscale_device(d_in.data_ptr(), d_out.data_ptr(), n)
    up.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(frames // n):
        up.upscale_device(d_in.data_ptr(), d_out.data_ptr(), n)
    e1.record(st)
    up.sync()
    fps = frames / (e0.elapsed_time(e1) / 1e3)
    up.close()
    return fps


if __name__ == "__main__":
    from oracle import srvgg
    w, h, n = 1920, 1080, 8
    cases = {
        "uniform_noise": np.stack([np.random.default_rng(i).integers(0, 256, (h, w, 3), dtype=np.uint8) for i in range(n)]),
        "flat_regions_and_edges": np.stack([srvgg.synthetic_frame(w, h, 10 + i, "edges") for i in range(n)]),
        "constant_grey": np.full((n, h, w, 3), 128, np.uint8),
    }
    for rep in range(2):
        for name, fr in cases.items():
            print(json.dumps({"content": name, "rep": rep, "fps": round(run(fr), 1)}), flush=True)
