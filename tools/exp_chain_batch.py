"""Chained body layers: frames/s and per-launch body time for 1, 2, 4 frames per launch (1080p x2, tile 200)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(env, frames=320, w=1920, h=1080, scale=2, tile=200):
    import torch
    import reve_b200
    from _opts import opts_from_env
    for k in ("REVE_DEBUG_BATCH", "REVE_DEBUG_FLAGS", "REVE_CHAIN"):
        os.environ.pop(k, None)
    os.environ.update(env)
    model = reve_b200.Model.random(scale, 1)
    up = reve_b200.Upscaler(model, w, h, tile=tile, prepad=10, ring_depth=8, **opts_from_env())
    n = 8
    d_in = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
    d_out = torch.empty((n, h * scale, w * scale, 3), dtype=torch.uint8, device="cuda")
    st = torch.cuda.ExternalStream(up.stream)
    for _ in range(6):
        up.upscale_device(d_in.data_ptr(), d_out.data_ptr(), n)
    up.sync()
    up.set_profiling(True)
    up.profile(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(frames // n):
        up.upscale_device(d_in.data_ptr(), d_out.data_ptr(), n)
    e1.record(st)
    up.sync()
    pr = up.profile(reset=True)
    fps = frames / (e0.elapsed_time(e1) / 1e3)
    up.close()
    return {"fps": round(fps, 1), "body_launch_ms": round(pr["ms_body"] / max(1, pr["timed_body"]), 4),
            "body_launches": pr["launches_body"], "ms_body_per_frame": round(pr["ms_body"] / frames, 4),
            "conv0_ms_per_frame": round(pr["ms_conv0"] / frames, 4), "tail_ms_per_frame": round(pr["ms_tail"] / frames, 4)}


if __name__ == "__main__":
    for chain in ("0", "4", "2"):
        for batch in ("1", "2", "4"):
            r = run({"REVE_CHAIN": chain, "REVE_DEBUG_BATCH": batch})
            r.update(chain=chain, batch=batch)
            print(json.dumps(r), flush=True)
