"""Timing experiments around L2 residency of the activation canvases (1080p x2, tile 200 / pad 10).
  default      4 frames per launch (1.3 GB canvases)
  batch1       1 frame per launch (328 MB canvases; the rows a layer wrote last are the ones the next reads first)
  alias        REVE_DEBUG_FLAGS=16: canvas rows alias each other (results garbage) -- nothing leaves L2
Device-resident frames, ~2 s sustained per case, A/B/A/B."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(env, frames=640, w=1920, h=1080, scale=2, tile=200):
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
    which = sys.argv[1] if len(sys.argv) > 1 else "l2"
    if which == "l2":
        cases = {"default": {}, "batch1": {"REVE_DEBUG_BATCH": "1"}, "alias": {"REVE_DEBUG_FLAGS": "16"},
                 "alias_batch1": {"REVE_DEBUG_FLAGS": "16", "REVE_DEBUG_BATCH": "1"}}
    else:
        cases = {"chain0": {"REVE_CHAIN": "0"}, "chain2": {"REVE_CHAIN": "2"}, "chain4": {"REVE_CHAIN": "4"}}
    for rep in range(2):
        for name, env in cases.items():
            print(json.dumps({"case": name, "rep": rep, "fps": round(run(env), 1)}), flush=True)
    if which == "l2":
        print(json.dumps({"case": "whole_frame_default", "fps": round(run({}, tile=0), 1)}), flush=True)
        print(json.dumps({"case": "whole_frame_alias", "fps": round(run({"REVE_DEBUG_FLAGS": "16"}, tile=0), 1)}), flush=True)
    else:
        for name, env in cases.items():
            print(json.dumps({"case": "whole_frame_" + name, "fps": round(run(env, tile=0), 1)}), flush=True)
            print(json.dumps({"case": "720p_x4_" + name, "fps": round(run(env, w=1280, h=720, scale=4), 1)}), flush=True)
