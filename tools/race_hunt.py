"""Repeat the same 8 full-size frames through the device-resident path and compare every pass with the first one on the
device: a lost / early hand-over flag in the chained kernel shows up as a differing block of output rows.
usage: python tools/race_hunt.py [passes]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _opts import opts_from_env  # noqa: E402
import reve_b200  # noqa: E402
from oracle import srvgg  # noqa: E402


def main():
    import torch
    passes = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    w, h, s, n = 1920, 1080, 2, 8
    model = reve_b200.Model.random(s, 11)
    frames = np.stack([srvgg.synthetic_frame(w, h, 60 + i, "random" if i % 2 else "edges") for i in range(n)])
    bad = []
    with reve_b200.Upscaler(model, w, h, tile=200, prepad=10, ring_depth=8, **opts_from_env()) as up:
        d_in = torch.from_numpy(frames).cuda()
        ref = torch.zeros((n, h * s, w * s, 3), dtype=torch.uint8, device="cuda")
        out = torch.zeros_like(ref)
        up.upscale_device(d_in.data_ptr(), ref.data_ptr(), n)
        up.sync()
        for it in range(passes):
            out.fill_(0)
            torch.cuda.synchronize()
            up.upscale_device(d_in.data_ptr(), out.data_ptr(), n)
            up.sync()
            if not torch.equal(out, ref):
                d = (out != ref).any(dim=3)                     # [n, H, W]
                idx = d.nonzero()
                f = idx[:, 0].unique().tolist()
                bad.append({"pass": it, "pixels": int(d.sum()), "frames": f,
                            "rows": [int(idx[:, 1].min()), int(idx[:, 1].max())],
                            "cols": [int(idx[:, 2].min()), int(idx[:, 2].max())],
                            "max_abs_diff": int((out.int() - ref.int()).abs().max())})
    print(json.dumps({"passes": passes, "mismatching_passes": len(bad), "first": bad[:6]}))


if __name__ == "__main__":
    main()
