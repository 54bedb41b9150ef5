"""Small end-to-end case for compute-sanitizer (memcheck / synccheck / racecheck) on the GPU box:
    compute-sanitizer --tool synccheck python tools/sanitize_case.py
Round 1: all three tools report 0 errors (synccheck found a barrier-init race in conv0 that was fixed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _opts import opts_from_env  # noqa: E402
import reve_b200  # noqa: E402
from oracle import srvgg  # noqa: E402

import numpy as np  # noqa: E402

for (w, h, s, t) in ((150, 60, 2, 64), (90, 50, 3, 0), (140, 40, 4, 50)):
    frame = srvgg.synthetic_frame(w, h, 1, "random")
    m = reve_b200.Model.random(s, 5)
    with reve_b200.Upscaler(m, w, h, tile=t, prepad=10, ring_depth=2, **opts_from_env()) as up:
        out = up.upscale(frame)
        out2 = up.upscale(frame)                      # second pass: the self-balancing split reads the first pass's table
        cw, ch, *_ = reve_b200.geometry(w, h, s, t, 10)
        up.debug_unpack(frame)                        # stand-alone conversion kernels
        up.debug_pack(np.random.default_rng(0).random((ch * s, cw * s, 3), dtype=np.float32))
        up.set_output_format(reve_b200.FMT_YUV420P10LE_BT601)
        up.upscale_yuv(frame)                         # colour-conversion kernel
        info = up.launch_info()
    ref = srvgg.upscale(frame, srvgg.make_weights(s, 5), tile=t, prepad=10)
    print(w, h, s, t, info, bool(np.array_equal(out, out2)), srvgg.parity(out, ref))
