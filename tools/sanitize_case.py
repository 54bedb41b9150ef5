"""Small end-to-end case for compute-sanitizer (memcheck / synccheck / racecheck) on the GPU box:
    compute-sanitizer --tool synccheck python tools/sanitize_case.py
Round 1: all three tools report 0 errors (synccheck found a barrier-init race in conv0 that was fixed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _opts import opts_from_env  # noqa: E402
import reve_b200  # noqa: E402
from oracle import srvgg  # noqa: E402

for (w, h, s, t) in ((150, 60, 2, 64), (90, 50, 3, 0)):
    frame = srvgg.synthetic_frame(w, h, 1, "random")
    m = reve_b200.Model.random(s, 5)
    with reve_b200.Upscaler(m, w, h, tile=t, prepad=10, ring_depth=2, **opts_from_env()) as up:
        out = up.upscale(frame)
    ref = srvgg.upscale(frame, srvgg.make_weights(s, 5), tile=t, prepad=10)
    print(w, h, s, t, srvgg.parity(out, ref))
