"""Per-CTA wall clock of body layer 5 (solo kernel, REVE_DEBUG_TRACE=1): which SMs finish late."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["REVE_DEBUG_TRACE"] = "1"
os.environ["REVE_CHAIN"] = "0"
from _opts import opts_from_env  # noqa: E402
import reve_b200  # noqa: E402
from reve_b200 import _lib  # noqa: E402

model = reve_b200.Model.random(2, 1)
up = reve_b200.Upscaler(model, 1920, 1080, tile=200, prepad=10, ring_depth=2, **opts_from_env())
frame = np.random.default_rng(0).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
for _ in range(3):
    up.upscale(frame)
tr = np.zeros(4096, np.int64)
assert _lib.load().reve_debug_trace(up._h, tr.ctypes.data, 4096) == 0
cta = tr[2048:2048 + 148 * 4].reshape(148, 4).copy()
smid = cta[:, 3] >> 32
steps = cta[:, 3] & 0xFFFFFFFF
t0 = cta[:, 1].min()
dur = (cta[:, 2] - cta[:, 1]) / np.maximum(steps, 1)
order = np.argsort(cta[:, 2])
print(f"end of last step: median {np.median(cta[:, 2] - t0) / 1e3:.1f} us, max {(cta[:, 2] - t0).max() / 1e3:.1f} us; ns/step median {np.median(dur):.0f}")
print("slowest 16 CTAs (block, smid, steps, end us, ns/step):")
for b in order[-16:]:
    print(f"  {b:3d} sm {smid[b]:3d} steps {steps[b]} end {(cta[b, 2] - t0) / 1e3:.1f} ns/step {dur[b]:.0f}")
print("fastest 4:", [(int(b), int(smid[b]), round(float(dur[b]))) for b in order[:4]])
