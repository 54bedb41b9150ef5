"""Dump the per-step timeline of CTA 0 of body layer 5 (REVE_DEBUG_TRACE).  GPU box only.

trace[0..999]   clock at the start of MMA step i (interleaved order of the CTA's two streams)
trace[1000]     look-ahead misses
trace[1024+4e]  epilogue group 0, event e: wait start, accumulator full, slot handed back, row stored
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("REVE_DEBUG_TRACE", "1")   # "tail": trace the tail kernel instead of body layer 5
from _opts import opts_from_env  # noqa: E402
import reve_b200  # noqa: E402
from reve_b200 import _lib  # noqa: E402

tile = int(sys.argv[1]) if len(sys.argv) > 1 else 200
model = reve_b200.Model.random(2, 1)
up = reve_b200.Upscaler(model, 1920, 1080, tile=tile, prepad=10, ring_depth=2, **opts_from_env())
frame = np.random.default_rng(0).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
for _ in range(3):
    up.upscale(frame)
tr = np.zeros(2048, np.int64)
rc = _lib.load().reve_debug_trace(up._h, tr.ctypes.data, 2048)
assert rc == 0
steps = tr[:1000]
steps = steps[steps > 0]
d = np.diff(steps)
print(f"flags={os.environ.get('REVE_DEBUG_FLAGS', '0')} steps traced: {len(steps)}  look-ahead misses: {tr[1000]}")
if len(d):
    core = d[4:-4] if len(d) > 16 else d
    print(f"clk/step: mean {core.mean():.1f}  median {np.median(core):.1f}  p10 {np.percentile(core, 10):.0f}  p90 {np.percentile(core, 90):.0f}  max {core.max()}")
    print("first 40 step deltas:", d[:40].tolist())
epi = tr[1024:].reshape(256, 4)
t0 = steps[0] if len(steps) else 0
print("event  wait_start  got_full  released  stored   (cycles rel. to first MMA step)")
for e in range(0, 40):
    a, b, c, s = epi[e]
    if a == 0:
        continue
    print(f"{e:5d} {a - t0:10d} {b - t0:9d} {c - t0:9d} {(s - t0) if s else 0:8d}")
