"""Dump the per-row timeline of CTA 0 of body layer 5 (REVE_DEBUG_TRACE).  GPU box only."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["REVE_DEBUG_TRACE"] = "1"
import reve_b200  # noqa: E402
from reve_b200 import _lib  # noqa: E402

tile = int(sys.argv[1]) if len(sys.argv) > 1 else 200
model = reve_b200.Model.random(2, 1)
up = reve_b200.Upscaler(model, 1920, 1080, tile=tile, prepad=10, ring_depth=2)
frame = np.random.default_rng(0).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
for _ in range(3):
    up.upscale(frame)
tr = np.zeros(2048, np.int64)
rc = _lib.load().reve_debug_trace(up._h, tr.ctypes.data, 2048)
assert rc == 0
print("look-ahead misses / interior rows (3 frames, CTA 0 of layer 5):", tr[2040], "/", tr[2041])
mma = tr[:1024].reshape(256, 4)
epi = tr[1024:].reshape(256, 4)
t0 = mma[mma[:, 0] > 0][:, 0].min()
print("row  start  waited  issued | flags(s0,okF,okE) | epi: wait_start got_full arrived  (cycles rel. to first MMA row)")
prev = None
for i in range(0, 150):
    if mma[i, 0] == 0:
        continue
    a, b, c, f = mma[i]
    d = (a - prev) if prev is not None else 0
    prev = a
    t = i - 1
    e = epi[t] if 0 <= t < 256 else [0, 0, 0, 0]
    print(f"{i:3d} {a - t0:7d} {b - a:6d} {c - b:6d}  dRow={d:5d} | s0={f >> 4} okF={(f >> 1) & 1} okE={f & 1} | t={t:3d} {e[0] - t0:8d} {e[1] - t0:8d} {e[2] - t0:8d}")
