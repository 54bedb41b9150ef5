"""Writes a reference-shaped temp directory for the segment scheduler: <dir>/video.temp (the reference's Video JSON) and
<dir>/tmp_frames/{i}/frame%08d.png with synthetic frames.  usage: make_segments.py DIR SEGMENTS FRAMES_PER_SEGMENT W H SCALE"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import reve_b200  # noqa: E402

d, nseg, per, w, h, s = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
st = reve_b200.VideoState.new("in.mkv", "out.mkv", nseg * per, 24.0, per, s)
st.segments = [(i, per) for i in range(nseg)]
os.makedirs(d, exist_ok=True)
open(os.path.join(d, "video.temp"), "w").write(st.to_json())
rng = np.random.default_rng(0)
base = np.zeros((h, w, 3), np.uint8)
for _ in range(40):
    x0, y0 = int(rng.integers(0, w - 50)), int(rng.integers(0, h - 50))
    base[y0:y0 + int(rng.integers(20, h // 2)), x0:x0 + int(rng.integers(20, w // 2))] = rng.integers(0, 256, 3)
for i in range(nseg):
    sd = os.path.join(d, "tmp_frames", str(i))
    os.makedirs(sd, exist_ok=True)
    for k in range(per):
        cv2.imwrite(os.path.join(sd, f"frame{k + 1:08d}.png"), np.roll(base, (7 * (i * per + k), 11 * k), axis=(0, 1)), [cv2.IMWRITE_PNG_COMPRESSION, 1])
print("wrote", nseg, "segments of", per, "frames", w, "x", h)
