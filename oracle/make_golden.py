"""Generates the golden input/output vectors under tests/golden/ from the oracle.

The reference holds no golden vector for this path; these fixtures pin
the oracle itself (regressions, cross-language agreement with oracle/srvgg_ref.c) and give the
`-m gpu` tests committed expected outputs.  Real frames are decoded here from the reference's
demo assets (reve-cli/assets/*.mp4) with OpenCV and stored as small crops, because
/root/reference does not exist on the GPU box.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cref, srvgg  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF_ASSETS = "/root/reference/reve-cli/assets"


def real_crop(name: str, frame_idx: int, x0: int, y0: int, w: int, h: int):
    import cv2
    path = os.path.join(REF_ASSETS, name)
    cap = cv2.VideoCapture(path)
    cap.set(cv2.CAP_PROP_POS_FRAMES, frame_idx)
    ok, bgr = cap.read()
    cap.release()
    if not ok:
        raise RuntimeError(f"cannot decode {path}")
    rgb = np.ascontiguousarray(bgr[:, :, ::-1])
    return np.ascontiguousarray(rgb[y0:y0 + h, x0:x0 + w])


CASES = [
    # name, frame source, scale, seed, tile, prepad
    ("x2_whole_random", ("random", 48, 36, 11), 2, 101, 0, 10),
    ("x2_tile16_edges", ("edges", 50, 41, 12), 2, 102, 16, 10),
    ("x3_tile24_random", ("random", 45, 33, 13), 3, 103, 24, 10),
    ("x4_whole_edges", ("edges", 37, 29, 14), 4, 104, 0, 10),
    ("x2_tile200_real_onepiece", ("real", "onepiece_demo.mp4", 60, 200, 150, 224, 96), 2, 105, 200, 10),
    ("x4_tile20_real_test", ("real", "test.mp4", 300, 10, 10, 64, 48), 4, 106, 20, 10),
    ("x2_min_frame_pad10", ("random", 11, 11, 15), 2, 107, 200, 10),
    ("x3_pad0_tiny", ("random", 5, 3, 16), 3, 108, 2, 0),
    # BASELINE.json configs[0] geometry: one whole 640x480 frame of the reference's demo asset, upstream tile 200 / pad 10
    # (12 tiles: 3 full columns + a 40-px one, 2 full rows + an 80-px one).  Too large for the scalar C cross-check in
    # reasonable time: pinned by the torch oracle only (the C restatement agrees on the 224x96 crop of the same frame above).
    ("x2_tile200_real_onepiece_480p", ("real", "onepiece_demo.mp4", 60, 0, 0, 640, 480), 2, 109, 200, 10),
]
NO_C_CHECK = {"x2_tile200_real_onepiece_480p"}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, src, scale, seed, tile, prepad in CASES:
        if src[0] == "real":
            frame = real_crop(*src[1:])
        else:
            frame = srvgg.synthetic_frame(src[1], src[2], src[3], src[0])
        w = srvgg.make_weights(scale, seed)
        out = srvgg.upscale(frame, w, tile=tile, prepad=prepad)
        par = None
        if name not in NO_C_CHECK:
            out_c = cref.upscale(frame, w, tile=tile, prepad=prepad)
            par = srvgg.parity(out, out_c)
            assert par["within1"] == 1.0 and par["exact"] > 0.999, (name, par)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), frame=frame, out=out, scale=scale, seed=seed,
                            tile=tile, prepad=prepad)
        print(f"{name}: frame {frame.shape} -> {out.shape}, torch-vs-C {par}")


if __name__ == "__main__":
    main()
