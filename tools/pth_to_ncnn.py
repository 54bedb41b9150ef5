"""Convert a Real-ESRGAN SRVGGNetCompact checkpoint (.pth) to the ncnn .param/.bin pair the reference's
models directory holds (`models/realesr-animevideov3-x{s}.param|.bin`, README.md:27-29 of the reference).

    python tools/pth_to_ncnn.py realesr-animevideov3.pth models/realesr-animevideov3-x4 [--fp32]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import reve_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("pth")
ap.add_argument("out_base", help="output path without extension")
ap.add_argument("--fp32", action="store_true", help="raw fp32 payload instead of the fp16-tagged one")
a = ap.parse_args()
m = reve_b200.Model.from_pth(a.pth)
m.save_ncnn(a.out_base + ".param", a.out_base + ".bin", fp16=not a.fp32)
print(f"x{m.scale} model written to {a.out_base}.param / .bin")
