"""ORACLE (test infrastructure only; never imported by the product path): numpy restatement of the
RGB -> yuv420p10le conversion of reve_b200/csrc/yuv.cu.

The reference converts on the host: `ffmpeg -f image2 -i frame%08d.png -c:v libx265 -pix_fmt yuv420p10le`
(reve-cli/src/main.rs:306-326) lets swscale turn the RGB PNGs into limited-range 10-bit 4:2:0 with the
BT.601 matrix (swscale's default for untagged RGB).  swscale's exact bicubic chroma filter and dithering are
not pinned by the reference (no golden frames, ffmpeg absent here), so parity for this stage is defined
against the published colour equations: `rgb_to_yuv420p10` below is the integer definition the GPU follows
bit for bit, `rgb_to_yuv420p10_float` the fp64 textbook form it must stay within +-1 code value of.
"""
from __future__ import annotations

import numpy as np

KR_KB = {601: (0.299, 0.114), 709: (0.2126, 0.0722)}


def coeffs(matrix: int):
    """Integer coefficients scaled by 2^16 (luma: x 876/255; chroma: x 896/255), rounded half away from zero."""
    kr, kb = KR_KB[matrix]
    kg = 1.0 - kr - kb
    ys, cs = 876.0 / 255.0 * 65536.0, 896.0 / 255.0 * 65536.0

    def r(v):
        return int(v - 0.5) if v < 0 else int(v + 0.5)

    y = [r(kr * ys), r(kg * ys), r(kb * ys)]
    u = [r(-kr / (2 * (1 - kb)) * cs), r(-kg / (2 * (1 - kb)) * cs), r(0.5 * cs)]
    v = [r(0.5 * cs), r(-kg / (2 * (1 - kr)) * cs), r(-kb / (2 * (1 - kr)) * cs)]
    return y, u, v


def _block_sums(rgb: np.ndarray) -> np.ndarray:
    """Sum of each 2x2 block (edge pixels replicated for odd sizes): int64 [ceil(H/2), ceil(W/2), 3]."""
    h, w, _ = rgb.shape
    ys = np.minimum(np.arange(2 * ((h + 1) // 2)), h - 1)
    xs = np.minimum(np.arange(2 * ((w + 1) // 2)), w - 1)
    p = rgb[ys][:, xs].astype(np.int64)
    return p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2]


def rgb_to_yuv420p10(rgb: np.ndarray, matrix: int = 601):
    """u8 [H,W,3] -> (Y u16 [H,W], U u16 [ceil(H/2),ceil(W/2)], V): the integer definition of yuv.cu."""
    cy, cu, cv = coeffs(matrix)
    p = rgb.astype(np.int64)
    y = (cy[0] * p[..., 0] + cy[1] * p[..., 1] + cy[2] * p[..., 2] + (64 << 16) + (1 << 15)) >> 16
    s = _block_sums(rgb)
    u = (cu[0] * s[..., 0] + cu[1] * s[..., 1] + cu[2] * s[..., 2] + (512 << 18) + (1 << 17)) >> 18
    v = (cv[0] * s[..., 0] + cv[1] * s[..., 1] + cv[2] * s[..., 2] + (512 << 18) + (1 << 17)) >> 18
    return y.astype(np.uint16), u.astype(np.uint16), v.astype(np.uint16)


def rgb_to_yuv420p10_float(rgb: np.ndarray, matrix: int = 601):
    """fp64 textbook form (ITU-R BT.601 / BT.709 limited range, 10 bit, 2x2 box-filtered chroma), unrounded."""
    kr, kb = KR_KB[matrix]
    kg = 1.0 - kr - kb
    p = rgb.astype(np.float64) / 255.0
    yl = kr * p[..., 0] + kg * p[..., 1] + kb * p[..., 2]
    y = 4.0 * (16.0 + 219.0 * yl)
    m = _block_sums(rgb).astype(np.float64) / (4.0 * 255.0)
    ym = kr * m[..., 0] + kg * m[..., 1] + kb * m[..., 2]
    u = 4.0 * (128.0 + 224.0 * (m[..., 2] - ym) / (2 * (1 - kb)))
    v = 4.0 * (128.0 + 224.0 * (m[..., 0] - ym) / (2 * (1 - kr)))
    return y, u, v
