"""ctypes binding of oracle/srvgg_ref.c (the independent plain-C restatement).
Test infrastructure only; see the header of oracle/srvgg.py."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import srvgg

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libsrvgg_ref.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "srvgg_ref.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


def pack_params(w: srvgg.Weights) -> np.ndarray:
    parts = []
    for k in range(srvgg.NUM_CONV + 2):
        parts += [w.conv_w[k].reshape(-1), w.conv_b[k].reshape(-1)]
        if k <= srvgg.NUM_CONV:
            parts.append(w.slopes[k].reshape(-1))
    return np.ascontiguousarray(np.concatenate(parts).astype(np.float32))


def upscale(frame: np.ndarray, w: srvgg.Weights, tile: int = 200, prepad: int = 10) -> np.ndarray:
    lib = ctypes.CDLL(build())
    lib.srvgg_ref_upscale.restype = ctypes.c_int
    lib.srvgg_ref_upscale.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    frame = np.ascontiguousarray(frame, dtype=np.uint8)
    h, wpx = frame.shape[:2]
    out = np.empty((h * w.scale, wpx * w.scale, 3), np.uint8)
    params = pack_params(w)
    rc = lib.srvgg_ref_upscale(frame.ctypes.data, wpx, h, w.scale, tile, prepad,
                               params.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise ValueError("srvgg_ref_upscale rejected its arguments")
    return out
