"""Shared helpers for the parity tests (oracle side).  Test infrastructure only."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import srvgg  # noqa: E402


def oracle_canvas(frame: np.ndarray, w: srvgg.Weights, tile: int, prepad: int, layer: int):
    """Oracle feature map after `layer` conv+PReLU stages laid out on the device canvas:
    float32 [CH, CW, 64], zeros at gap rows/columns.  Follows reve_b200/csrc/geometry.cpp's layout
    (padded tiles side by side, one gap pixel between tiles)."""
    h, wpx = frame.shape[:2]
    t = tile if tile > 0 else max(h, wpx)
    xs = list(range(0, wpx, t))
    ys = list(range(0, h, t))
    tws = [min(t, wpx - x0) for x0 in xs]
    ths = [min(t, h - y0) for y0 in ys]
    cw = sum(tw + 2 * prepad for tw in tws) + len(xs) - 1
    ch = sum(th + 2 * prepad for th in ths) + len(ys) - 1
    canvas = np.zeros((ch, cw, 64), np.float32)
    cy = 0
    for y0, th in zip(ys, ths):
        cx = 0
        for x0, tw in zip(xs, tws):
            tl = srvgg.padded_tile(frame, x0, y0, tw, th, prepad)
            x = (tl.astype(np.float32) * np.float32(1.0 / 255.0)).transpose(2, 0, 1)
            _, feats, _ = srvgg.forward(x, w, taps=True)
            f = feats[layer - 1]  # [64, ph, pw]
            canvas[cy:cy + th + 2 * prepad, cx:cx + tw + 2 * prepad] = f.transpose(1, 2, 0)
            cx += tw + 2 * prepad + 1
        cy += th + 2 * prepad + 1
    return canvas


def needed_rows(h: int, scale: int, tile: int, prepad: int, layer: int) -> np.ndarray:
    """Canvas rows the device computes after `layer` convolutions (bool [CH]).  Upstream keeps only the centre
    of every padded tile, so a row r pixels outside the kept region only matters to layers with at least r
    convolutions still to come (18 - layer); the device skips the others (reve_b200/csrc/api.cu:set_row_space)
    and leaves stale values there."""
    import reve_b200
    _, ch, _, _, src_y, out_y = reve_b200.geometry(max(h, prepad + 1), h, scale, tile, prepad)
    margin = 18 - layer
    if margin >= prepad:
        return np.ones(ch, bool)
    need = np.zeros(ch, bool)
    i = 0
    while i < ch:
        if src_y[i] < 0:
            i += 1
            continue
        j = i
        while j < ch and src_y[j] >= 0:
            j += 1
        kept = np.where(out_y[i:j] >= 0)[0] + i
        need[max(i, kept[0] - margin):min(j, kept[-1] + 1 + margin)] = True
        i = j
    return need


def feature_report(dev: np.ndarray, ref: np.ndarray, rows: np.ndarray | None = None) -> dict:
    """`rows`: bool mask of the canvas rows to compare (see needed_rows)."""
    if rows is not None:
        dev, ref = dev[rows], ref[rows]
    err = np.abs(dev - ref)
    # fp16 storage of 17 layers of activations against the fp32 oracle: the largest error / (1 + |ref|) measured over
    # five geometries x two frame kinds x six layers is 3.84e-3 (tools/feature_errors.py, profiles/r02_feature_errors.txt);
    # the tolerance is twice that
    tol = 8e-3 + 8e-3 * np.abs(ref)
    bad = err > tol
    rep = {"max_err": float(err.max()), "mean_err": float(err.mean()), "bad_frac": float(bad.mean()),
           "ref_absmax": float(np.abs(ref).max())}
    if bad.any():
        rows = np.where(bad.any(axis=(1, 2)))[0]
        cols = np.where(bad.any(axis=(0, 2)))[0]
        chans = np.where(bad.any(axis=(0, 1)))[0]
        rep["bad_rows"] = f"{len(rows)} rows, first {rows[:12].tolist()} last {rows[-4:].tolist()}"
        rep["bad_cols"] = f"{len(cols)} cols, first {cols[:12].tolist()} last {cols[-4:].tolist()}"
        rep["bad_chans"] = f"{len(chans)} ch, first {chans[:12].tolist()}"
    return rep
