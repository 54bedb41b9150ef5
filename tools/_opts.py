"""Experiment tools only: translate the REVE_* environment variables these scripts have always used into
reve_ctx_options (keyword arguments of reve_b200.Upscaler).  The library itself reads no environment variables."""
import os


def opts_from_env() -> dict:
    o = {}
    e = os.environ
    if e.get("REVE_CHAIN", "") != "":
        o["layers_per_launch"] = max(1, int(e["REVE_CHAIN"]))     # 0 meant "one launch per layer"
    if e.get("REVE_DEBUG_BATCH"):
        o["max_batch"] = int(e["REVE_DEBUG_BATCH"])
    if e.get("REVE_DEBUG_GRID"):
        o["debug_grid"] = int(e["REVE_DEBUG_GRID"])
    flags = int(e.get("REVE_DEBUG_FLAGS", "0") or 0)
    if e.get("REVE_CTA_PAIRS", "0") not in ("", "0"):
        flags |= 2
    if flags:
        o["debug_flags"] = flags
    t = e.get("REVE_DEBUG_TRACE", "")
    if t:
        if t == "tail":
            o["trace"] = 2
        else:
            o["trace"] = 1
            o["trace_launch"] = int(t[1:]) if t[0] == "c" else 1
            o["trace_chain"] = int(e.get("REVE_DEBUG_TRACE_CHAIN", "0"))
    if e.get("REVE_SHARED_DEVICE", "0") not in ("", "0"):
        o["shared_device"] = True
    return o
