"""reve_b200 -- B200-native (sm_100a) implementation of REVE's per-segment upscale step.

The product is ``libreve_cuda.so`` (hand-written CUDA behind the C ABI in ``include/reve_cuda.h``);
this package is the thin Python host mirror used by the tests and ``bench.py``.  It never falls
back to a CPU path: if the shared library or an sm_100 device is missing, it raises.
"""
from .segments import VideoState, last_segment_size, segment_table, shard_segments
from .upscaler import (CTX_SHARED_DEVICE, DBG_ALIAS_ROWS, DBG_ALL_ROWS, DBG_CONV0_IM2COL, DBG_CTA_PAIRS, DBG_EQUAL_SPLIT, DBG_FAULT, DBG_NO_REVERSE,
                       DBG_SWAP_PAIR_B, FMT_RGB24, FMT_YUV420P10LE_BT601, FMT_YUV420P10LE_BT709, Model, Upscaler, ReveError,
                       load_library, library_path, geometry, launch_plan, upscale_segment)

__all__ = ["CTX_SHARED_DEVICE", "DBG_ALIAS_ROWS", "DBG_ALL_ROWS", "DBG_CONV0_IM2COL", "DBG_CTA_PAIRS", "DBG_EQUAL_SPLIT", "DBG_FAULT", "DBG_NO_REVERSE",
           "DBG_SWAP_PAIR_B", "Model", "Upscaler", "ReveError", "load_library", "library_path", "geometry", "launch_plan",
           "upscale_segment", "FMT_RGB24", "FMT_YUV420P10LE_BT601", "FMT_YUV420P10LE_BT709", "segment_table", "last_segment_size", "shard_segments", "VideoState"]
