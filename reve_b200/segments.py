"""Segment arithmetic of the reference and the multi-GPU sharding of segments.

The reference cuts a video into fixed-size segments (reference reve-shared/src/lib.rs:59-73) and
upscales them strictly one after another (reve-cli/src/main.rs:218-274).  Segments are independent
(no temporal state in the network, parts are only joined by the concat step, lib.rs:173-206), so on
a multi-GPU box segment k goes to GPU k mod G -- no collective, no peer traffic.
"""
from __future__ import annotations

import math
from typing import List, Tuple


def last_segment_size(frame_count: int, segment_size: int) -> int:
    """Reference reve-shared/src/lib.rs:282-289 (get_last_segment_size): the remainder minus one
    (compensating the one-frame-early seek at lib.rs:97), or a full segment if it divides evenly."""
    last = frame_count % segment_size
    return segment_size if last == 0 else last - 1


def segment_table(frame_count: int, segment_size: int) -> List[Tuple[int, int]]:
    """[(index, size)] exactly as Video::new builds it (lib.rs:59-73)."""
    if frame_count < 0 or segment_size < 1:
        raise ValueError("frame_count >= 0 and segment_size >= 1 required")
    parts = math.ceil(frame_count / segment_size) if frame_count else 0
    out = [(i, segment_size) for i in range(max(parts - 1, 0))]
    if parts:
        out.append((parts - 1, last_segment_size(frame_count, segment_size)))
    return out


def shard_segments(segment_indices: List[int], world_size: int, rank: int) -> List[int]:
    """Segments owned by `rank`: position k in the (remaining) segment queue -> GPU k mod G.
    Works on the resume queue as well (a list of not-yet-encoded indices, main.rs:340-343)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    return [s for k, s in enumerate(segment_indices) if k % world_size == rank]
