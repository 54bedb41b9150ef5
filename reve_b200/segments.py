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


# ----------------------------------------------------------------------------------------------
# Resume file as a *set* of unfinished segments
# ----------------------------------------------------------------------------------------------
import json
from dataclasses import dataclass, field


@dataclass
class VideoState:
    """The reference's `temp\\video.temp` JSON (serde of `Video`, reve-shared/src/lib.rs:16-25): same
    field names, so a file written by reve-cli loads here and vice versa.  `segments` is the queue of
    segments that are not encoded yet (main.rs:340-343); with G GPUs it is treated as a set: segments
    finish out of order, each one is removed when its part file `video_parts\\{i}.mp4` is complete."""

    path: str
    output_path: str
    segments: List[Tuple[int, int]]          # (index, size)
    frame_rate: float
    frame_count: int
    segment_size: int
    segment_count: int
    upscale_ratio: int

    @classmethod
    def new(cls, path: str, output_path: str, frame_count: int, frame_rate: float, segment_size: int,
            upscale_ratio: int) -> "VideoState":
        segs = segment_table(frame_count, segment_size)
        return cls(path, output_path, segs, frame_rate, frame_count, segment_size, len(segs), upscale_ratio)

    @classmethod
    def from_json(cls, text: str) -> "VideoState":
        d = json.loads(text)
        return cls(d["path"], d["output_path"], [(s["index"], s["size"]) for s in d["segments"]],
                   float(d["frame_rate"]), int(d["frame_count"]), int(d["segment_size"]),
                   int(d["segment_count"]), int(d["upscale_ratio"]))

    def to_json(self) -> str:
        return json.dumps({"path": self.path, "output_path": self.output_path,
                           "segments": [{"index": i, "size": n} for i, n in self.segments],
                           "frame_rate": self.frame_rate, "frame_count": self.frame_count,
                           "segment_size": self.segment_size, "segment_count": self.segment_count,
                           "upscale_ratio": self.upscale_ratio})

    def remaining(self) -> List[int]:
        return [i for i, _ in self.segments]

    def mark_done(self, index: int) -> None:
        """Segment `index` is fully encoded: drop it from the queue (any order)."""
        before = len(self.segments)
        self.segments = [(i, n) for i, n in self.segments if i != index]
        if len(self.segments) == before:
            raise KeyError(f"segment {index} is not pending")

    def resume_fixup(self, parts_present: List[int]) -> List[int]:
        """What a resumed run must redo.  The reference re-inserts the segment before the first
        pending one because its encode may have been cut (reve-cli/src/main.rs:142-155) and deletes
        that part file (main.rs:156-159).  With out-of-order completion the rule becomes: every
        segment that is still pending, plus nothing else -- a segment is only removed from the queue
        after its part file is complete -- but part files of pending segments are stale and must be
        deleted.  Returns the indices whose part files have to be removed."""
        pending = set(self.remaining())
        return sorted(i for i in parts_present if i in pending)

    def assignment(self, world_size: int) -> List[List[int]]:
        """Pending segments per GPU (queue position k -> GPU k mod G)."""
        return [shard_segments(self.remaining(), world_size, r) for r in range(world_size)]
