"""Segment arithmetic (mirrors the reference) and the N>1 sharding path on CPU with gloo."""
import os
import socket
import sys

import pytest

import reve_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_segment_table_matches_reference_arithmetic():
    # reference asset: test.mp4 has 1440 frames; default segment size 1000 -> 1000 + 439
    assert reve_b200.segment_table(1440, 1000) == [(0, 1000), (1, 439)]
    assert reve_b200.segment_table(2000, 1000) == [(0, 1000), (1, 1000)]
    assert reve_b200.segment_table(181, 1000) == [(0, 180)]
    assert reve_b200.segment_table(0, 1000) == []
    assert reve_b200.last_segment_size(1001, 1000) == 0          # the reference's off-by-one quirk, kept
    with pytest.raises(ValueError):
        reve_b200.segment_table(10, 0)


def test_sharding_is_a_partition():
    segs = list(range(11))
    for g in (1, 2, 4, 8):
        parts = [reve_b200.shard_segments(segs, g, r) for r in range(g)]
        assert sorted(sum(parts, [])) == segs
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    # resume queue (a set of unfinished indices) shards by queue position
    assert reve_b200.shard_segments([3, 4, 7, 9], 2, 1) == [4, 9]
    with pytest.raises(ValueError):
        reve_b200.shard_segments(segs, 2, 2)


def _worker(rank: int, world: int, port: int, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import reve_b200 as rb
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = rb.shard_segments(list(range(7)), world, rank)
        # the only cross-rank traffic of the data-parallel path: timing reduction (MAX) and a barrier
        t = torch.tensor([float(10 + rank), float(len(mine))], dtype=torch.float64)
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        dist.barrier()
        q.put((rank, mx.tolist(), gathered))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing_reduction():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, mx, gathered in res:
        assert mx[0] == 11.0 and mx[1] == 4.0                 # MAX over ranks (time, count)
        assert sorted(sum(gathered, [])) == list(range(7))    # the two shards partition the queue
        assert gathered[0] == [0, 2, 4, 6] and gathered[1] == [1, 3, 5]


def test_video_state_is_compatible_with_the_reference_json_and_behaves_as_a_set():
    # a video.temp as reve-cli writes it (serde field order of reve-shared/src/lib.rs:16-25)
    ref_json = ('{"path":"in.mkv","output_path":"out.mkv","segments":[{"index":2,"size":1000},{"index":3,"size":1000},'
                '{"index":4,"size":439}],"frame_rate":23.976,"frame_count":4440,"segment_size":1000,'
                '"segment_count":5,"upscale_ratio":2}')
    v = reve_b200.VideoState.from_json(ref_json)
    assert v.remaining() == [2, 3, 4] and v.segment_count == 5 and v.upscale_ratio == 2
    import json
    assert json.loads(v.to_json()) == json.loads(ref_json)          # round trip, same schema
    assert v.assignment(2) == [[2, 4], [3]]
    v.mark_done(3)                                                   # out-of-order completion
    assert v.remaining() == [2, 4]
    with pytest.raises(KeyError):
        v.mark_done(3)
    # resume: part files of still-pending segments are stale (the reference deletes video_parts\{first}.mp4)
    assert v.resume_fixup(parts_present=[0, 1, 2, 3]) == [2]
    n = reve_b200.VideoState.new("a.mp4", "b.mp4", 1440, 23.976, 1000, 3)
    assert n.segments == [(0, 1000), (1, 439)] and n.segment_count == 2


def test_bench_reference_arm_emits_the_contract_json():
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-budget", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "scaling", "dtype",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
