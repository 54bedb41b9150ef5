"""Aggregate host <-> device copy bandwidth of the box with all GPUs copying at once (pinned memory, one thread and one
stream per GPU), with and without binding each thread to its GPU's NUMA node.  This is the ceiling of the staged
(`e2e`) path at N GPUs: 1080p x2 moves 6.2 MB in + 24.9 MB out per frame, 720p x4 2.8 MB in + 44.2 MB out.
GPU box only:  python tools/pcie_ceiling.py [n_gpus]"""
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reve_b200 import numa  # noqa: E402


def run(n, bind, direction, mb=512, reps=8):
    res = {}
    start = threading.Barrier(n)

    def lane(d):
        if bind:
            numa.bind_thread_to_gpu_node(d)
        torch.cuda.set_device(d)
        host = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True)
        host.fill_(1)
        dev = torch.empty(mb << 20, dtype=torch.uint8, device=f"cuda:{d}")
        st = torch.cuda.Stream(device=d)
        with torch.cuda.stream(st):
            (host if direction == "d2h" else dev).copy_(dev if direction == "d2h" else host, non_blocking=True)
        st.synchronize()
        start.wait()
        t0 = time.perf_counter()
        with torch.cuda.stream(st):
            for _ in range(reps):
                (host if direction == "d2h" else dev).copy_(dev if direction == "d2h" else host, non_blocking=True)
        st.synchronize()
        res[d] = time.perf_counter() - t0

    ts = [threading.Thread(target=lane, args=(d,)) for d in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    worst = max(res.values())
    return n * reps * mb / 1024.0 / worst


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    for g in sorted({1, 2, 4, n} & set(range(1, n + 1))):
        for direction in ("d2h", "h2d"):
            for bind in (False, True):
                print(json.dumps({"gpus": g, "direction": direction, "numa_bound": bind,
                                  "aggregate_GBps": round(run(g, bind, direction), 1)}), flush=True)
