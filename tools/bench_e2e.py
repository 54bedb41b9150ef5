"""Stand-in for BASELINE.json configs[4] ("end-to-end reve-cli 1080p x2 with ffmpeg decode + x265 encode overlapped,
8 B200"), as concrete as this image allows (SURVEY.md 8(d) "Config 5 -> concrete").  NOT x265 and NOT ffmpeg: neither
binary is in the image, and OpenCV here encodes mp4v / MJPG only.  What is kept is the STRUCTURE of the reference's
pipeline (reve-cli/src/main.rs:218-347: export(k+1) || upscale(k) || encode(k-1)), per GPU and in ONE process:

    decoder thread  : cv2.VideoCapture on a synthetic 1080p clip (mp4v, written once before the timed run)
                      -> RGB frame into a pinned input buffer                      [stands in for `ffmpeg -i ... rgb24`]
    upscale thread  : reve_submit / reve_wait on the GPU's context (H2D, kernels, D2H)
    consumer thread : reads the pinned 4K output: `checksum` (sum of a strided view: every frame is touched on the host)
                      or `mjpg` (cv2.VideoWriter MJPG at 4K)                         [stands in for `... -c:v libx265`]

and a where-the-time-goes table: every stage alone (frames/s on the host cores it gets), then the pipeline.
    python tools/bench_e2e.py --gpus 1|8 [--frames 240] [--sink checksum|mjpg]
GPU box only.  Prints one JSON line per measurement."""
import argparse
import json
import os
import queue
import sys
import tempfile
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# every cv2.VideoCapture brings its own ffmpeg decoder threads (one per core by default): G captures on one host
# oversubscribe it.  A handful per capture is what a real `ffmpeg -threads n` decode stage would be given.
os.environ.setdefault("OPENCV_FFMPEG_CAPTURE_OPTIONS", "threads;" + os.environ.get("REVE_DECODE_THREADS", "2"))

W, H, S = 1920, 1080, 2


def make_clip(path: str, n: int):
    import cv2
    rng = np.random.default_rng(0)
    base = np.zeros((H, W, 3), np.uint8)
    for _ in range(40):                                   # flat regions + hard edges, moving: anime-like, compressible
        x0, y0 = int(rng.integers(0, W - 200)), int(rng.integers(0, H - 200))
        base[y0:y0 + int(rng.integers(50, 400)), x0:x0 + int(rng.integers(50, 600))] = rng.integers(0, 256, 3)
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 24.0, (W, H))
    if not wr.isOpened():
        raise SystemExit("cannot open the mp4v writer")
    for i in range(n):
        wr.write(np.roll(base, (3 * i, 5 * i), axis=(0, 1)))
    wr.release()


def decode_only(path: str, n: int) -> float:
    import cv2
    cap = cv2.VideoCapture(path)
    buf = np.empty((H, W, 3), np.uint8)
    t0 = time.perf_counter()
    k = 0
    while k < n:
        ok, bgr = cap.read()
        if not ok:
            cap.set(cv2.CAP_PROP_POS_FRAMES, 0)
            continue
        cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB, dst=buf)
        k += 1
    return n / (time.perf_counter() - t0)


class Sink:
    def __init__(self, kind: str, idx: int, tmp: str):
        self.kind, self.acc, self.wr = kind, 0, None
        if kind == "mjpg":
            import cv2
            self.wr = cv2.VideoWriter(os.path.join(tmp, f"out{idx}.avi"), cv2.VideoWriter_fourcc(*"MJPG"), 24.0, (W * S, H * S))

    def consume(self, frame: np.ndarray):
        if self.wr is not None:
            self.wr.write(frame)                          # (RGB read as BGR: irrelevant for a throughput stand-in)
        else:
            self.acc += int(frame[::8, ::8].astype(np.uint32).sum())

    def close(self):
        if self.wr is not None:
            self.wr.release()


def sink_only(kind: str, n: int, tmp: str) -> float:
    frame = np.random.default_rng(1).integers(0, 256, (H * S, W * S, 3), dtype=np.uint8)
    s = Sink(kind, 99, tmp)
    t0 = time.perf_counter()
    for _ in range(n):
        s.consume(frame)
    dt = time.perf_counter() - t0
    s.close()
    return n / dt


def lane(dev: int, path: str, n: int, sink_kind: str, tmp: str, start: threading.Barrier, out: dict, stages: str):
    """stages: 'gpu' (pinned buffers cycled, no host stages), 'all' (decode -> gpu -> sink)."""
    import cv2
    import reve_b200
    if NUMA:
        from reve_b200 import numa
        numa.bind_thread_to_gpu_node(dev)                  # inherited by the decoder / consumer threads started below
    model = reve_b200.Model.for_scale(S, "models", allow_random=True, seed=1234)
    ring = 8
    with reve_b200.Upscaler(model, W, H, tile=200, prepad=10, device=dev, ring_depth=ring) as up:
        depth = ring + 4
        hin = [up.pinned((H, W, 3)) for _ in range(depth)]
        hout = [up.pinned((H * S, W * S, 3)) for _ in range(depth)]
        free_q, ready_q, done_q = queue.Queue(), queue.Queue(), queue.Queue()
        for i in range(depth):
            free_q.put(i)
        sink = Sink(sink_kind, dev, tmp)
        busy = {"decode": 0.0, "sink": 0.0}

        def decoder():
            cap = cv2.VideoCapture(path)
            k = 0
            while k < n:
                slot = free_q.get()
                t0 = time.perf_counter()
                ok, bgr = cap.read()
                if not ok:
                    cap.set(cv2.CAP_PROP_POS_FRAMES, 0)
                    ok, bgr = cap.read()
                cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB, dst=hin[slot])
                busy["decode"] += time.perf_counter() - t0
                ready_q.put(slot)
                k += 1
            ready_q.put(-1)

        def consumer():
            while True:
                slot = done_q.get()
                if slot < 0:
                    break
                t0 = time.perf_counter()
                sink.consume(hout[slot])
                busy["sink"] += time.perf_counter() - t0
                free_q.put(slot)

        for i in range(4):                                 # warm-up: clocks, first-launch costs
            up.submit(hin[i], hout[i], i)
        for i in range(4):
            up.wait()
        start.wait()
        t0 = time.perf_counter()
        if stages == "gpu":
            inflight = 0
            for k in range(n):
                if inflight == ring:
                    up.wait(); inflight -= 1
                up.submit(hin[k % ring], hout[k % ring], k % ring)
                inflight += 1
            while inflight:
                up.wait(); inflight -= 1
        else:
            td, tc = threading.Thread(target=decoder), threading.Thread(target=consumer)
            td.start(); tc.start()
            inflight, eof = 0, False
            while not eof or inflight:
                slot = None
                if not eof and inflight < ring:
                    try:
                        slot = ready_q.get(block=(inflight == 0))
                    except queue.Empty:
                        slot = None
                if slot is not None and slot < 0:
                    eof = True
                    continue
                if slot is not None:
                    up.submit(hin[slot], hout[slot], slot)
                    inflight += 1
                    continue
                done_q.put(up.wait())
                inflight -= 1
            done_q.put(-1)
            td.join(); tc.join()
        up.sync()
        out[dev] = {"s": time.perf_counter() - t0, "decode_busy_s": busy["decode"], "sink_busy_s": busy["sink"]}
        sink.close()


def run(gpus: int, path: str, n: int, sink_kind: str, tmp: str, stages: str) -> dict:
    out = {}
    start = threading.Barrier(gpus)
    ts = [threading.Thread(target=lane, args=(d, path, n, sink_kind, tmp, start, out, stages)) for d in range(gpus)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if len(out) != gpus:
        raise SystemExit("a lane failed")
    worst = max(v["s"] for v in out.values())
    return {"fps": gpus * n / worst, "per_gpu_fps": n / worst,
            "decode_busy_frac": float(np.mean([v["decode_busy_s"] / v["s"] for v in out.values()])),
            "sink_busy_frac": float(np.mean([v["sink_busy_s"] / v["s"] for v in out.values()]))}


NUMA = False


def main():
    global NUMA
    ap = argparse.ArgumentParser()
    ap.add_argument("--numa", action="store_true", help="bind each GPU's three threads to the GPU's NUMA node")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--frames", type=int, default=240, help="frames per GPU")
    ap.add_argument("--sink", default="checksum", choices=["checksum", "mjpg"])
    a = ap.parse_args()
    NUMA = a.numa
    tmp = tempfile.mkdtemp()
    clip = os.path.join(tmp, "clip.mp4")
    make_clip(clip, 48)
    cores = os.cpu_count()
    base = {"workload": "stand-in for BASELINE.json configs[4]: 1080p x2, decode + consumer overlapped with the upscale; "
                        "mp4v decode (OpenCV) and a " + a.sink + " sink -- NOT ffmpeg/x265 (absent from the image)",
            "gpus": a.gpus, "frames_per_gpu": a.frames, "host_cores": cores, "numa": a.numa,
            "decode_threads_per_capture": os.environ["OPENCV_FFMPEG_CAPTURE_OPTIONS"]}
    print(json.dumps(dict(base, stage="decode only, 1 thread", fps=round(decode_only(clip, 120), 1))), flush=True)
    print(json.dumps(dict(base, stage=f"{a.sink} sink only, 1 thread", fps=round(sink_only(a.sink, 60, tmp), 1))), flush=True)
    r = run(a.gpus, clip, a.frames, a.sink, tmp, "gpu")
    print(json.dumps(dict(base, stage="H2D + kernels + D2H only (pinned buffers, no host stages)", fps=round(r["fps"], 1),
                          per_gpu_fps=round(r["per_gpu_fps"], 1))), flush=True)
    r = run(a.gpus, clip, a.frames, a.sink, tmp, "all")
    print(json.dumps(dict(base, stage="pipeline: decode || upscale || sink, one thread each per GPU", fps=round(r["fps"], 1),
                          per_gpu_fps=round(r["per_gpu_fps"], 1), decode_thread_busy=round(r["decode_busy_frac"], 2),
                          sink_thread_busy=round(r["sink_busy_frac"], 2),
                          bound_by=("decode" if r["decode_busy_frac"] > max(0.85, r["sink_busy_frac"]) else
                                    "sink" if r["sink_busy_frac"] > 0.85 else "GPU"))), flush=True)


if __name__ == "__main__":
    main()
