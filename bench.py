#!/usr/bin/env python
"""bench.py -- frames/s of the realesr-animevideov3 upscale step on B200 (default: x2 1080p->4K, BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --gpus N --single-process                      # the product's shape: ONE process, one thread + context per GPU
    python bench.py --workload 720p_x4|540p_x3|480p_x2             # the other BASELINE.json geometries
    python bench.py --impl reference [...]                         # the reference's CPU path (restated)

A "step" is one pass of the hot path over one batch of `--batch` synthetic frames.  Protocol:
W >= 3 warm-up steps, then THREE timed regions of exactly K steps each (>= 200 frames per region with the defaults,
K x batch = 20 x 12), bracketed by a barrier + synchronize, timed with CUDA events on the library's compute stream, MAX
over ranks; `value` is the MEDIAN region (all three are listed under `regions`).
  value     whole-job frames/s with the frames resident in HBM (reve_upscale_device)
  e2e       the same metric through the reference-facing C-ABI call (reve_submit / reve_wait) with pinned HOST buffers,
            H2D and D2H copies inside the timed region (median of three regions as well)
  roofline  the dominant kernel (chained 64->64 tcgen05 body convolution), every launch of the timed region bracketed by
            CUDA events; `roofline_others` carries conv0 and the tail against the measured HBM bandwidth
  latency   one frame, ring depth 1: reve_submit -> reve_wait wall time (H2D + 6 launches + D2H)
  cpu_baseline / --impl reference: the CPU restatement of the reference path (oracle/srvgg.py, torch CPU fp32, upstream
            tile 200 / pre-pad 10) on the box's host cores on a bounded sample; the real realesrgan-ncnn-vulkan binary and
            its weights are not available offline.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {   # name -> (w, h, scale, BASELINE.json config it belongs to)
    "1080p_x2": (1920, 1080, 2, "configs[1]: realesr-animevideov3 x2 1080p->4K, 1000-frame segment, 1 B200"),
    "720p_x4": (1280, 720, 4, "configs[2]: realesr-animevideov3 x4 720p->2880p, segments sharded across the GPUs"),
    "540p_x3": (960, 540, 3, "configs[3]: realesr-animevideov3 x3 540p->1620p small-frame path"),
    "480p_x2": (640, 480, 2, "configs[0] geometry: realesr-animevideov3 x2 480p (the reference's demo asset size)"),
}
FLOP_PER_PX = {2: 1196928, 3: 1214208, 4: 1238400}   # SURVEY.md section 8(d), algorithmic
BODY_FLOP_PER_PX = 2 * 9 * 64 * 64                   # one 64->64 3x3 layer
WEIGHTS_NOTE = "seeded He-normal random init of SRVGGNetCompact (real .param/.bin unavailable offline)"


def metric_name(workload: str) -> str:
    w, h, s, _ = WORKLOADS[workload]
    return "frames/s animevideov3 x2 1080p->4K" if workload == "1080p_x2" else f"frames/s animevideov3 x{s} {h}p->{h * s}p"


def make_frame(kind: str, w: int, h: int, seed: int) -> np.ndarray:
    """Synthetic input frames.  'noise' (default, and what every committed number uses unless it says otherwise): uniform u8
    noise -- the worst case for an energy-bound kernel, every operand bit toggles.  'edges': anime-like frames
    (a gradient, flat regions with hard edges, line art).  'real': the decoded frame of the reference's demo clip held by
    tests/golden (640x480), tiled and shifted to the workload's size.  The arithmetic is identical; what changes is the
    switching activity, hence the power, hence the clock the 1000 W cap allows (profiles/r02_notes.md section 17)."""
    if kind == "noise":
        return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == "edges":
        rng = np.random.default_rng(seed)
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.empty((h, w, 3), np.uint8)
        img[..., 0] = xx * 255 // max(1, w - 1)                      # a gradient background
        img[..., 1] = yy * 255 // max(1, h - 1)
        img[..., 2] = 128
        for _ in range(40):                                          # flat regions with hard edges
            x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
            img[y0:y0 + int(rng.integers(h // 20, h // 3)), x0:x0 + int(rng.integers(w // 20, w // 3))] = rng.integers(0, 256, 3)
        img[(xx // 7 + yy // 5) % 11 == 0] = 0                       # line art
        return img
    g = np.load(os.path.join(ROOT, "tests", "golden", "x2_tile200_real_onepiece_480p.npz"))
    src = g["frame"]
    reps = (-(-h // src.shape[0]) + 1, -(-w // src.shape[1]) + 1, 1)
    big = np.tile(src, reps)
    dy, dx = (seed * 37) % src.shape[0], (seed * 101) % src.shape[1]
    return np.ascontiguousarray(big[dy:dy + h, dx:dx + w])


def config_block(args, world: int, parallelism: str) -> dict:
    """The same keys on both arms (ours / reference), so that the driver's same-config check compares like with like."""
    w, h, s, cfg = WORKLOADS[args.workload]
    return {"workload": f"{args.workload}: {cfg}", "frame": [w, h], "scale": s, "tile": args.tile, "prepad": args.prepad,
            "frames_per_step": args.batch, "ring_depth": 8, "weights": WEIGHTS_NOTE, "frames": getattr(args, "frames", "noise"),
            "l2": f"working set per launch set (2 fp16 activation canvases of 4 stacked frames) exceeds the 126 MB L2; "
                  f"{args.batch} distinct frames cycled",
            "parallelism": parallelism}


def parallelism_string(args, n_gpus: int) -> str:
    """Identical on both arms for the same command line (the reference arm describes the configuration of OUR arm)."""
    if n_gpus <= 1:
        return "segments x1, no collective"
    mode = ("one process, one thread + reve_ctx per GPU (the product's shape)" if args.single_process else
            f"one rank per GPU under torchrun ({args.pg} process group for barrier / MAX only)")
    return f"segments x{n_gpus}, no collective; {mode}"


def measured_traffic(chained=False):
    """dram__bytes_read.sum + dram__bytes_write.sum of one body launch from the committed ncu capture."""
    for name in (("r02_chain_traffic.json", "r01_chain_traffic.json") if chained else ("r01_body_traffic.json",)):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f)
        except OSError:
            continue
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # median over the samples taken under load (power above half of the observed maximum)
        load = [c for c, p in zip(sm, pw) if pw and p >= 0.5 * max(pw)]
        return {"sm_mhz": float(np.median(load)) if load else (float(np.median(sm)) if sm else None),
                "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(pw)) if pw else None,
                "samples": len(sm), "samples_under_load": len(load), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_sample_fps(budget_s: float, w_in: int, h_in: int, scale: int, tile: int, prepad: int, seed: int = 0):
    """Times the CPU restatement (oracle) on as many upstream tiles of one frame as fit in `budget_s`, returns
    (frames/s, cores, description).  Uses all host threads."""
    import torch
    from oracle import srvgg
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wts = srvgg.make_weights(scale, 1234)
    frame = srvgg.synthetic_frame(w_in, h_in, seed, "random")
    tiles = srvgg.tile_grid(w_in, h_in, tile)
    # warm-up on one tile (thread pool, oneDNN primitive cache)
    x0, y0, tw, th = tiles[0]
    t = srvgg.padded_tile(frame, x0, y0, tw, th, prepad)
    xin = (t.astype(np.float32) * np.float32(1 / 255.0)).transpose(2, 0, 1)
    srvgg.forward(xin, wts)
    done_px, n = 0, 0
    t0 = time.perf_counter()
    rounds = 0
    while True:
        for (x0, y0, tw, th) in tiles:
            t = srvgg.padded_tile(frame, x0, y0, tw, th, prepad)
            xin = (t.astype(np.float32) * np.float32(1 / 255.0)).transpose(2, 0, 1)
            y = srvgg.forward(xin, wts)
            srvgg.quantise(y)
            done_px += tw * th
            n += 1
            if time.perf_counter() - t0 > budget_s:
                break
        rounds += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    frames = done_px / float(w_in * h_in)
    return frames / dt, cores, (f"{n} upstream tiles (tile {tile}, pre-pad {prepad}; a frame has {len(tiles)}) of a synthetic "
                                f"{w_in}x{h_in} frame = {frames:.3f} frames in {dt:.1f} s, torch CPU fp32, {cores} threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w_in, h_in, scale, _ = WORKLOADS[args.workload]
    total = max(1, args.steps + args.warmup)
    per_step = min(8.0, 150.0 / total)
    vals, desc, cores = [], "", 1
    for i in range(total):
        fps, cores, desc = cpu_sample_fps(per_step, w_in, h_in, scale, args.tile, args.prepad, seed=i)
        if i >= args.warmup:
            vals.append(fps)
    v = float(np.mean(vals)) if vals else 0.0
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * args.batch / v if v else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(args, args.gpus, parallelism_string(args, args.gpus)),
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "per step: " + desc},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of the reference path (oracle/srvgg.py); realesrgan-ncnn-vulkan itself is absent offline; "
                "each step is a bounded sample (a few seconds) of the workload, not a whole batch",
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Lane:
    """One GPU: context, resident frames, pinned ring.  Used by the torchrun rank and by every thread of the
    single-process mode alike."""

    def __init__(self, args, dev: int, seed_base: int):
        import torch
        import reve_b200
        self.torch = torch
        self.args = args
        self.dev = dev
        self.w, self.h, self.scale, _ = WORKLOADS[args.workload]
        torch.cuda.set_device(dev)
        self.numa_cpus = None
        self.model = reve_b200.Model.for_scale(self.scale, args.model_dir, allow_random=True, seed=1234)
        self.up = reve_b200.Upscaler(self.model, self.w, self.h, tile=args.tile, prepad=args.prepad, device=dev, ring_depth=8,
                                     shared_device=args.shared_device)
        B = args.batch
        self.frames = np.stack([make_frame(args.frames, self.w, self.h, seed_base + i) for i in range(min(B, 8))])
        self.nres = len(self.frames)
        self.d_in = torch.from_numpy(self.frames).to(f"cuda:{dev}")
        self.d_out = torch.empty((self.nres, self.h * self.scale, self.w * self.scale, 3), dtype=torch.uint8, device=f"cuda:{dev}")
        self.stream = torch.cuda.ExternalStream(self.up.stream, device=f"cuda:{dev}")
        self.ring = self.up.ring_depth
        self.h_in = self.h_out = None

    def alloc_pinned(self):
        """Called from the lane's own thread (after it has bound itself to the GPU's NUMA node, --numa): the pinned ring is
        first touched, hence placed, there."""
        self.h_in = [self.up.pinned((self.h, self.w, 3)) for _ in range(self.ring)]
        self.h_out = [self.up.pinned((self.h * self.scale, self.w * self.scale, 3)) for _ in range(self.ring)]
        for i in range(self.ring):
            self.h_in[i][...] = self.frames[i % self.nres]

    def step_device(self):
        # one step = `batch` frames: the resident frames cycled (a pass over 8 distinct frames + the first ones again)
        B, n = self.args.batch, self.nres
        for f0 in range(0, B, n):
            self.up.upscale_device(self.d_in.data_ptr(), self.d_out.data_ptr(), min(n, B - f0))

    def timed_device(self, steps: int) -> float:
        torch = self.torch
        with torch.cuda.device(self.dev):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(self.stream)
            for _ in range(steps):
                self.step_device()
            ev1.record(self.stream)
            torch.cuda.synchronize(self.dev)
            return ev0.elapsed_time(ev1)

    def run_e2e(self, n_frames: int):
        """The segment is streamed: up to `ring` frames in flight (H2D of frame i+k overlaps the kernels of frame i and the
        D2H of frame i-k), exactly as a decode -> upscale -> encode pipeline drives it."""
        up, ring = self.up, self.ring
        inflight = 0
        for i in range(n_frames):
            if inflight == ring:
                up.wait(); inflight -= 1
            up.submit(self.h_in[i % ring], self.h_out[i % ring], i)
            inflight += 1
        while inflight:
            up.wait(); inflight -= 1
        up.sync()

    def latency_ms(self, reps: int = 20) -> float:
        lat = []
        for _ in range(reps):
            t0 = time.perf_counter()
            self.up.submit(self.h_in[0], self.h_out[0], 0)
            self.up.wait()
            lat.append((time.perf_counter() - t0) * 1e3)
        return float(np.median(lat))

    def close(self):
        self.up.close()


def measure(lanes, args, barrier, reduce_max, rank0: bool):
    """Runs the protocol on this process's lanes (1 under torchrun, G in single-process mode; lanes run in threads).
    Returns the per-process measurements; `reduce_max` folds region times over ranks."""
    K, Wm, B = args.steps, args.warmup, args.batch
    R = 3
    results = [None] * len(lanes)
    tb = threading.Barrier(len(lanes))
    sampler_box = {}

    errors = []

    def lane_main(li):
        try:
            lane_body(li)
        except BaseException as e:      # noqa: BLE001  -- a failed lane must not leave the others waiting at a barrier
            errors.append(f"lane {li}: {e!r}")
            tb.abort()

    def lane_body(li):
        lane = lanes[li]
        torch = lane.torch
        torch.cuda.set_device(lane.dev)
        if args.numa:
            from reve_b200 import numa
            lane.numa_cpus = numa.bind_thread_to_gpu_node(lane.dev)
        lane.alloc_pinned()
        out = {"numa_cpus": lane.numa_cpus}
        for _ in range(Wm):
            lane.step_device()
        torch.cuda.synchronize(lane.dev)
        lane.up.profile(reset=True)
        lane.up.set_profiling(True)
        tb.wait()
        if li == 0:
            barrier()
            if rank0:
                sampler_box["s"] = ClockSampler(lane.dev)
        tb.wait()
        regions = []
        for _ in range(R):
            tb.wait()
            if li == 0:
                barrier()
            tb.wait()
            regions.append(lane.timed_device(K))
        tb.wait()
        if li == 0 and "s" in sampler_box:
            sampler_box["clocks"] = sampler_box["s"].stop()
        out["regions_ms"] = regions
        out["prof"] = lane.up.profile(reset=True)
        lane.up.set_profiling(False)
        # ---- end to end
        lane.run_e2e(min(Wm, 3) * B)
        e2e = []
        for _ in range(R):
            tb.wait()
            if li == 0:
                barrier()
            tb.wait()
            t0 = time.perf_counter()
            lane.run_e2e(K * B)
            e2e.append((time.perf_counter() - t0) * 1e3)
        out["e2e_ms"] = e2e
        out["checksum"] = int(lane.h_out[0][::97, ::89].astype(np.int64).sum())   # a D2H result actually read on the host
        tb.wait()
        if li == 0:
            out["latency_ms"] = lane.latency_ms()
        out["launch_info"] = lane.up.launch_info()
        results[li] = out

    threads = [threading.Thread(target=lane_main, args=(i,)) for i in range(len(lanes))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors or any(r is None for r in results):
        raise SystemExit("bench.py: " + ("; ".join(errors) or "a lane failed"))
    # per region: the slowest lane of this process, then the slowest rank
    regions = [max(r["regions_ms"][k] for r in results) for k in range(R)]
    e2e = [max(r["e2e_ms"][k] for r in results) for k in range(R)]
    regions = reduce_max(regions)
    e2e = reduce_max(e2e)
    return results, regions, e2e, sampler_box.get("clocks")


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    dist = None
    single = args.single_process and world == 1
    if world > 1:
        # The data path has no collective; the process group exists only for the barrier and the MAX
        # reduction of the timings the bench contract asks for.  NCCL because the driver launches one rank per GPU "over
        # NCCL" and checks the communicator's rank count; --pg gloo does the same reductions on the host.
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if args.pg == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    n_gpus = args.gpus if single else world
    devs = list(range(args.gpus)) if single else [local]
    if single and torch.cuda.device_count() < args.gpus:
        raise SystemExit(f"bench.py: --single-process --gpus {args.gpus} needs {args.gpus} devices")
    lanes = [Lane(args, d, 100 * (rank * 8 + i)) for i, d in enumerate(devs)]   # each GPU its own segment (seeds)

    def barrier():
        for ln in lanes:
            torch.cuda.synchronize(ln.dev)
        if dist is not None:
            dist.barrier()

    def reduce_max(vals):
        if dist is None:
            return vals
        t = torch.tensor(vals, dtype=torch.float64, device=f"cuda:{local}" if args.pg == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    results, regions, e2e, clocks = measure(lanes, args, barrier, reduce_max, rank == 0)
    K, Wm, B = args.steps, args.warmup, args.batch
    launches = sum(r["prof"]["launches_conv0"] + r["prof"]["launches_body"] + r["prof"]["launches_tail"] for r in results)
    body_ms = max(r["prof"]["ms_body"] / max(1, r["prof"]["timed_body"]) for r in results)
    if dist is not None:
        dev_t = f"cuda:{local}" if args.pg == "nccl" else "cpu"
        lt = torch.tensor([launches], dtype=torch.int64, device=dev_t)
        dist.all_reduce(lt)
        launches = int(lt.item())
        bt = torch.tensor([body_ms], dtype=torch.float64, device=dev_t)
        dist.all_reduce(bt, op=dist.ReduceOp.MAX)
        body_ms = float(bt.item())

    if rank == 0:
        w_in, h_in, scale, _ = WORKLOADS[args.workload]
        pr = results[0]["prof"]
        peaks, psrc = measured_peaks()
        ms = float(np.median(regions))
        e2e_ms = float(np.median(e2e))
        fps = n_gpus * K * B / (ms / 1000.0)
        e2e_fps = n_gpus * K * B / (e2e_ms / 1000.0)
        px = w_in * h_in
        in_bytes, out_bytes = px * 3, px * 3 * scale * scale
        frames_timed = max(1, pr["frames"])
        frames_per_launch = pr["body_frames"] / max(1, pr["launches_body"])        # frames stacked per launch
        layers_per_launch = pr["body_layer_frames"] / max(1, pr["body_frames"])     # chained launches run 2 or 4 layers
        flop_per_launch = BODY_FLOP_PER_PX * px * frames_per_launch * layers_per_launch
        body_tflops = flop_per_launch / (body_ms * 1e-3) / 1e12 if body_ms > 0 else 0.0
        chained = layers_per_launch > 1.5
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        frame_tflops = fps / n_gpus * FLOP_PER_PX[scale] * px / 1e12
        import reve_b200
        cw, chh, *_ = reve_b200.geometry(w_in, h_in, scale, args.tile, args.prepad)
        canvas_px = cw * chh
        # algorithmic bytes of the two HBM-bound kernels, per frame: conv0 reads the u8 frame and writes the
        # fp16 NHWC canvas; the tail reads that canvas and the u8 frame (residual) and writes the u8 output
        conv0_bytes = in_bytes + canvas_px * 128
        tail_bytes = canvas_px * 128 + in_bytes + out_bytes
        conv0_ms, tail_ms = pr["ms_conv0"] / frames_timed, pr["ms_tail"] / frames_timed
        line = {
            "metric": metric_name(args.workload), "value": fps, "unit": "frames/s", "n_gpus": n_gpus, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": config_block(args, n_gpus, parallelism_string(args, n_gpus)),
            "regions": {"count": len(regions), "frames_each": K * B, "ms": regions, "e2e_ms": e2e, "reported": "median"},
            "output_mpixel_per_s": fps * px * scale * scale / 1e6,
            "frame_tflops_algorithmic": frame_tflops,
            "frame_frac_of_peak": frame_tflops / peak,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": B * in_bytes,
                    "d2h_bytes_per_step": B * out_bytes, "api": "reve_submit/reve_wait, pinned host buffers",
                    "host_checksum": results[0]["checksum"]},
            "latency_ms_1frame": results[0].get("latency_ms"),
            "host": {"cores": os.cpu_count(), "numa_binding": [r.get("numa_cpus") for r in results] if args.numa else "off"},
            "gpu_launches": launches,
            "launch": results[0]["launch_info"],
            "roofline": {"kernel": (f"conv3x3_chain_kernel ({layers_per_launch:.0f} chained 64->64 3x3 + PReLU layers per launch, tcgen05, "
                                    "rotating TMEM banks, layer-to-layer hand-over through L2 scratch rings, cooperative launch)") if chained else
                                   "conv3x3_umma_kernel<64,false,false> (64->64 3x3 + PReLU, tcgen05, rotating TMEM banks)",
                         "bound": "tensor", "achieved": body_tflops, "peak": peak, "unit": "TFLOP/s",
                         "frac": body_tflops / peak,
                         "traffic": (lambda t: None if not t else t["dram_bytes_per_launch"] * frames_per_launch / t["frames_per_launch"])(measured_traffic(chained)),
                         "traffic_unit": "bytes of DRAM read+write per launch (ncu capture under profiles/); algorithmic activation "
                                         "bytes per launch = 2 x 128 B x canvas pixels (one canvas read, one written)",
                         "dram_bytes_per_frame": (lambda t: None if not t or "dram_bytes_per_frame" not in t or args.workload != "1080p_x2" else
                                                  {"measured_ncu": t["dram_bytes_per_frame"], "algorithmic_minimum": t["algorithmic_minimum_bytes_per_frame"],
                                                   "note": "all six launches of a frame; the minimum is u8 in + u8 out, everything above it is fp16 activations "
                                                           "crossing HBM at the 4 chain boundaries, conv0's output and the tail's input"})(measured_traffic(chained)),
                         "peak_source": f"{psrc} bf16_tflops_sustained (kernel timed inside a long step)",
                         "avg_launch_ms": body_ms, "launches_timed": int(pr["timed_body"]),
                         "timed_in": "the three timed regions of `value` (every launch bracketed by CUDA events)",
                         "kernel_ms_share_of_step": (pr["ms_conv0"] + pr["ms_body"] + pr["ms_tail"]) / sum(results[0]["regions_ms"]),
                         "algorithmic_flop_per_launch": flop_per_launch,
                         "frames_per_launch": frames_per_launch, "layers_per_launch": layers_per_launch,
                         "ms_per_frame": {"conv0": conv0_ms, "body_x16": pr["ms_body"] / frames_timed, "tail": tail_ms}},
            "roofline_others": {
                "conv0_rows_kernel": {"bound": "hbm", "achieved": conv0_bytes / (conv0_ms * 1e-3) / 1e9 if conv0_ms else None,
                                      "peak": hbm, "unit": "GB/s", "frac": conv0_bytes / (conv0_ms * 1e-3) / 1e9 / hbm if conv0_ms else None,
                                      "algorithmic_bytes_per_frame": conv0_bytes,
                                      "note": "a write stream (128 B written, 3 B read per canvas pixel); the write-only ceiling with this access "
                                              "pattern is 6.06 TB/s (profiles/r02_store_paths.txt)"},
                "conv3x3_umma_kernel<tail>": {"bound": "hbm", "achieved": tail_bytes / (tail_ms * 1e-3) / 1e9 if tail_ms else None,
                                              "peak": hbm, "unit": "GB/s", "frac": tail_bytes / (tail_ms * 1e-3) / 1e9 / hbm if tail_ms else None,
                                              "algorithmic_bytes_per_frame": tail_bytes},
                "peak_source": f"{psrc} hbm_gbs"},
            "clocks": clocks,
        }
        if not args.no_cpu and n_gpus == 1:
            v, cores, desc = cpu_sample_fps(args.cpu_budget, w_in, h_in, scale, 200, 10)
            line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc}
            if args.workload == "1080p_x2":   # BASELINE.json configs[0]: the reference's own CPU-runnable case (480p segment)
                v0, cores0, desc0 = cpu_sample_fps(min(args.cpu_budget, 10.0), 640, 480, 2, 200, 10)
                line["cpu_baseline_config0"] = {"value": v0, "unit": "frames/s (640x480 x2)", "cores": cores0, "kind": "port",
                                                "sample": desc0 + "; BASELINE.json configs[0] is a 100-frame segment of these"}
        else:
            line["cpu_baseline"] = None
        emit(line)
    for ln in lanes:
        ln.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_JSON_OUT = None


def quiet_stdout():
    """Libraries print to stdout (NCCL: 'NCCL version ...' when the first communicator is built); the contract is ONE
    JSON line there.  File descriptor 1 is pointed at stderr for the whole run and the line goes to the saved one."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1080p_x2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=12, help="frames per step (20 steps x 12 = 240 frames per timed region)")
    ap.add_argument("--frames", default="noise", choices=["noise", "edges", "real"],
                    help="synthetic input: uniform noise (default, worst case for power), anime-like flat regions + line art, or the reference's demo frame tiled")
    ap.add_argument("--tile", type=int, default=200, help="upstream tile size (0 = whole frame)")
    ap.add_argument("--prepad", type=int, default=10)
    ap.add_argument("--model-dir", default="models")
    ap.add_argument("--single-process", action="store_true",
                    help="N GPUs from ONE process: one thread + one context per GPU, no process group at all")
    ap.add_argument("--pg", default="nccl", choices=["nccl", "gloo"], help="process group for barrier / MAX under torchrun")
    ap.add_argument("--shared-device", action="store_true", help="REVE_CTX_SHARED_DEVICE: single-layer launches only")
    ap.add_argument("--numa", action="store_true",
                    help="bind every lane's thread (and so its pinned ring) to the CPUs of its GPU's NUMA node")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    relaunch = args.impl == "ours" and args.gpus > 1 and world == 1 and not args.single_process
    if not relaunch:   # (the re-launching parent passes its children's stdout through)
        quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if relaunch:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
