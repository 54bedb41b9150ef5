#!/usr/bin/env python
"""bench.py -- frames/s of the realesr-animevideov3 x2 1080p->4K upscale step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                         # the reference's CPU path

A "step" is one pass of the hot path over one batch of `--batch` synthetic 1080p frames
(BASELINE.json configs[1]).  `value` = whole-job frames/s with the frames resident in HBM
(reve_upscale_device); `e2e` = the same metric through the reference-facing C-ABI call
(reve_submit / reve_wait) with pinned HOST buffers, H2D and D2H copies inside the timed region.
`roofline` is the dominant kernel (the 64->64 tcgen05 body convolution, 16 launches per frame)
timed live with CUDA events on the library's compute stream.  `cpu_baseline` / `--impl reference`
time the CPU restatement of the reference path (oracle/srvgg.py, torch CPU fp32, upstream tile 200
/ pre-pad 10) on the box's host cores on a bounded sample; the real realesrgan-ncnn-vulkan binary
and its weights are not available offline.
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

W_IN, H_IN, SCALE = 1920, 1080, 2
FLOP_PER_PX = {2: 1196928, 3: 1214208, 4: 1238400}   # SURVEY.md section 8(d), algorithmic
BODY_FLOP_PER_PX = 2 * 9 * 64 * 64                   # one 64->64 3x3 layer
METRIC = "frames/s animevideov3 x2 1080p->4K"


def measured_traffic(chained=False):
    """dram__bytes_read.sum + dram__bytes_write.sum of one body launch from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "r01_chain_traffic.json" if chained else "r01_body_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except OSError:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
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
        # under-load samples: the upper half of the observed clocks (idle samples drag the median)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_sample_fps(budget_s: float, tile: int, prepad: int, seed: int = 0):
    """Times the CPU restatement (oracle) on as many upstream tiles of one 1080p frame as fit in
    `budget_s`, returns (frames/s, cores, description).  Uses all host threads."""
    import torch
    from oracle import srvgg
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wts = srvgg.make_weights(SCALE, 1234)
    frame = srvgg.synthetic_frame(W_IN, H_IN, seed, "random")
    tiles = srvgg.tile_grid(W_IN, H_IN, tile)
    # warm-up on one tile (thread pool, oneDNN primitive cache)
    x0, y0, tw, th = tiles[0]
    t = srvgg.padded_tile(frame, x0, y0, tw, th, prepad)
    xin = (t.astype(np.float32) * np.float32(1 / 255.0)).transpose(2, 0, 1)
    srvgg.forward(xin, wts)
    done_px, n = 0, 0
    t0 = time.perf_counter()
    for (x0, y0, tw, th) in tiles:
        t = srvgg.padded_tile(frame, x0, y0, tw, th, prepad)
        xin = (t.astype(np.float32) * np.float32(1 / 255.0)).transpose(2, 0, 1)
        y = srvgg.forward(xin, wts)
        srvgg.quantise(y)
        done_px += tw * th
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    frames = done_px / float(W_IN * H_IN)
    return frames / dt, cores, (f"{n} of {len(tiles)} upstream tiles (tile {tile}, pre-pad {prepad}) of one synthetic "
                                f"1080p frame = {frames:.3f} frame in {dt:.1f} s, torch CPU fp32, {cores} threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    total = max(1, args.steps + args.warmup)
    per_step = min(8.0, 150.0 / total)
    vals, desc, cores = [], "", 1
    for i in range(total):
        fps, cores, desc = cpu_sample_fps(per_step, args.tile, args.prepad, seed=i)
        if i >= args.warmup:
            vals.append(fps)
    v = float(np.mean(vals)) if vals else 0.0
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v if v else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "realesr-animevideov3 x2 1080p->4K, upstream tile 200 / pre-pad 10",
                   "frame": [W_IN, H_IN], "scale": SCALE, "tile": args.tile, "prepad": args.prepad,
                   "weights": "seeded He-normal random init of SRVGGNetCompact (real .param/.bin unavailable offline)"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "per step: " + desc},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of the reference path (oracle/srvgg.py); realesrgan-ncnn-vulkan itself is absent offline",
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import reve_b200

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local
    torch.cuda.set_device(dev)

    B, K, Wm = args.batch, args.steps, args.warmup
    model = reve_b200.Model.for_scale(SCALE, args.model_dir, seed=1234)
    up = reve_b200.Upscaler(model, W_IN, H_IN, tile=args.tile, prepad=args.prepad, device=dev, ring_depth=8)
    in_bytes, out_bytes = W_IN * H_IN * 3, W_IN * H_IN * 3 * SCALE * SCALE

    # synthetic segment: B distinct seeded frames (each rank its own seeds = its own segment)
    frames = np.stack([np.random.default_rng(100 * rank + i).integers(0, 256, (H_IN, W_IN, 3), dtype=np.uint8)
                       for i in range(B)])
    d_in = torch.from_numpy(frames).to(f"cuda:{dev}")
    d_out = torch.empty((B, H_IN * SCALE, W_IN * SCALE, 3), dtype=torch.uint8, device=f"cuda:{dev}")
    stream = torch.cuda.ExternalStream(up.stream, device=f"cuda:{dev}")

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device():
        up.upscale_device(d_in.data_ptr(), d_out.data_ptr(), B)

    # ---- device-resident throughput (value) -------------------------------------------------
    # Every launch of the timed region is bracketed by CUDA events on the library's compute stream
    # (reve_ctx_set_profiling), so `value` and the roofline's per-kernel durations come from the SAME
    # sustained, power-capped region.  (A short separate profiling pass after a pause runs at burst
    # clocks and overstates the kernel by 10-18 %; --no-prof-in-timed-region restores that behaviour
    # to measure what the events cost: nothing measurable.)
    for _ in range(Wm):
        step_device()
    barrier()
    up.profile(reset=True)
    prof_in_region = not args.no_prof_in_timed_region
    if prof_in_region:
        up.set_profiling(True)
    sampler = ClockSampler(dev) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(K):
        step_device()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    pr = up.profile(reset=True)
    launches = pr["launches_conv0"] + pr["launches_body"] + pr["launches_tail"]
    if not prof_in_region:
        up.set_profiling(True)
        for _ in range(max(1, min(K, 4))):
            step_device()
        pr = up.profile(reset=True)
    up.set_profiling(False)
    body_ms = pr["ms_body"] / max(1, pr["timed_body"])
    frames_timed = max(1, pr["frames"])
    frames_per_launch = pr["body_frames"] / max(1, pr["launches_body"])   # frames stacked per launch
    layers_per_launch = pr["body_layer_frames"] / max(1, pr["body_frames"])  # chained launches run 2 or 4 layers

    # ---- end to end through reve_submit / reve_wait with pinned host buffers -------------------
    ring = up.ring_depth
    h_in = [up.pinned((H_IN, W_IN, 3)) for _ in range(ring)]
    h_out = [up.pinned((H_IN * SCALE, W_IN * SCALE, 3)) for _ in range(ring)]
    for i in range(ring):
        h_in[i][...] = frames[i % B]

    # The segment is streamed: up to `ring` frames are in flight (H2D of frame i+k overlaps the kernels
    # of frame i and the D2H of frame i-k), exactly as a decode -> upscale -> encode pipeline drives it.
    def run_e2e(n_frames):
        inflight = 0
        for i in range(n_frames):
            if inflight == ring:
                up.wait(); inflight -= 1
            up.submit(h_in[i % ring], h_out[i % ring], i)
            inflight += 1
        while inflight:
            up.wait(); inflight -= 1

    run_e2e(min(Wm, 3) * B)
    barrier()
    t0 = time.perf_counter()
    run_e2e(K * B)
    up.sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    checksum = int(h_out[0][::97, ::89].astype(np.int64).sum())  # a D2H result actually read on the host

    # ---- reduce over ranks -----------------------------------------------------------------------
    if dist is not None:
        t = torch.tensor([ms, e2e_s * 1000.0, body_ms], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, body_ms = (float(x) for x in t.tolist())
        e2e_s = e2e_ms / 1000.0
        lt = torch.tensor([launches], dtype=torch.int64, device=f"cuda:{dev}")
        dist.all_reduce(lt)
        launches = int(lt.item())

    if rank == 0:
        peaks, psrc = measured_peaks()
        fps = world * K * B / (ms / 1000.0)
        e2e_fps = world * K * B / e2e_s
        px = W_IN * H_IN
        flop_per_launch = BODY_FLOP_PER_PX * px * frames_per_launch * layers_per_launch
        body_tflops = flop_per_launch / (body_ms * 1e-3) / 1e12 if body_ms > 0 else 0.0
        chained = layers_per_launch > 1.5
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        frame_tflops = fps / world * FLOP_PER_PX[SCALE] * px / 1e12
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": "realesr-animevideov3 x2 1080p->4K (BASELINE.json configs[1]), one segment per GPU",
                       "frame": [W_IN, H_IN], "scale": SCALE, "tile": args.tile, "prepad": args.prepad,
                       "frames_per_step": B, "ring_depth": ring,
                       "weights": "seeded He-normal random init of SRVGGNetCompact (real .param/.bin unavailable offline)",
                       "l2": "per-frame working set (2 fp16 activation canvases, >= 0.5 GB) exceeds the 126 MB L2; "
                             f"{B} distinct frames cycled",
                       "parallelism": f"segments x{world}, no collective"},
            "output_mpixel_per_s": fps * px * SCALE * SCALE / 1e6,
            "frame_tflops_algorithmic": frame_tflops,
            "frame_frac_of_peak": frame_tflops / peak,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": B * in_bytes,
                    "d2h_bytes_per_step": B * out_bytes, "api": "reve_submit/reve_wait, pinned host buffers",
                    "host_checksum": checksum},
            "gpu_launches": launches,
            "roofline": {"kernel": (f"conv3x3_chain_kernel ({layers_per_launch:.0f} chained 64->64 3x3 + PReLU layers per launch, tcgen05, "
                                    "rotating TMEM banks, layer-to-layer hand-over through L2 scratch rings)") if chained else
                                   "conv3x3_umma_kernel<64,false,false> (64->64 3x3 + PReLU, tcgen05, rotating TMEM banks)",
                         "bound": "tensor", "achieved": body_tflops, "peak": peak, "unit": "TFLOP/s",
                         "frac": body_tflops / peak,
                         "traffic": (lambda t: None if not t else t["dram_bytes_per_launch"] * frames_per_launch / t["frames_per_launch"])(measured_traffic(chained)),
                         "traffic_unit": "bytes of DRAM read+write per launch (ncu capture under profiles/, see r01_body_traffic*.json); "
                                         "algorithmic activation bytes per launch = 2 x 128 B x canvas pixels (one canvas read, one written)",
                         "peak_source": f"{psrc} bf16_tflops_sustained (kernel timed inside a long step)",
                         "avg_launch_ms": body_ms, "launches_timed": int(pr["timed_body"]),
                         "timed_in": "the timed region of `value` (every launch bracketed by CUDA events)" if prof_in_region
                                     else "a separate short pass (burst clocks)",
                         "kernel_ms_share_of_step": (pr["ms_conv0"] + pr["ms_body"] + pr["ms_tail"]) / ms if prof_in_region else None,
                         "algorithmic_flop_per_launch": flop_per_launch,
                         "frames_per_launch": frames_per_launch, "layers_per_launch": layers_per_launch,
                         "ms_per_frame": {"conv0": pr["ms_conv0"] / frames_timed, "body_x16": pr["ms_body"] / frames_timed,
                                          "tail": pr["ms_tail"] / frames_timed}},
            "clocks": clocks,
        }
        if not args.no_cpu and world == 1:
            v, cores, desc = cpu_sample_fps(args.cpu_budget, 200, 10)
            line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc}
        elif world > 1:
            line["cpu_baseline"] = None
        emit(line)
    up.close()
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
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="frames per step")
    ap.add_argument("--tile", type=int, default=200, help="upstream tile size (0 = whole frame)")
    ap.add_argument("--prepad", type=int, default=10)
    ap.add_argument("--model-dir", default="models")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-prof-in-timed-region", action="store_true",
                    help="time the kernels in a short separate pass instead of inside the timed region")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not (args.impl == "ours" and args.gpus > 1 and world == 1):   # (the re-launching parent passes its children's stdout through)
        quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
