"""Segment-level end-to-end stand-in for BASELINE.json configs[4] (reve-cli 1080p x2 with decode and encode
overlapped).  ffmpeg / x265 are not in the image, so this measures the upscale stage together with the
frame I/O on both sides of it: (a) the reference's own contract, PNG directories in and out (C++ driver,
zlib), and (b) raw rgb24 streaming (what ffmpeg pipes would carry).  Host-bound by design: it shows where
the time goes once the GPU does > 300 frames/s.   python tools/bench_segment.py [n_frames]"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "reve_b200", "host", "reve-upscale")


def main():
    import cv2
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    w, h, s = 1920, 1080, 2
    base = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    rng = np.random.default_rng(0)
    # anime-like content compresses far better than noise: flat regions + edges
    frames = []
    for i in range(4):
        f = np.zeros((h, w, 3), np.uint8)
        f[:, :, 0] = np.linspace(0, 255, w, dtype=np.uint8)[None, :]
        f[:, :, 1] = np.linspace(0, 255, h, dtype=np.uint8)[:, None]
        for _ in range(40):
            x0, y0 = int(rng.integers(0, w - 200)), int(rng.integers(0, h - 200))
            f[y0:y0 + int(rng.integers(20, 200)), x0:x0 + int(rng.integers(20, 200))] = rng.integers(0, 256, 3)
        frames.append(f)
    indir = os.path.join(base, "tmp_frames", "0")
    os.makedirs(indir)
    for i in range(n):
        cv2.imwrite(os.path.join(indir, f"frame{i + 1:08d}.png"), frames[i % 4][:, :, ::-1], [cv2.IMWRITE_PNG_COMPRESSION, 1])
    raw = os.path.join(base, "in.rgb")
    with open(raw, "wb") as fo:
        for i in range(n):
            fo.write(frames[i % 4].tobytes())
    res = {"frames": n, "frame": [w, h], "scale": s, "host_cores": os.cpu_count()}
    env = dict(os.environ, REVE_HOST_TIMING="1")

    def run(cmd, shell=False):
        """(frames/s of the whole process, frames/s after CUDA start-up and context creation)"""
        t0 = time.perf_counter()
        p = subprocess.run(cmd, shell=shell, env=env, stderr=subprocess.PIPE, stdout=subprocess.DEVNULL if not shell else None, text=True)
        total = time.perf_counter() - t0
        assert p.returncode == 0, p.stderr[-2000:]
        ms = re.findall(r"\[timing\].*? ([0-9.]+) ms", p.stderr)
        if "start-up" in p.stderr:                       # PNG mode: one line, time after start-up
            stream = float(ms[-1]) / 1e3
        else:                                             # raw mode: ready / gpu done / written
            stream = (float(ms[-1]) - float(ms[0])) / 1e3
        return n / total, n / stream

    res["png_dirs_fps"], res["png_dirs_fps_after_startup"] = run(
        [EXE, "-i", indir, "-o", os.path.join(base, "out_frames", "0"), "-s", str(s), "-m", "/nonexistent"])
    res["raw_rgb24_files_fps"], res["raw_rgb24_files_fps_after_startup"] = run(
        [EXE, "--raw", f"{w}x{h}", "-i", raw, "-o", os.path.join(base, "out.rgb"), "-s", str(s), "-m", "/nonexistent"])
    res["raw_rgb24_pipes_fps"], res["raw_rgb24_pipes_fps_after_startup"] = run(
        f"cat {raw} | {EXE} --raw {w}x{h} -i - -o - -s {s} -m /nonexistent | cat > /dev/null", shell=True)
    res["note"] = ("*_fps includes process start, CUDA initialisation and context creation (0.8-3.5 s on a fresh box), "
                   "*_after_startup is the streaming rate; PNG = zlib level 1 on a pool of host threads (decode 1/4, "
                   "encode 3/4 of the cores); raw mode = reader / submit / writer threads around one context; "
                   "no x265/ffmpeg in the image")
    print(json.dumps(res))
    subprocess.call(["rm", "-rf", base])


if __name__ == "__main__":
    main()
