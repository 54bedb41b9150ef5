"""Same-box A/B of two builds of libreve_cuda.so (REVE_LIB): 1080p x2 (or AB_SIZE=WxHxS), tile 200, device-resident,
~1.7 s per run, A/B/A/B.  usage: python tools/ab_libs.py reve_b200/libreve_cuda_prev.so reve_b200/libreve_cuda.so"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import json, os, sys
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(sys.path[0], "tools"))
import torch, reve_b200
from _opts import opts_from_env
w, h, scale = (int(v) for v in os.environ.get("AB_SIZE", "1920x1080x2").split("x"))
n = 8
frames = 640 * (1920 * 1080) // (w * h) // 8 * 8
up = reve_b200.Upscaler(reve_b200.Model.random(scale, 1), w, h, tile=200, prepad=10, ring_depth=8, **opts_from_env())
d_in = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
d_out = torch.empty((n, h * scale, w * scale, 3), dtype=torch.uint8, device="cuda")
st = torch.cuda.ExternalStream(up.stream)
for _ in range(10):
    up.upscale_device(d_in.data_ptr(), d_out.data_ptr(), n)
up.sync(); up.set_profiling(True); up.profile(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(frames // n):
    up.upscale_device(d_in.data_ptr(), d_out.data_ptr(), n)
e1.record(st); up.sync()
pr = up.profile(reset=True)
print(json.dumps({"size": [w, h, scale], "lib": os.path.basename(os.environ.get("REVE_LIB", "default")), "fps": round(frames / (e0.elapsed_time(e1) / 1e3), 1),
                  "conv0": round(pr["ms_conv0"] / frames, 4), "body": round(pr["ms_body"] / frames, 4), "tail": round(pr["ms_tail"] / frames, 4)}))
''' % ROOT
for rep in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, REVE_LIB=os.path.abspath(lib))
        print(subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
