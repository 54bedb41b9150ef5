"""Timeline of chain 0 of the second chained launch (REVE_DEBUG_TRACE=1, REVE_CHAIN=2|4).  GPU box only.
Per layer j of the chain: trace[512 j + 0..499] = clock at MMA steps 200..699; [500..503] courier of group 0:
cycles waiting for the staging buffer / store + read-out / publish, rows; [504..507] loader: cycles retiring /
polling `published` / waiting for a free A slot, steps."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("REVE_DEBUG_TRACE", "1")   # "c<n>": trace chained launch n
from _opts import opts_from_env  # noqa: E402
import reve_b200  # noqa: E402
from reve_b200 import _lib  # noqa: E402

L = int(os.environ.get("REVE_CHAIN", "4"))
model = reve_b200.Model.random(2, 1)
up = reve_b200.Upscaler(model, 1920, 1080, tile=200, prepad=10, ring_depth=2, **opts_from_env())
frame = np.random.default_rng(0).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
for _ in range(3):
    up.upscale(frame)
tr = np.zeros(4096, np.int64)
assert _lib.load().reve_debug_trace(up._h, tr.ctypes.data, 4096) == 0
for j in range(L):
    t = tr[512 * j:512 * j + 484]
    t = t[t > 0]
    d = np.diff(t)
    d = d[(d > 0) & (d < 100000)]     # (entries of an earlier launch with more steps may linger at the end)
    c = tr[512 * j + 500:512 * j + 508]
    line = f"layer {j}: steps {len(t)}"
    if len(d):
        line += f"  clk/step mean {d.mean():.0f} median {np.median(d):.0f} p10 {np.percentile(d, 10):.0f} p90 {np.percentile(d, 90):.0f} max {d.max()}"
    if c[3]:
        line += f" | courier/row: wait_full {c[0] / c[3]:.0f} store {c[1] / c[3]:.0f} publish {c[2] / c[3]:.0f}"
    if c[7]:
        line += f" | loader/step: retire {c[4] / c[7]:.0f} poll {c[5] / c[7]:.0f} a_empty {c[6] / c[7]:.0f}"
    ep = tr[512 * j + 484:512 * j + 489]
    if ep[4]:
        n = ep[4]
        line += (f" | epilogue warp 0/event: wait acc {ep[0] / n:.0f} drain {ep[1] / n:.0f} wait quarter {ep[2] / n:.0f} "
                 f"output {ep[3] / n:.0f}")
    m = tr[512 * j + 492:512 * j + 496]
    if c[7]:
        runs = 3     # the counters accumulate over the three traced launches above
        line += (f" | MMA issuer/step: wait for A {m[0] / c[7] / runs:.0f} wait for slot {m[1] / c[7] / runs:.0f} "
                 f"({100 * m[2] / c[7] / runs:.0f} % of steps waited)")
    print(line)

# per-CTA wall clock (globaltimer, ns): kernel-relative start of the MMA warp, first step, end of the last step
cta = tr[2048:2048 + 148 * 4].reshape(148, 4).copy()
smid = cta[:, 3] >> 32
cta[:, 3] &= 0xFFFFFFFF
nz = cta[:, 3] > 0
cta, smid = cta[nz], smid[nz]
t0 = cta[:, 0].min()
print(f"CTAs traced {len(cta)}: MMA-warp start spread {(cta[:, 0].max() - t0) / 1e3:.1f} us")
for j in range(L):
    c = cta[j::L]
    print(f"layer {j}: first step at {np.median(c[:, 1] - t0) / 1e3:.1f} us (max {(c[:, 1] - t0).max() / 1e3:.1f}), "
          f"last step ends {np.median(c[:, 2] - t0) / 1e3:.1f} us (min {(c[:, 2] - t0).min() / 1e3:.1f} max {(c[:, 2] - t0).max() / 1e3:.1f}), "
          f"steps {c[:, 3].min()}..{c[:, 3].max()}, ns/step {np.median((c[:, 2] - c[:, 1]) / c[:, 3]):.0f}")

print("chain: SMs of its CTAs | steps of layer 0 | end of the last layer (us) | ns/step of each layer")
order = np.argsort(cta[L - 1::L, 2])
if os.environ.get("TRACE_TOP"):
    order = order[-int(os.environ["TRACE_TOP"]):]
for c in order:
    rows = cta[c * L:(c + 1) * L]
    print(f"{c:3d}: {smid[c * L:(c + 1) * L].tolist()} | {rows[0, 3]} | {(rows[-1, 2] - t0) / 1e3:.1f} | "
          f"{[int((r[2] - r[1]) / r[3]) for r in rows]}")
