"""Per-kernel histogram of the SASS opcodes that prove a Blackwell-native kernel (B200_PROFILING.md: tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP; legacy HMMA must be absent) from the shipped
library.  CPU box:  python tools/sass_histogram.py > profiles/sass_opcodes_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "reve_b200", "libreve_cuda.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "UTMACMDFLUSH", "SYNCS", "ELECT",
        "HMMA", "HGMMA", "FENCE.VIEW.ASYNC", "ERRBAR", "MEMBAR", "STG", "LDG", "STS", "LDS", "HFMA2", "HMNMX2", "FFMA", "F2FP", "PRMT", "REDUX"]
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} ({os.path.getsize(lib)} bytes), opcode counts per kernel (static instructions)")
blocks = re.split(r"\n\s*Function : ", sass)[1:]
for blk, name in zip(blocks, names):
    ops = collections.Counter()
    for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", blk):
        ops[m.group(1)] += 1
    total = sum(ops.values())
    short = re.sub(r"\(.*", "", name.replace("reve::(anonymous namespace)::", ""))
    print(f"\n{short}: {total} instructions")
    agg = collections.Counter()
    for op, n in ops.items():
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                agg[k] += n
        if ".2CTA" in op and op.startswith("UTCHMMA"):
            agg["UTCHMMA.2CTA"] += n
    print("  " + "  ".join(f"{k}={agg[k]}" for k in KEYS if agg[k]))
    assert agg["HMMA"] == 0 and agg["HGMMA"] == 0, "legacy tensor-core instructions in " + short
