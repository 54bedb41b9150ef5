"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + per-role stall sampling.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel_index]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime", "dram__bytes_read.sum [", "dram__bytes_write.sum [",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.avg [", "sm__cycles_active.avg [", "smsp__inst_executed.sum [",
        "lts__t_sectors_srcunit_tex_op_read.sum [", "lts__t_sectors_srcunit_tex_op_write.sum [",
        "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "sm__throughput.avg.pct",
        "lts__t_sectors_op_read.sum [", "lts__t_sectors_op_write.sum [", "lts__t_sector_op_read_hit_rate",
        "lts__t_sector_op_write_hit_rate", "sm__inst_executed_pipe_uniform", "launch__shared_mem_per_block_dynamic"]
r = rows[2 + kidx]
for h, u, v in zip(hdr, units, r):
    label = f"{h} [{u}]"
    if any(w in label for w in want):
        print(f"{label} = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
out, k = [], -1
for r in rows:
    if r and r[0] == "Kernel Name":
        k += 1
        continue
    if r and r[0] == "Address":
        shdr = r
        continue
    if k == kidx and len(r) > 6:
        out.append(r)
ix = {h: i for i, h in enumerate(shdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in out)
print(f"\ntotal samples {tot}, {len(out)} SASS instructions")
stall_cols = [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
print("top instructions by samples:")
for r in sorted(out, key=lambda r: -int(r[ix["# Samples"]] or 0))[:25]:
    n = int(r[ix["# Samples"]] or 0)
    st = {h[6:]: int(r[ix[h]] or 0) for h in stall_cols}
    st = {a: b for a, b in st.items() if b > n * 0.15}
    print(f"  {n:6d} {r[ix['Address']][-5:]} {r[ix['Source']][:70]:70s} {st}")
# role split: find marker instructions
idx_first = {}
for i, r in enumerate(out):
    s = r[ix["Source"]]
    for key in ("UTMALDG", "UTCHMMA", "LDTM"):
        if key in s and key not in idx_first:
            idx_first[key] = i
print("\nfirst marker index:", idx_first)
