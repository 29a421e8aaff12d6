"""Compact reading of an .ncu-rep: key metrics (raw page) and the top stalled SASS instructions (source page).
usage: python profiles/ncu_summary.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__cycles_elapsed.max.per_second", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_lsu.sum",
        "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_xu.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
for r in rows[2:]:
    print("# kernel:", r[h.index("Kernel Name")][:100])
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print(f"{k:85s} {u[i]:16s} {r[i]}")
    for i, name in enumerate(h):
        if name.startswith("smsp__pcsamp_warps_issue_stalled") and not name.endswith("not_issued") and r[i] not in ("", "0"):
            print(f"{name:85s} {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the source page repeats its header per captured launch (twice per launch: SASS and the high-level view)
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Address":
        cur = {"h": r, "d": []}
        blocks.append(cur)
    elif cur is not None and len(r) == len(cur["h"]):
        cur["d"].append(r)
seen = set()
for bi, blk in enumerate(blocks):
    h, data = blk["h"], blk["d"]
    iS, iI, isrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    tot = sum(int(r[iS]) for r in data)
    key = (len(data), tot)
    if key in seen or tot < 1000:          # duplicate view of the same launch, or a launch too short to sample
        continue
    seen.add(key)
    stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    print(f"# source block {bi}: total samples", tot, "warp instructions", sum(int(r[iI]) for r in data))
    for i in sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:top_n]:
        r = data[i]
        st = sorted(((h[c][6:], int(r[c])) for c in stall_cols if int(r[c] or 0) > 0), key=lambda kv: -kv[1])[:3]
        print(f"{i:5d} {int(r[iS]):7d} {int(r[iI]):10d}  {r[isrc].strip()[:64]:64s} {st}")
