"""Summarise an ncu report: per-kernel headline metrics + SASS opcode histogram (instructions per particle).
usage: python profiles/ncu_summarize.py <report.ncu-rep> <kernel-regex> <particles>"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, kre, npart = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]
want += [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in h]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:78s}", [r[i] for r in rows[1:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("Instructions Executed")].isdigit()]
isrc, ins, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[ins]) for r in data)
print("warp-instructions", tot, " per particle", tot / (npart / 32))
c, s = Counter(), Counter()
for r in data:
    t = r[isrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += int(r[ins])
    s[op] += int(r[ismp] or 0)
for op, v in c.most_common(24):
    print(f"  {op:10s} {v / (npart / 32):8.1f} /particle   samples {s[op]}")
print("hottest SASS by samples:")
for r in sorted(data, key=lambda r: -int(r[ismp] or 0))[:18]:
    print("  ", r[ismp], r[ins], r[isrc][:100])
