"""Per-source-line instruction counts from an ncu report (needs -lineinfo + --import-source on).
usage: python profiles/ncu_lines.py <report.ncu-rep> <kernel-regex> <particles> [top]"""
import csv, io, subprocess, sys
rep, kre, npart = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname, hdr, out = "?", None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1] or "?"
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        ins = r[hdr.index("Instructions Executed")]
        smp = r[hdr.index("# Samples")]
        if ins.isdigit():
            out.append((int(ins), int(smp or 0), fname, int(r[0]), r[1].strip()))
tot = sum(o[0] for o in out)
print(f"total warp-instructions {tot}  = {tot / (npart / 32):.1f} thread-instructions per particle")
byfile = {}
for o in out:
    byfile[o[2]] = byfile.get(o[2], 0) + o[0]
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1]):
    print(f"  {f:40s} {v / (npart / 32):8.1f} /particle")
for o in sorted(out, key=lambda o: -o[0])[:top]:
    print(f"{o[0] / (npart / 32):8.1f} /particle  samples {o[1]:6d}  {o[2]}:{o[3]}  {o[4][:110]}")
