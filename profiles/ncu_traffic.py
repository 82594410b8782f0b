"""profiles/ncu_traffic.json from an `ncu --set full` report: DRAM bytes per particle and duration of one launch of each hot kernel.
usage: python profiles/ncu_traffic.py <report.ncu-rep> <particles> [<report2> ...] > profiles/ncu_traffic.json"""
import csv, io, json, subprocess, sys
reps = [a for a in sys.argv[1:] if a.endswith(".ncu-rep")]
npart = float([a for a in sys.argv[1:] if not a.endswith(".ncu-rep")][0])
out = {"source": "ncu --set full --clock-control none, bench.py --particles %d (%s), one launch per kernel" % (npart, ", ".join(reps)), "kernels": {}}
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    def val(r, name):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[i], 1.0)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        if name in out["kernels"]:
            continue
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        out["kernels"][name] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "particles": int(npart),
                                "bytes_per_particle": (rd + wr) / npart, "duration_us": val(r, "gpu__time_duration.sum")}
print(json.dumps(out, indent=1))
