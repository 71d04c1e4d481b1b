"""Extract per-launch DRAM traffic of our kernels from an ncu --set full report into profiles/ncu_traffic.json."""
import csv, io, json, os, subprocess, sys

rep = sys.argv[1]
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "ncu_traffic.json")
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = json.load(open(out)) if os.path.exists(out) else {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1].strip()
    rd = float(r[hdr.index("dram__bytes_read.sum")]) * scale[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(r[hdr.index("dram__bytes_write.sum")]) * scale[units[hdr.index("dram__bytes_write.sum")]]
    res[name] = {"dram_bytes": rd + wr, "read": rd, "write": wr, "us": float(r[hdr.index("gpu__time_duration.sum")]),
                 "report": os.path.basename(rep)}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
