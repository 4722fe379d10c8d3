"""profiles/r1_traffic.json from an `ncu --set full` capture of the dominant kernel on the bench workload."""
import csv, json, subprocess, sys
rep, workload = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
names, units, vals = rows[0], rows[1], rows[2]
def get(n):
    i = names.index(n); v = float(vals[i].replace(",", "")); u = units[i]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
out = dict(workload=workload, kernel=vals[names.index("Kernel Name")][:80], dram_read_bytes=get("dram__bytes_read.sum"),
           dram_write_bytes=get("dram__bytes_write.sum"), source=rep.split("/")[-1],
           duration_ms_under_ncu=float(vals[names.index("gpu__time_duration.sum")]) * {"us": 1e-3, "ms": 1, "s": 1e3, "ns": 1e-6}[units[names.index("gpu__time_duration.sum")]])
out["dram_bytes_per_launch"] = out["dram_read_bytes"] + out["dram_write_bytes"]
json.dump(out, open("profiles/r1_traffic.json", "w"), indent=1)
print(out)
