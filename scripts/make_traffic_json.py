"""profiles/r2_traffic.json from the `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`
launch record of the dominant kernel on the bench workload (one cfg-4 launch of tc64_var_kernel):

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:tc64_var -s 5 -c 1 --csv --log-file gpurun_out/traffic.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-width-sweep --no-operator-configs --e2e-steps 1
    python scripts/make_traffic_json.py gpurun_out/traffic.csv synthetic_2dt_1e6x64_mlp4x64_tanh [other.csv label ...]
"""
import csv, json, sys


def read(path):
    vals, kernel = {}, None
    for r in csv.reader(open(path)):
        if len(r) >= 15 and r[12] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
            vals[r[12]] = float(r[14].replace(",", "")); kernel = r[4]
    return dict(kernel=kernel[:80], dram_read_bytes=vals["dram__bytes_read.sum"], dram_write_bytes=vals["dram__bytes_write.sum"],
                duration_ms_under_ncu=vals["gpu__time_duration.sum"] * 1e-6,
                dram_bytes_per_launch=vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"], source=path.split("/")[-1])


out = read(sys.argv[1])
out["workload"] = sys.argv[2]
out["algorithmic_bytes_per_launch"] = 24 * 64000000
out["ratio_to_algorithmic"] = out["dram_bytes_per_launch"] / out["algorithmic_bytes_per_launch"]
rest = sys.argv[3:]
out["variants"] = {label: read(path) for path, label in zip(rest[0::2], rest[1::2])}
json.dump(out, open("profiles/r2_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
