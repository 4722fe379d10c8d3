"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.
usage: python scripts/make_profile_summary.py <tag> <launches.csv|-> <report.ncu-rep|-> [bench.json]"""
import collections
import csv
import json
import os
import subprocess
import sys

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
bench = sys.argv[4] if len(sys.argv) > 4 else None
out = ["# ncu summary %s" % tag, ""]
if launches != "-":
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[i_val].replace(",", ""))
        v = v / 1e3 if r[i_unit] == "ns" else v * 1e3 if r[i_unit] == "ms" else v
        a = agg.setdefault(r[i_name][:90], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
            "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.4f |" % (n, a[0], a[1], a[1] / tot))
    out.append("")
if rep != "-":
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    names, units = rows[0], rows[1]
    keep = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic"]
    for k, line in enumerate(rows[2:]):
        out += ["## kernel %d: `%s`" % (k, line[names.index("Kernel Name")][:100]), "", "| metric | unit | value |", "|---|---|---:|"]
        for nm in keep:
            if nm in names:
                i = names.index(nm)
                out.append("| %s | %s | %s |" % (nm, units[i], line[i]))
        stalls = [(names[i], float(line[i] or 0)) for i in range(len(names)) if names[i].startswith("smsp__pcsamp_warps_issue_stalled_") and not names[i].endswith("_not_issued")]
        tot = sum(v for _, v in stalls) or 1
        out += ["", "warp-state samples: " + ", ".join("%s %.1f%%" % (n.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot) for n, v in sorted(stalls, key=lambda kv: -kv[1])[:9]), ""]
if bench:
    line = [l for l in open(bench) if l.startswith("{")][-1]
    b = json.loads(line)
    out += ["## bench line of the same build", "", "```json", json.dumps({k: b[k] for k in ("metric", "value", "unit", "ms_per_step", "roofline", "e2e", "clocks", "cpu_baseline") if k in b}, indent=1), "```", ""]
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "%s.md" % tag)
open(path, "w").write("\n".join(out))
print("wrote", path)
