"""Attribute the warp-stall samples of an `ncu --page source --csv` export to CUDA source lines.

usage: ncu_line_summary.py <src.csv> <nvdisasm --print-line-info output> <mangled kernel name substring> [top]
The SASS addresses of the export are matched to the disassembly by offset from the first instruction."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
recs = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    recs.append((int(r[idx["Address"]], 16), r[idx["Source"]].strip(), int(r[idx["# Samples"]] or 0),
                 int(r[idx["Instructions Executed"]] or 0), {c: int(r[idx[c]] or 0) for c in stall_cols}))
base = recs[0][0]
# disassembly: offset -> line
line_of, cur, on = {}, None, False
for ln in open(sys.argv[2]):
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        on = sys.argv[3] in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File ".*?", line (\d+)', ln)
    if m:
        cur = int(m.group(1)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
by_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for addr, src, n, ex, st in recs:
    l = line_of.get(addr - base)
    by_line[l][0] += n; by_line[l][1] += ex
    for k, v in st.items():
        by_line[l][2][k] += v
tot = sum(v[0] for v in by_line.values())
src_lines = open(sys.argv[5] if len(sys.argv) > 5 else "varnet_b200/csrc/vn_tc64.cu").read().split("\n")
print("total samples", tot)
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
for l, (n, ex, st) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src_lines[l - 1].strip()[:90] if l else "?"
    print("%5s %6d %5.1f%% ex=%9d %-40s | %s" % (l, n, 100.0 * n / tot, ex, " ".join("%s:%d" % (k[6:], v) for k, v in st.most_common(3)), text))
