"""Summarise `ncu --page source --csv` output: samples by opcode and the hottest SASS instructions."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
samp = collections.Counter(); cnt = collections.Counter(); execs = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_stall = collections.Counter()
top = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[idx["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:3]) if op.startswith(("LD", "ST", "RED", "ATOM")) else "")
    n = int(r[idx["# Samples"]] or 0)
    samp[op] += n; cnt[op] += 1; execs[op] += int(r[idx["Instructions Executed"]] or 0)
    for c in stall_cols:
        by_stall[c] += int(r[idx[c]] or 0)
    top.append((n, src, {c: int(r[idx[c]] or 0) for c in stall_cols if int(r[idx[c]] or 0) > 0}))
tot = sum(samp.values())
print("total samples", tot)
print("by stall:", [(k, v) for k, v in by_stall.most_common(8)])
print("%-22s %8s %6s %12s" % ("opcode", "samples", "static", "executed"))
for op, n in samp.most_common(18):
    print("%-22s %8d %6d %12d" % (op, n, cnt[op], execs[op]))
print("total warp-instr executed", sum(execs.values()))
top.sort(key=lambda t: -t[0])
for n, src, st in top[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(n, src[:70], sorted(st.items(), key=lambda kv: -kv[1])[:3])
