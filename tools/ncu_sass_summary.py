"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` output: opcode histogram by executed
warp instructions and the hottest stall sites.   usage: python tools/ncu_sass_summary.py file.csv [top_n]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr_i]
ix = {n: i for i, n in enumerate(h)}
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
ops = collections.Counter(); samples = collections.Counter()
data = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(h): continue
    src = r[ix["Source"]]
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src[:20]
    base = ".".join(op.split(".")[:2]) if op.startswith(("MUFU", "UTC", "LDTM", "STTM", "SYNCS", "F2FP")) else op.split(".")[0]
    ie = int(float(r[ix["Instructions Executed"]] or 0)); s = int(float(r[ix["# Samples"]] or 0))
    ops[base] += ie; samples[base] += s
    data.append((r[ix["Address"]], src, ie, s, r))
tot = sum(ops.values()); ts = sum(samples.values())
print(f"total warp instructions {tot}, samples {ts}")
for k, v in ops.most_common(top):
    print(f"  {k:22s} {v:12d} {100.0 * v / tot:6.2f}%   samples {100.0 * samples[k] / max(ts, 1):6.2f}%")
print("hottest instructions by samples:")
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
for a, src, ie, s, r in sorted(data, key=lambda d: -d[3])[:top]:
    st = sorted(((int(float(r[ix[n]] or 0)), n[6:]) for n in stall_cols), reverse=True)[:2]
    print(f"  {s:7d} {100.0 * s / max(ts, 1):5.2f}%  exec {ie:10d}  {src[:70]:70s} {st}")
