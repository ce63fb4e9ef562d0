"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: count, total, average, share.
usage: python tools/ncu_launch_summary.py launches.csv ["header line"]"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).strip()
    if "gemm_kernel" in r[ki] or "attention_kernel" in r[ki]:
        name = re.sub(r"\(CUtensorMap.*", "", r[ki]).replace("(int)", "").replace("(bool)", "").strip()
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui], 1e-6)
    tot[name] += v; cnt[name] += 1
allms = sum(tot.values())
if len(sys.argv) > 2: print(sys.argv[2])
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:70]:70s} n={cnt[k]:4d} total={v:9.3f} ms avg={v / cnt[k] * 1e3:9.1f} us share={100 * v / allms:5.1f}%")
print(f"{'TOTAL':70s} n={sum(cnt.values()):4d} total={allms:9.3f} ms")
