"""ncu --set full report of tools/ncu_target.py -> profiles/r2_ncu_traffic.json (what bench.py's roofline.traffic reads)
and a text summary.  The JSON carries the sha of the kernel sources it was captured on: bench.py reports traffic only
while that sha matches the tree it runs from.
    python tools/ncu_traffic.py gpurun_out/<name>.ncu-rep [profiles/<summary>.txt]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha  # noqa: E402

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
units = rows[1]
ki = hdr.index("Kernel Name")


def scaled(name, r):
    v, u = float(r[hdr.index(name)].replace(",", "")), units[hdr.index(name)]
    if name.startswith("dram__bytes"):
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if name == "gpu__time_duration.sum":
        return v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)   # microseconds
    return v


launches = []
for r in rows[2:]:
    launches.append((r[ki], {w: scaled(w, r) for w in WANT if w in hdr}))

# name the launches of tools/ncu_target.py by their order
gemm_order = ["gemm_qkv", "gemm_qkv", "gemm_out", "gemm_out", "gemm_fc", "gemm_fc", "gemm_proj", "gemm_proj"]
kernels, gi, txt = {}, 0, []
for name, m in launches:
    if "gemm_kernel" in name:
        cls = gemm_order[gi] if gi < len(gemm_order) else "gemm_other"
        gi += 1
    elif "attention_kernel" in name:
        cls = "attention"
    elif "head_stream" in name:
        cls = "head_stream"
    elif "maps_from_dots" in name:
        cls = "head_maps"
    else:
        continue
    e = kernels.setdefault(cls, {"kernel": name[:100], "launches": 0, "us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
    e["launches"] += 1
    e["us"] += m.get("gpu__time_duration.sum", 0.0)
    e["dram_read"] += m.get("dram__bytes_read.sum", 0.0)
    e["dram_write"] += m.get("dram__bytes_write.sum", 0.0)
    txt.append(f"kernel: {name[:110]}   [{cls}]")
    for w in WANT:
        if w in m:
            txt.append(f"  {w:74s} {m[w]:16.3f}")
for e in kernels.values():
    n = e.pop("launches")
    e["us"] /= n
    e["dram_read"] /= n
    e["dram_write"] /= n
    e["traffic"] = e["dram_read"] + e["dram_write"]
try:
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
except Exception:
    commit = None
out = {"kernel_source_sha": kernel_source_sha(), "commit": commit, "batch": 64, "report": os.path.basename(rep),
       "how": "ncu --set full --clock-control none on tools/ncu_target.py (one launch at a time, cold caches): per-launch averages; "
              "dram bytes = dram__bytes_read.sum + dram__bytes_write.sum",
       "kernels": kernels}
if "gemm_fc" in kernels and "gemm_proj" in kernels:
    out["gemm_traffic_avg_bytes"] = (kernels["gemm_fc"]["traffic"] + kernels["gemm_proj"]["traffic"]) / 2
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"), "w"), indent=1)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(f"== {os.path.basename(rep)} (ncu --set full --clock-control none, tools/ncu_target.py, B = 64; "
                                 f"kernel sources {out['kernel_source_sha']}, commit {commit})\n" + "\n".join(txt) + "\n")
print(json.dumps(out["kernels"], indent=1))
