"""Sustained (power-capped) throughput of our GEMM vs torch.matmul (cuBLAS) at the bench shapes: each runs back to
back for SECS seconds, TF/s over the last 2/3 of the window, with nvidia-smi clocks/power sampled alongside.
usage: python tools/sustained_probe.py [secs]"""
import subprocess, sys, threading, time, torch
sys.path.insert(0, ".")
from aaclip_b200 import ops
SECS = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
M = 64 * 577

def sample(stop, rows):
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                         stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        if stop.is_set(): break
        try: rows.append(tuple(float(x) for x in line.split(",")))
        except ValueError: pass
    p.terminate()

def sustained(name, fn, flop):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    stop, rows = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, rows), daemon=True); th.start()
    t0 = time.perf_counter(); n = 0; marks = []
    while time.perf_counter() - t0 < SECS:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): fn()
        e1.record(); e1.synchronize()
        marks.append((time.perf_counter() - t0, e0.elapsed_time(e1) / 50))
    stop.set()
    late = [ms for t, ms in marks if t > SECS / 3]
    first = marks[0][1]
    ms = sum(late) / len(late)
    r = rows[len(rows) // 3:] or [(0, 0)]
    clk = sorted(x[0] for x in r)[len(r) // 2]; pw = sorted(x[1] for x in r)[len(r) // 2]
    print(f"{name:28s} first-50 {flop / first / 1e9:7.0f} TF/s   sustained {flop / ms / 1e9:7.0f} TF/s  ({ms * 1e3:.1f} us)  clk {clk:.0f} MHz  power {pw:.0f} W")

for nm, N, K, act, om in (("fc", 4096, 1024, ops.ACT_GELU_ERF, ops.OUT_BF16), ("proj", 1024, 4096, ops.ACT_NONE, ops.OUT_F32_RESID),
                          ("qkv", 3072, 1024, ops.ACT_NONE, ops.OUT_BF16)):
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if om == ops.OUT_BF16 else torch.float32)
    flop = 2.0 * M * N * K
    sustained(f"ours  {nm} {M}x{N}x{K}", lambda: ops.gemm(a, w, bias, act, om, out=out), flop)
    wt = w.t().contiguous().t()
    o2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    sustained(f"torch {nm} (plain matmul)", lambda: torch.matmul(a, w.t(), out=o2), flop)
    time.sleep(1.0)
