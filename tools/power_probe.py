"""Power / clock of each hot kernel class run back to back for SECS seconds (NVML in-process)."""
import sys, threading, time, torch, pynvml
sys.path.insert(0, ".")
from aaclip_b200 import ops
SECS = float(sys.argv[1]) if len(sys.argv) > 1 else 2.5
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
M, W, L, H = 64 * 577, 1024, 577, 16
def run(name, fn, unit_work, unit):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    rows, stop = [], [False]
    def poll():
        while not stop[0]:
            rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)); time.sleep(0.05)
    th = threading.Thread(target=poll, daemon=True); th.start()
    t0 = time.perf_counter(); marks = []
    while time.perf_counter() - t0 < SECS:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): fn()
        e1.record(); e1.synchronize()
        marks.append((time.perf_counter() - t0, e0.elapsed_time(e1) / 50))
    stop[0] = True; th.join()
    late = [ms for t, ms in marks if t > SECS / 2] or [marks[-1][1]]
    ms = sum(late) / len(late)
    r = rows[len(rows) // 2:] or [(0, 0)]
    print(f"{name:22s} first {marks[0][1] * 1e3:7.1f} us  sustained {ms * 1e3:7.1f} us ({unit_work / ms / 1e9:7.1f} {unit})  "
          f"clk {sorted(x[0] for x in r)[len(r) // 2]:5.0f} MHz  power {sorted(x[1] for x in r)[len(r) // 2]:5.0f} W")
    time.sleep(1.0)
qkv = (torch.randn(M, 3 * W, device="cuda") * 1.5).bfloat16()
run("attention", lambda: ops.attention(qkv, 64, L, H), 4.0 * 64 * H * L * L * 64, "TF/s")
x = torch.randn(M, W, device="cuda"); g = torch.ones(W, device="cuda"); b = torch.zeros(W, device="cuda")
run("layernorm", lambda: ops.layernorm(x, g, b), M * W * 6.0 * 1e3, "GB/s")
a = (torch.randn(M, W, device="cuda") * 0.5).bfloat16(); w = (torch.randn(4 * W, W, device="cuda") * 0.03).bfloat16()
bias = torch.randn(4 * W, device="cuda"); out = torch.empty(M, 4 * W, device="cuda", dtype=torch.bfloat16)
run("gemm fc (GELU)", lambda: ops.gemm(a, w, bias, ops.ACT_GELU_ERF, ops.OUT_BF16, out=out), 2.0 * M * 4 * W * W, "TF/s")
