"""Does halving the batch make the fp32 residual stream L2 resident?  Times one layer's worth of kernels
(LN -> qkv GEMM -> attention -> out GEMM(+=x) -> LN -> fc GEMM -> proj GEMM(+=x)) at B=64 in one go vs two halves of 32."""
import sys, torch
sys.path.insert(0, ".")
from aaclip_b200 import ops
L, W, FF, H = 577, 1024, 4096, 16
def mk(B):
    M = B * L
    d = dict(M=M, B=B)
    d["x"] = torch.randn(M, W, device="cuda")
    d["w_qkv"] = (torch.randn(3 * W, W, device="cuda") * 0.03).bfloat16(); d["b_qkv"] = torch.randn(3 * W, device="cuda")
    d["w_out"] = (torch.randn(W, W, device="cuda") * 0.03).bfloat16(); d["b_out"] = torch.randn(W, device="cuda")
    d["w_fc"] = (torch.randn(FF, W, device="cuda") * 0.03).bfloat16(); d["b_fc"] = torch.randn(FF, device="cuda")
    d["w_pr"] = (torch.randn(W, FF, device="cuda") * 0.03).bfloat16(); d["b_pr"] = torch.randn(W, device="cuda")
    d["g"] = torch.ones(W, device="cuda"); d["be"] = torch.zeros(W, device="cuda")
    d["qkv"] = torch.empty(M, 3 * W, device="cuda", dtype=torch.bfloat16)
    d["h"] = torch.empty(M, FF, device="cuda", dtype=torch.bfloat16)
    return d
def layer(d):
    xn, _ = ops.layernorm(d["x"], d["g"], d["be"])
    ops.gemm(xn, d["w_qkv"], d["b_qkv"], ops.ACT_NONE, ops.OUT_BF16, out=d["qkv"])
    att = ops.attention(d["qkv"], d["B"], L, H)
    ops.gemm(att, d["w_out"], d["b_out"], ops.ACT_NONE, ops.OUT_F32_RESID, out=d["x"])
    xn, _ = ops.layernorm(d["x"], d["g"], d["be"])
    ops.gemm(xn, d["w_fc"], d["b_fc"], ops.ACT_GELU_ERF, ops.OUT_BF16, out=d["h"])
    ops.gemm(d["h"], d["w_pr"], d["b_pr"], ops.ACT_NONE, ops.OUT_F32_RESID, out=d["x"])
def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
full = mk(64)
h1, h2 = mk(32), mk(32)
def six_full():
    for _ in range(6): layer(full)
def six_halves():   # same work: 6 layers on each half, half by half (x of one half stays hot)
    for _ in range(6): layer(h1)
    for _ in range(6): layer(h2)
print(f"6 layers, B=64 in one go : {timeit(six_full):8.3f} ms")
print(f"6 layers, 2 x B=32       : {timeit(six_halves):8.3f} ms")
# the same under the sustained power cap (2.5 s each)
import time
def sustained(fn, secs=2.5):
    t0 = time.perf_counter(); marks = []
    while time.perf_counter() - t0 < secs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4): fn()
        e1.record(); e1.synchronize()
        marks.append((time.perf_counter() - t0, e0.elapsed_time(e1) / 4))
    late = [ms for t, ms in marks if t > secs / 2]
    return sum(late) / len(late)
print(f"sustained: B=64 in one go {sustained(six_full):8.3f} ms   2 x B=32 {sustained(six_halves):8.3f} ms")
