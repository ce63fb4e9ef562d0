"""LayerNorm alone at the bench shape (36928 x 1024 fp32 -> bf16) and in its real context (LN -> qkv GEMM), CUDA events."""
import sys, torch
sys.path.insert(0, ".")
from aaclip_b200 import ops
M, W = 64 * 577, 1024
x = torch.randn(M, W, device="cuda"); g = torch.ones(W, device="cuda"); b = torch.zeros(W, device="cuda")
w = (torch.randn(3 * W, W, device="cuda") * 0.03).bfloat16(); bias = torch.randn(3 * W, device="cuda")
qkv = torch.empty(M, 3 * W, device="cuda", dtype=torch.bfloat16)
big = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
def t(fn, n=20, flush=False):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(n):
        if flush: big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3
us = t(lambda: ops.layernorm(x, g, b), flush=True)
print(f"layernorm alone (L2 flushed): {us:.1f} us  {(M * W * 6) / us / 1e3:.0f} GB/s")
def pair():
    xn, _ = ops.layernorm(x, g, b)
    ops.gemm(xn, w, bias, ops.ACT_NONE, ops.OUT_BF16, out=qkv)
print(f"layernorm + qkv GEMM: {t(pair, flush=True):.1f} us")
