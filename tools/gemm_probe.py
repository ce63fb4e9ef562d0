"""The encoder GEMMs alone at the bench shape (M = 64*577 rows): CUDA-event timing (for ncu captures).
usage: python tools/gemm_probe.py [which ...]   which in {qkv,out,fc,proj}; default all"""
import sys, torch
sys.path.insert(0, ".")
from aaclip_b200 import ops
M = 64 * 577
SHAPES = {"qkv": (3072, 1024, ops.ACT_NONE, ops.OUT_BF16), "out": (1024, 1024, ops.ACT_NONE, ops.OUT_F32_RESID),
          "fc": (4096, 1024, ops.ACT_GELU_ERF, ops.OUT_BF16), "proj": (1024, 4096, ops.ACT_NONE, ops.OUT_F32_RESID)}
which = [a for a in sys.argv[1:] if a in SHAPES] or list(SHAPES)
for name in which:
    N, K, act, om = SHAPES[name]
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if om == ops.OUT_BF16 else torch.float32)
    for _ in range(3):
        ops.gemm(a, w, bias, act, om, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        ops.gemm(a, w, bias, act, om, out=out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"gemm {name} M={M} N={N} K={K}: {us:.1f} us  {2 * M * N * K / us / 1e6:.0f} TFLOP/s")
