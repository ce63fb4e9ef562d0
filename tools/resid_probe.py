"""Residual GEMMs at the bench shape, plain TMA reduce-add (OUT_F32_RESID) vs the folded-LayerNorm producer epilogue
(OUT_F32_RESID_LN: x load / store + bf16 copy + partial sums); buffers rotate so that x streams from HBM."""
import sys, torch
sys.path.insert(0, ".")
from aaclip_b200 import ops
M = 64 * 577
for name, K in (("out", 1024), ("proj", 4096)):
    N = 1024
    a = [(torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16) for _ in range(2)]
    w = (torch.randn(N, K, device="cuda") * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    xs = [torch.zeros(M, N, device="cuda") for _ in range(3)]
    def plain(i): ops.gemm(a[i % 2], w, bias, ops.ACT_NONE, ops.OUT_F32_RESID, out=xs[i % 3])
    def rln(i): ops.gemm_resid_ln(a[i % 2], w, bias, xs[i % 3])
    for label, fn in (("reduce-add", plain), ("resid_ln", rln)):
        for i in range(3): fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        e0.record()
        for i in range(n): fn(i)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        print(f"{name:5s} {label:11s} {us:7.1f} us  {2 * M * N * K / us / 1e6:6.0f} TFLOP/s")
