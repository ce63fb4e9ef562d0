"""Tile-shape A/B of the residual GEMM (out_proj K=1024, c_proj K=4096) at the bench batch: cta_group 2 = 256x256 CTA
pairs, 1 = 128x256, 3 = 128x128.  x rotates over 3 buffers (x_old from HBM).   python tools/tile_probe.py [B]"""
import sys
import torch
sys.path.insert(0, ".")
from aaclip_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
M, N = B * 577, 1024
for name, K in (("out_proj", 1024), ("c_proj", 4096)):
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    xs = [torch.randn(M, N, device="cuda") for _ in range(3)]
    for cg in (2, 1, 3):
        x0 = xs[0].clone()
        ref = x0 + (a.float() @ w.float().t()) + bias
        ops.gemm_resid_ln(a, w, bias, x0, cta_group=cg)
        err = (x0 - ref).abs().max().item()
        for i in range(6):
            ops.gemm_resid_ln(a, w, bias, xs[i % 3], cta_group=cg)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 60
        e0.record()
        for i in range(n):
            ops.gemm_resid_ln(a, w, bias, xs[i % 3], cta_group=cg)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        print(f"B={B} {name} cta_group={cg}: {us:7.1f} us  {2.0 * M * N * K / us / 1e6:6.0f} TF/s  err {err:.1e}", flush=True)
