"""A/B of the residual + LayerNorm-statistics GEMM epilogue variants (AACLIP_RLN_VARIANT, gemm_sm100.cuh) at the bench
shapes: out_proj (K = 1024) and c_proj (K = 4096), M = 64 * 577, x rotating over 3 buffers so x_old comes from HBM.
    python tools/rln_probe.py            # runs every variant in its own process
    python tools/rln_probe.py one        # the variant selected by the environment"""
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] == "one":
    import torch
    sys.path.insert(0, ".")
    from aaclip_b200 import ops
    M, N = 64 * 577, 1024
    res = []
    for name, K in (("out_proj", 1024), ("c_proj", 4096)):
        a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") * 0.03).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        xs = [torch.randn(M, N, device="cuda") for _ in range(3)]
        ref = xs[0].clone() + (a.float() @ w.float().t()) + bias
        xb, part = ops.gemm_resid_ln(a, w, bias, xs[0])
        torch.cuda.synchronize()
        err = (xs[0] - ref).abs().max().item()
        errb = (xb.float() - ref).abs().max().item()
        perr = (part[:, :, 0].sum(1) - ref.sum(1)).abs().max().item()
        for i in range(6):
            ops.gemm_resid_ln(a, w, bias, xs[i % 3])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 60
        e0.record()
        for i in range(n):
            ops.gemm_resid_ln(a, w, bias, xs[i % 3])
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        res.append(f"{name} {us:7.1f} us  {2.0 * M * N * K / us / 1e6:6.0f} TF/s  err {err:.1e}/{errb:.1e}/{perr:.1e}")
    print(f"variant {os.environ.get('AACLIP_RLN_VARIANT', '0')}: " + "   ".join(res), flush=True)
else:
    for v in (sys.argv[1:] or ["1", "0", "3", "2", "4"]):
        env = dict(os.environ, AACLIP_RLN_VARIANT=v)
        r = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True, timeout=120)
        print(r.stdout.strip() or ("variant " + v + " FAILED: " + r.stderr.strip()[-400:]), flush=True)
