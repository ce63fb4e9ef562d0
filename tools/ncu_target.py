"""Small, fixed launch sequence for ncu captures (one launch list + one --set full capture per change), at the bench
shapes (B = 64, M = 64 x 577), inputs rotating so every launch streams from HBM:
  head : 3 x aaclip_anomaly_head (bf16, 4 levels)  -> head_stream_kernel + maps_from_dots_kernel
  gemm : 2 x in_proj (ln_1 folded) | 2 x out_proj (residual + statistics) | 2 x c_fc (ln_2 folded, erf-GELU) |
         2 x c_proj (residual + statistics)          <- tools/ncu_traffic.py relies on this order
  attn : 2 x attention
Usage: python tools/ncu_target.py [head] [gemm] [attn]"""
import sys

import torch

sys.path.insert(0, ".")
from aaclip_b200 import ops  # noqa: E402

what = set(sys.argv[1:]) or {"head", "gemm", "attn"}
B, L, P, E, S = 64, 577, 576, 768, 336
M = B * L
torch.manual_seed(0)
if "head" in what:
    T = torch.nn.functional.normalize(torch.randn(E, 2, device="cuda"), dim=0)
    sets = [[torch.nn.functional.normalize(torch.randn(B, P, E, device="cuda"), dim=-1).to(torch.bfloat16) for _ in range(4)]
            for _ in range(2)]
    det = torch.randn(B, E, device="cuda")
    for i in range(3):
        ops.anomaly_head(sets[i % 2], T, S, ops.HEAD_TEST_INDUSTRIAL, det=det, want_extrema=True)
    torch.cuda.synchronize()
    del sets
if "gemm" in what:
    xs = [torch.randn(M, 1024, device="cuda") for _ in range(3)]
    g, bta = torch.randn(1024, device="cuda") * 0.1 + 1, torch.randn(1024, device="cuda") * 0.1
    xb, part = ops.rowstats_cast(xs[0], 8)
    for N, K, kind in ((3072, 1024, "lnfold"), (1024, 1024, "resid"), (4096, 1024, "lnfold_gelu"), (1024, 4096, "resid")):
        w32 = torch.randn(N, K, device="cuda") * 0.03
        bias = torch.randn(N, device="cuda")
        if kind.startswith("lnfold"):
            wf, cs, bf = ops.fold_ln_weight(w32, bias, g, bta)
            for i in range(2):
                ops.gemm_lnfold(xb, wf, bf, cs, part, act=ops.ACT_GELU_ERF if kind.endswith("gelu") else ops.ACT_NONE)
        else:
            a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
            w = w32.to(torch.bfloat16)
            for i in range(2):
                ops.gemm_resid_ln(a, w, bias, xs[i % 3])
            del a, w
        torch.cuda.synchronize()
        del w32
    del xs
if "attn" in what:
    qkvs = [(torch.randn(M, 3072, device="cuda") * 1.5).to(torch.bfloat16) for _ in range(2)]
    for i in range(2):
        ops.attention(qkvs[i], B, L, 16)
    torch.cuda.synchronize()
print("ok")
