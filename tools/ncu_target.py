"""Small, fixed launch sequence for ncu captures (one launch list + one --set full capture per change):
  3 x head_stream (bf16, 4 levels, B=64)   3 x out_proj residual GEMM   2 x c_proj residual GEMM   2 x attention (B=64)
Inputs rotate so every launch streams from HBM.  Usage: python tools/ncu_target.py [head] [gemm] [attn]"""
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
    for K, reps in ((1024, 3), (4096, 2)):
        a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
        w = (torch.randn(1024, K, device="cuda") * 0.03).to(torch.bfloat16)
        bias = torch.randn(1024, device="cuda")
        xs = [torch.randn(M, 1024, device="cuda") for _ in range(3)]
        for i in range(reps):
            ops.gemm_resid_ln(a, w, bias, xs[i % 3])
        torch.cuda.synchronize()
        del a, w, xs
if "attn" in what:
    qkvs = [(torch.randn(M, 3072, device="cuda") * 1.5).to(torch.bfloat16) for _ in range(2)]
    for i in range(2):
        ops.attention(qkvs[i], B, L, 16)
    torch.cuda.synchronize()
print("ok")
