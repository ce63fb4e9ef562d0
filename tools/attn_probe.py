"""Attention kernel alone at the bench shape (B=64, L=577, 16 heads): CUDA-event timing (for ncu captures)."""
import sys, torch
sys.path.insert(0, ".")
from aaclip_b200 import ops
B, L, H = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 577, 16
qkv = (torch.randn(B * L, 3 * H * 64, device="cuda") * 1.5).to(torch.bfloat16)
for _ in range(3):
    out = ops.attention(qkv, B, L, H)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    out = ops.attention(qkv, B, L, H)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / n * 1e3
flop = 4 * B * H * L * L * 64
print(f"attention B={B}: {us:.1f} us  {flop / us / 1e6:.0f} TFLOP/s")
