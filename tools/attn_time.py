"""Attention kernel alone at the bench shape (B = 64, L = 577, 16 heads), inputs rotating over 3 QKV buffers
(3 x 227 MB > L2), CUDA events over 30 launches; and the text shape (240 x 77, 12 heads, causal)."""
import sys

import torch

sys.path.insert(0, ".")
from aaclip_b200 import ops  # noqa: E402

for B, L, H, causal in ((64, 577, 16, False), (240, 77, 12, True)):
    qkvs = [(torch.randn(B * L, 3 * H * 64, device="cuda") * 1.5).to(torch.bfloat16) for _ in range(3)]
    for i in range(5):
        ops.attention(qkvs[i % 3], B, L, H, causal)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for i in range(n):
        ops.attention(qkvs[i % 3], B, L, H, causal)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    flop = 4.0 * B * H * L * L * 64
    print(f"attention B={B} L={L} heads={H} causal={causal}: {us:.1f} us  {flop / us / 1e6:.0f} TF/s")
