"""Reference points for the attention kernel at the bench shape (B=64, 16 heads, L=577, d=64, bf16): torch SDPA with the
cuDNN and flash backends (library kernels, for comparison only - never on the product path)."""
import sys, torch, torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel
sys.path.insert(0, ".")
from aaclip_b200 import ops
B, L, H = 64, 577, 16
qkv = (torch.randn(B * L, 3 * H * 64, device="cuda") * 1.5).to(torch.bfloat16)
q, k, v = [t.reshape(B, L, H, 64).transpose(1, 2).contiguous() for t in qkv.view(B * L, 3, H * 64).unbind(1)]
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
flop = 4 * B * H * L * L * 64
us = timeit(lambda: ops.attention(qkv, B, L, H))
print(f"ours (qkv packed [B*L,3*H*64] in, [B*L,H*64] out): {us:.1f} us  {flop / us / 1e6:.0f} TFLOP/s")
for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION), ("efficient", SDPBackend.EFFICIENT_ATTENTION)):
    try:
        with sdpa_kernel(be):
            us = timeit(lambda: F.scaled_dot_product_attention(q, k, v))
        print(f"torch SDPA {name:9s} ([B,H,L,64] contiguous in): {us:.1f} us  {flop / us / 1e6:.0f} TFLOP/s")
    except Exception as e:
        print(f"torch SDPA {name}: unavailable ({str(e)[:80]})")
try:
    from flash_attn import flash_attn_func
    q2, k2, v2 = [t.reshape(B, L, H, 64) for t in qkv.view(B * L, 3, H * 64).unbind(1)]
    us = timeit(lambda: flash_attn_func(q2, k2, v2))
    print(f"flash_attn 2 ([B,L,H,64] strided in): {us:.1f} us  {flop / us / 1e6:.0f} TFLOP/s")
except Exception as e:
    print("flash_attn: unavailable", str(e)[:80])
