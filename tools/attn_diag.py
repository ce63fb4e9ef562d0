import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from aaclip_b200 import ops
from test_kernels_gpu import _attention_ref, _rand_bf16
B, L, H = 2, 577, 16
qkv = _rand_bf16(B * L, 3 * H * 64, scale=1.5, seed=30)
ref = _attention_ref(qkv, B, L, H, False).view(B, L, H, 64)
for rep in range(3):
    out = ops.attention(qkv, B, L, H, False).float().view(B, L, H, 64)
    err = (out - ref).abs().amax(-1)          # [B, L, H]
    bad = (err > 0.05)
    print(f"rep {rep}: max err {err.max():.3f}, bad rows {int(bad.sum())} of {bad.numel()}")
    idx = bad.nonzero()
    import collections
    c = collections.Counter((int(b), int(h), int(l) // 128, (int(l) % 128) // 32) for b, l, h in idx.tolist())
    for k, v in sorted(c.items())[:40]:
        print("  b,h,qtile,warp =", k, "rows", v)
