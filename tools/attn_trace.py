"""Per-tile clock64 trace of one attention CTA (diagnostics)."""
import sys, torch
sys.path.insert(0, ".")
from aaclip_b200 import _lib
from aaclip_b200._lib import check, ptr, cur_stream
B, L, H = 64, 577, 16
qkv = (torch.randn(B * L, 3 * H * 64, device="cuda") * 1.5).to(torch.bfloat16)
out = torch.empty(B * L, H * 64, device="cuda", dtype=torch.bfloat16)
lib = _lib.load()
for cta in (100, 101, 250):
    tr = torch.zeros(24 * 32, dtype=torch.int64, device="cuda")
    for _ in range(2):
        check(lib.aaclip_attention_trace(ptr(qkv), ptr(out), B, L, H, 0, ptr(tr), cta, cur_stream()))
    torch.cuda.synchronize()
    t = tr.cpu().view(24, 32)[:, :19]
    t0 = int(t[0, 0])
    print(f"--- CTA {cta}: softmax warp0 stamps (cycles since first), per tile")
    names = ["tile top", "masked+max", "exp pass", "checked", "S next ld", "arrived", "ld landed", "-"]
    for s_ in range(7):
        print(f"{names[s_]:11s}", " ".join(f"{int(x) - t0:7d}" for x in t[s_]))
    print("MMA thread:")
    names = ["pre p_full", "p_full ok", "v_full ok", "PV issued", "S issued"]
    for s_ in range(5):
        print(f"{names[s_]:11s}", " ".join(f"{int(x) - t0:7d}" for x in t[16 + s_]))
