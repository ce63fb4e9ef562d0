"""Graph-replayed fused forward at small batches (device-resident inputs); AACLIP_PDL=0/1 A/B.
usage: python tools/small_e2e.py [B ...]"""
import os
import sys

import torch

sys.path.insert(0, ".")
from aaclip_b200 import synth  # noqa: E402
from aaclip_b200.engine import Engine  # noqa: E402

cfg = synth.VIT_L_14_336
eng = Engine(cfg, device=0, max_batch=64, text=False)
eng.load_state_dicts(synth.clip_state_dict(cfg, 0, text=False), synth.image_adapter_state_dict(cfg, 0), None)
T = synth.anchors(cfg, 1).cuda()
side = torch.cuda.Stream()
res = []
for B in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16]:
    img = torch.randn(B, 3, 336, 336, device="cuda")
    out = (torch.empty(B, 336, 336, device="cuda"), torch.empty(B, device="cuda"))
    with torch.cuda.stream(side):
        for _ in range(4):
            eng.forward_fused(img, T, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            eng.forward_fused(img, T, out=out)
        e1.record()
    torch.cuda.synchronize()
    res.append(f"B={B}: {e0.elapsed_time(e1) / 30:.3f} ms")
print(f"AACLIP_PDL={os.environ.get('AACLIP_PDL', '0')}  " + "  ".join(res))
