"""Latency of the fused forward at small batches, eager launches vs CUDA-graph replay (AACLIP_GRAPH=0/1; the graph
path needs a non-default stream and stable pointers)."""
import os, sys, torch
sys.path.insert(0, ".")
from aaclip_b200 import synth
from aaclip_b200.engine import Engine
cfg = synth.VIT_L_14_336
eng = Engine(cfg, device=0, max_batch=64, text=False)
eng.load_state_dicts(synth.clip_state_dict(cfg, 0, text=False), synth.image_adapter_state_dict(cfg, 0), None)
T = synth.anchors(cfg, 1).cuda()
side = torch.cuda.Stream()
print("AACLIP_GRAPH =", os.environ.get("AACLIP_GRAPH", "1 (default)"))
for B in (1, 2, 4, 8, 16, 64):
    img = synth.images(B, cfg, seed=B).cuda()
    out = (torch.empty(B, 336, 336, device="cuda"), torch.empty(B, device="cuda"))
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        for _ in range(4): eng.forward_fused(img, T, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        e0.record()
        for _ in range(n): eng.forward_fused(img, T, out=out)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"B={B:3d}: {ms:7.3f} ms/batch  {B / ms * 1e3:8.1f} images/s   launches so far {eng.launch_count}")
