"""Per-kernel-class times of the fused forward at small batches (eager launches with CUDA events around every kernel,
aaclip_profile_*), next to the graph-replayed step time.  usage: python tools/small_batch_probe.py [B ...]"""
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
for B in [int(a) for a in sys.argv[1:]] or [1, 8]:
    img = torch.randn(B, 3, 336, 336, device="cuda")
    out = (torch.empty(B, 336, 336, device="cuda"), torch.empty(B, device="cuda"))
    with torch.cuda.stream(side):
        for _ in range(3):
            eng.forward_fused(img, T, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.forward_fused(img, T, out=out)
        e1.record()
    torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / 20
    eng.profile(True)
    for _ in range(2):
        eng.forward_fused(img, T)
    eng.profile_read()
    for _ in range(3):
        eng.forward_fused(img, T)
    prof = eng.profile_read()
    eng.profile(False)
    tot = sum(v[0] for v in prof.values()) / 3
    print(f"B={B}: graph replay {graph_ms:.3f} ms/step; eager profiled span {eng.profile_span_ms / 3:.3f} ms, sum of kernels {tot:.3f} ms")
    for k, (ms, n) in prof.items():
        if n:
            print(f"   {k:14s} {ms / 3 * 1e3:8.1f} us/step  x{n // 3:3d}  = {ms / n * 1e3:6.1f} us each")
