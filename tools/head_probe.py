"""Anomaly-map head (aaclip_anomaly_head -> head_stream_kernel) streaming from HBM: bf16 and fp32 tokens, the 4-level
call (test.py:89-93 in one launch) and the 1-level call (one calculate_similarity_map), plus the fused engine path's
maps_from_dots tail.  Inputs rotate over enough copies to exceed the 126 MB L2.
    python tools/head_probe.py [B ...]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from aaclip_b200 import ops  # noqa: E402

P, E, S = 576, 768, 336
T = torch.nn.functional.normalize(torch.randn(E, 2, device="cuda"), dim=0)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]


def timeit(fn, n=24):
    """Device time per call: the n calls are captured into ONE CUDA graph (the entry allocates nothing and never
    synchronises, so it is capturable) and the graph is replayed - no Python / launch overhead between the kernels."""
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(n):
                fn(i)
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        g.replay()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * n)


out = {}
for B in [int(a) for a in sys.argv[1:] if a.isdigit()] or [64, 128, 256]:
    for dtype, es in ((torch.bfloat16, 2), (torch.float32, 4)):
        for nl in (4, 1):
            copies = max(2, (400 << 20) // (B * nl * P * E * es) + 1)
            sets = [[torch.nn.functional.normalize(torch.randn(B, P, E, device="cuda"), dim=-1).to(dtype) for _ in range(nl)]
                    for _ in range(copies)]
            det = torch.randn(B, E, device="cuda")
            ms = timeit(lambda i: ops.anomaly_head(sets[i % copies], T, S, ops.HEAD_TEST_INDUSTRIAL, det=det, want_extrema=True))
            byts = B * (nl * P * E * es + S * S * 4 + E * 4 + 4 + 8)
            name = f"head B={B} {str(dtype).split('.')[-1]} levels={nl}"
            print(f"{name} ({copies} input sets): {ms * 1e3:.1f} us  {byts / ms / 1e6:.0f} GB/s "
                  f"({byts / ms / 1e6 / peak:.3f} of {peak})  {B / ms * 1e3:.0f} img/s", flush=True)
            out[name] = {"us": ms * 1e3, "GBps": byts / ms / 1e6, "frac": byts / ms / 1e6 / peak}
            del sets
if len(sys.argv) > 1 and sys.argv[-1] == "json":
    print(json.dumps(out))
