"""Anomaly-map head (aaclip_anomaly_head: 4 bf16 levels in, map + score out) and map extrema alone, streaming from HBM
(inputs rotate over enough copies to exceed L2)."""
import sys, json, torch
sys.path.insert(0, ".")
from aaclip_b200 import ops
P, E, S = 576, 768, 336
T = torch.nn.functional.normalize(torch.randn(E, 2, device="cuda"), dim=0)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
def timeit(fn, n=24):
    for i in range(3): fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B in [int(a) for a in sys.argv[1:]] or [64, 128, 256]:
    copies = max(2, (400 << 20) // (B * 4 * P * E * 2) + 1)
    sets = [[torch.nn.functional.normalize(torch.randn(B, P, E, device="cuda"), dim=-1).to(torch.bfloat16) for _ in range(4)]
            for _ in range(copies)]
    det = torch.randn(B, E, device="cuda")
    ms = timeit(lambda i: ops.anomaly_head(sets[i % copies], T, S, ops.HEAD_TEST_INDUSTRIAL, det=det))
    byts = B * (4 * P * E * 2 + S * S * 4 + E * 4 + 4)
    print(f"head B={B} ({copies} input sets): {ms * 1e3:.1f} us  {byts / ms / 1e6:.0f} GB/s ({byts / ms / 1e6 / peak:.3f} of {peak})  {B / ms * 1e3:.0f} img/s")
    del sets
maps = torch.randn(512, S, S, device="cuda")
ms = timeit(lambda i: ops.map_minmax(maps))
byts = maps.numel() * 4
print(f"map_minmax B=512: {ms * 1e3:.1f} us  {byts / ms / 1e6:.0f} GB/s ({byts / ms / 1e6 / peak:.3f} of {peak})")
