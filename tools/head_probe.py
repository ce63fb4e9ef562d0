"""Anomaly-map head (aaclip_anomaly_head: 4 bf16 levels in, map + score out) and map extrema alone at B=64."""
import sys, json, torch
sys.path.insert(0, ".")
from aaclip_b200 import ops
B, P, E, S = 64, 576, 768, 336
feats = [torch.nn.functional.normalize(torch.randn(B, P, E, device="cuda"), dim=-1).to(torch.bfloat16) for _ in range(4)]
T = torch.nn.functional.normalize(torch.randn(E, 2, device="cuda"), dim=0)
det = torch.randn(B, E, device="cuda")
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = timeit(lambda: ops.anomaly_head(feats, T, S, ops.HEAD_TEST_INDUSTRIAL, det=det))
byts = B * (4 * P * E * 2 + S * S * 4 + E * 4 + 4)
print(f"head B={B}: {ms * 1e3:.1f} us  {byts / ms / 1e6:.0f} GB/s ({byts / ms / 1e6 / peak:.3f} of {peak})")
maps, _ = ops.anomaly_head(feats, T, S, ops.HEAD_TEST_INDUSTRIAL, det=det)
big = maps.repeat(8, 1, 1).contiguous()     # 231 MB > L2
ms = timeit(lambda: ops.map_minmax(big))
byts = big.numel() * 4
print(f"map_minmax B={big.shape[0]}: {ms * 1e3:.1f} us  {byts / ms / 1e6:.0f} GB/s ({byts / ms / 1e6 / peak:.3f} of {peak})")
