"""Small head_stream invocations for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
from aaclip_b200 import ops
import aaclip_oracle as orc

gen = torch.Generator().manual_seed(0)
for dtype in (torch.float32, torch.bfloat16):
    for B, G, S, nl in ((3, 24, 336, 4), (2, 9, 63, 2)):
        feats = [torch.nn.functional.normalize(torch.randn(B, G * G, 768, generator=gen), dim=-1).to(dtype) for _ in range(nl)]
        T = torch.nn.functional.normalize(torch.randn(768, 2, generator=gen), dim=0)
        det = torch.randn(B, 768, generator=gen)
        maps, scores, ext = ops.anomaly_head([f.cuda() for f in feats], T.cuda(), S, ops.HEAD_TEST_INDUSTRIAL, det=det.cuda(),
                                             want_extrema=True)
        torch.cuda.synchronize()
        ref, sref = orc.predict([f.float() for f in feats], det, T, S, "Industrial")
        print(dtype, B, G, S, nl, "err", (maps.cpu() - ref).abs().max().item(), (scores.cpu() - sref).abs().max().item(), flush=True)
