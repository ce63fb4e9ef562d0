"""Tile-shape sweep of the four GEMM classes of a visual block at small batches: cta_group 2 (256x256 CTA pairs),
1 (128x256), 3 (128x128) and 0 (the automatic choice), each timed as 24 launches inside one CUDA graph (what the engine
replays).   python tools/tile_sweep.py [B ...]"""
import sys

import torch

sys.path.insert(0, ".")
from aaclip_b200 import ops  # noqa: E402

W, FF = 1024, 4096
side = torch.cuda.Stream()


def graph_us(fn, n=24, reps=5):
    with torch.cuda.stream(side):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(n):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3


for B in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32]:
    M = B * 577
    xb = (torch.randn(M, W, device="cuda") * 0.5).to(torch.bfloat16)
    part = torch.zeros(M, W // 128, 2, device="cuda")
    part[:, :, 0] = xb.float().view(M, W // 128, 128).sum(-1)
    part[:, :, 1] = (xb.float() ** 2).view(M, W // 128, 128).sum(-1)
    h = (torch.randn(M, FF, device="cuda") * 0.5).to(torch.bfloat16)
    x = torch.randn(M, W, device="cuda")
    mk = lambda n, k: (torch.randn(n, k, device="cuda") * 0.03).to(torch.bfloat16)
    w_qkv, w_out, w_fc, w_proj = mk(3 * W, W), mk(W, W), mk(FF, W), mk(W, FF)
    cs_qkv, cs_fc = w_qkv.float().sum(1), w_fc.float().sum(1)
    b_qkv, b_out, b_fc, b_proj = (torch.randn(n, device="cuda") * 0.1 for n in (3 * W, W, FF, W))
    classes = {
        "qkv": lambda cg: ops.gemm_lnfold(xb, w_qkv, b_qkv, cs_qkv, part, 1e-5, ops.ACT_NONE, cg),
        "out": lambda cg: ops.gemm_resid_ln(xb, w_out, b_out, x, cg),
        "fc": lambda cg: ops.gemm_lnfold(xb, w_fc, b_fc, cs_fc, part, 1e-5, ops.ACT_GELU_ERF, cg),
        "proj": lambda cg: ops.gemm_resid_ln(h, w_proj, b_proj, x, cg),
    }
    line = [f"B={B:3d} M={M:6d}"]
    tot_auto = tot_best = 0.0
    for name, fn in classes.items():
        t = {cg: graph_us(lambda cg=cg: fn(cg)) for cg in (2, 1, 3, 0)}
        best = min((2, 1, 3), key=lambda c: t[c])
        tot_auto += t[0]; tot_best += t[best]
        line.append(f"{name}: cg2 {t[2]:6.1f} cg1 {t[1]:6.1f} cg3 {t[3]:6.1f} auto {t[0]:6.1f} (best cg{best})")
    line.append(f"sum auto {tot_auto:6.1f} best {tot_best:6.1f} us/layer")
    print("  ".join(line), flush=True)
