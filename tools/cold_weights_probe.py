"""Does a small-batch GEMM pay for weights that come from HBM instead of L2?  Each class of the block is timed as 24
launches in one CUDA graph, once re-reading ONE weight matrix (L2 resident after the first launch) and once rotating over
24 distinct matrices (> 126 MB in total for the large ones: every launch streams its weights from HBM, as in the real step).
    python tools/cold_weights_probe.py [B ...]"""
import sys

import torch

sys.path.insert(0, ".")
from aaclip_b200 import ops  # noqa: E402

W, FF = 1024, 4096
side = torch.cuda.Stream()
NW = 24


def graph_us(fns, reps=5):
    with torch.cuda.stream(side):
        for f in fns[:2]:
            f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for f in fns:
                f()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (len(fns) * reps) * 1e3


for B in [int(a) for a in sys.argv[1:]] or [1, 8]:
    M = B * 577
    xb = (torch.randn(M, W, device="cuda") * 0.5).to(torch.bfloat16)
    part = torch.zeros(M, W // 128, 2, device="cuda")
    part[:, :, 0] = xb.float().view(M, W // 128, 128).sum(-1)
    part[:, :, 1] = (xb.float() ** 2).view(M, W // 128, 128).sum(-1)
    h = (torch.randn(M, FF, device="cuda") * 0.5).to(torch.bfloat16)
    x = torch.randn(M, W, device="cuda")
    mk = lambda n, k: [(torch.randn(n, k, device="cuda") * 0.03).to(torch.bfloat16) for _ in range(NW)]
    w_qkv, w_out, w_fc, w_proj = mk(3 * W, W), mk(W, W), mk(FF, W), mk(W, FF)
    cs_qkv, cs_fc = w_qkv[0].float().sum(1), w_fc[0].float().sum(1)
    b_qkv, b_out, b_fc, b_proj = (torch.randn(n, device="cuda") * 0.1 for n in (3 * W, W, FF, W))
    classes = {
        "qkv": lambda w: ops.gemm_lnfold(xb, w, b_qkv, cs_qkv, part, 1e-5, ops.ACT_NONE, 0),
        "out": lambda w: ops.gemm_resid_ln(xb, w, b_out, x, 0),
        "fc": lambda w: ops.gemm_lnfold(xb, w, b_fc, cs_fc, part, 1e-5, ops.ACT_GELU_ERF, 0),
        "proj": lambda w: ops.gemm_resid_ln(h, w, b_proj, x, 0),
    }
    sets = {"qkv": w_qkv, "out": w_out, "fc": w_fc, "proj": w_proj}
    line = [f"B={B:3d}"]
    hot_sum = cold_sum = 0.0
    for name, fn in classes.items():
        hot = graph_us([lambda fn=fn, w=sets[name][0]: fn(w)] * NW)
        cold = graph_us([lambda fn=fn, w=w: fn(w) for w in sets[name]])
        hot_sum += hot; cold_sum += cold
        line.append(f"{name}: L2-hot {hot:6.1f}  HBM {cold:6.1f} us")
    # the whole block in sequence with distinct weights per "layer" (what the step does)
    seq = []
    for i in range(NW):
        for name, fn in classes.items():
            seq.append(lambda fn=fn, w=sets[name][i]: fn(w))
    blk = graph_us(seq) * 4
    line.append(f"sum hot {hot_sum:6.1f} cold {cold_sum:6.1f}; 24 blocks in sequence: {blk:6.1f} us/block")
    print("  ".join(line), flush=True)
