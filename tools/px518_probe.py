"""the paper's 518-px operating point (37 x 37 + 1 = 1370 tokens) at full depth: images/s of the fused
forward (encoder + adapters + head) and its tensor-roofline fraction on the per-image FLOPs at L = 1370."""
import json, sys, torch
sys.path.insert(0, ".")
from aaclip_b200 import synth
from aaclip_b200.engine import Engine
pk = json.load(open("MEASURED_PEAKS.json"))
cfg = synth.ModelCfg(image_size=518, t_layers=0)
L, P, d, ff, E = cfg.tokens, cfg.patches, cfg.width, cfg.mlp_width, cfg.embed_dim
blk = 2 * L * d * 3 * d + 4 * L * L * d + 2 * L * d * d + 4 * L * d * ff          # qkv, QK^T + PV, out, fc + proj
F_IMG = cfg.layers * blk + 2 * P * 3 * 14 * 14 * d + 6 * 2 * L * d * d + 5 * 2 * P * d * E
print(f"# 518 px: L={L}, {F_IMG / 1e9:.1f} GFLOP/image (336 px: 393.7)")
for B in (16, 32, 64):
    eng = Engine(cfg, device=0, max_batch=B, text=False)
    eng.load_state_dicts(synth.clip_state_dict(cfg, 0, text=False), synth.image_adapter_state_dict(cfg, 0), None)
    img, T = synth.images(B, cfg, seed=3).cuda(), synth.anchors(cfg, seed=1).cuda()
    for _ in range(3): eng.forward_fused(img, T)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 8
    e0.record()
    for _ in range(n): maps, scores = eng.forward_fused(img, T)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = F_IMG * B / ms / 1e9
    print(f"B={B:3d}: {ms:8.2f} ms/batch  {B / ms * 1e3:7.1f} images/s  {tf:6.0f} TFLOP/s  ({tf / pk['bf16_tflops_sustained']:.3f} of sustained, "
          f"{tf / pk['bf16_tflops']:.3f} of burst)  finite={bool(torch.isfinite(maps).all())}")
    eng.close(); del eng, img
