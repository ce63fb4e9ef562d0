"""BASELINE.json configs[3] and configs[4] on one B200 (not bench lines; reported under profiles/):
  * batch sweep B = 1 .. 2048 (chunks of <= 256 images) with all four taps: images/s, tensor-roofline fraction
    (393.7 GFLOP/image over the measured bf16 peaks) and the head kernel's HBM fraction at that batch;
  * text-anchor path: 15 classes x (7 normal + 9 abnormal ... 240 sentences in all) through the text tower with the
    text adapter, anchors [768,2] per class, then similarity against cached patch features.
usage: python tools/sweep.py [max_B] > profiles/rN_sweep.txt"""
import json, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aaclip_b200 import ops, synth
from aaclip_b200.engine import Engine
F_IMG, HEAD_BYTES = 393_708_404_736, 3_993_604
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else \
    {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}
max_B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cfg = synth.VIT_L_14_336

def ev_time(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

eng = Engine(cfg, device=0, max_batch=min(256, max_B), max_text=256, text=True)
eng.load_state_dicts(synth.clip_state_dict(cfg, 0), synth.image_adapter_state_dict(cfg, 0), synth.text_adapter_state_dict(cfg, 0))
T = synth.anchors(cfg, 1).cuda()
print(f"# batch sweep (ViT-L/14-336, 4 taps, fused head); peaks: bf16 {pk['bf16_tflops']} burst / {pk['bf16_tflops_sustained']} sustained TF/s, HBM {pk['hbm_gbs']} GB/s")
print(f"{'B':>6} {'ms/batch':>10} {'images/s':>10} {'TF/s':>8} {'of burst':>9} {'of sust.':>9} {'head us':>9} {'head GB/s':>10} {'of HBM':>7}")
B = 1
while B <= max_B:
    g = torch.Generator(device="cuda").manual_seed(B)
    img = torch.randn(B, 3, 336, 336, device="cuda", generator=g)
    for _ in range(2): eng.forward_fused(img, T)
    reps = max(2, min(20, 2048 // B))
    ms = ev_time(lambda: eng.forward_fused(img, T), reps)
    tf = F_IMG * B / (ms / 1e3) / 1e12
    hb = min(B, 256)
    feats = [torch.nn.functional.normalize(torch.randn(hb, 576, 768, device="cuda"), dim=-1).bfloat16() for _ in range(4)]
    det = torch.randn(hb, 768, device="cuda")
    for _ in range(2): ops.anomaly_head(feats, T, 336, ops.HEAD_TEST_INDUSTRIAL, det=det)
    hms = ev_time(lambda: ops.anomaly_head(feats, T, 336, ops.HEAD_TEST_INDUSTRIAL, det=det), 10)
    hgb = HEAD_BYTES * hb / (hms / 1e3) / 1e9
    print(f"{B:6d} {ms:10.3f} {B / (ms / 1e3):10.1f} {tf:8.1f} {tf / pk['bf16_tflops']:9.3f} {tf / pk['bf16_tflops_sustained']:9.3f} "
          f"{hms * 1e3:9.1f} {hgb:10.1f} {hgb / pk['hbm_gbs']:7.3f}")
    del img, feats
    B *= 2

# ---- text-anchor path: 15 classes, 240 sentences (MVTec prompt tables: dataset/constants.py:135-147), synthetic ids
n_cls, n_norm, n_abn = 15, 7, 9
tok = synth.tokens(n_cls * (n_norm + n_abn), cfg, seed=2).cuda()
def anchors_all():
    emb = eng.text_forward(tok)                     # one batched pass over all 240 sentences
    out = torch.empty(n_cls, 768, 2, device="cuda")
    lib = eng.lib
    from aaclip_b200._lib import check, cur_stream, ptr
    for c in range(n_cls):
        base = c * (n_norm + n_abn)
        check(lib.aaclip_text_anchor(ptr(emb[base:base + n_norm]), n_norm, 768, ptr(out[c]), 0, cur_stream()))
        check(lib.aaclip_text_anchor(ptr(emb[base + n_norm:base + n_norm + n_abn]), n_abn, 768, ptr(out[c]), 1, cur_stream()))
    return out
for _ in range(2): A = anchors_all()
ms = ev_time(anchors_all, 5)
print(f"\n# text-anchor path: {tok.shape[0]} sentences x 77 tokens, 12-layer text tower + text adapter -> {n_cls} anchors [768,2]: "
      f"{ms:.3f} ms ({tok.shape[0] / (ms / 1e3):.0f} sentences/s)")
feats = [torch.nn.functional.normalize(torch.randn(64, 576, 768, device="cuda"), dim=-1).bfloat16() for _ in range(4)]
ms2 = ev_time(lambda: [ops.anomaly_head(feats, A[c].contiguous(), 336, ops.HEAD_TEST_INDUSTRIAL) for c in range(n_cls)], 5)
print(f"# similarity of 64 cached images x 4 levels against each of the {n_cls} class anchors: {ms2:.3f} ms "
      f"({64 * n_cls / (ms2 / 1e3):.0f} maps/s)")
eng.close()
