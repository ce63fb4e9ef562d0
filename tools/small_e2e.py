"""A small end-to-end pass over every kernel family with ragged shapes (B = 3 against max_batch = 2, 97 x 131 raw images,
2-layer ViT-L-width model, both LayerNorm schedules, text tower, fused and drop-in paths, head modes, extrema)."""
import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
from aaclip_b200 import synth, ops, forward_utils as fu
from aaclip_b200.engine import Engine
cfg = synth.ModelCfg(layers=2, t_layers=2, image_adapt_until=1, levels=[1, 2], text_adapt_until=1)
for fold in (True, False):
    eng = Engine(cfg, device=0, max_batch=2, max_text=4, ln_fold=fold)
    eng.load_state_dicts(synth.clip_state_dict(cfg, 0), synth.image_adapter_state_dict(cfg, 0), synth.text_adapter_state_dict(cfg, 0))
    img, T = synth.images(3, cfg, seed=1).cuda(), synth.anchors(cfg, seed=1).cuda()
    seg, det = eng.visual_forward(img)
    maps, scores = eng.forward_fused(img, T)
    emb = eng.text_forward(synth.tokens(3, cfg, seed=2))
    raw = torch.randint(0, 256, (3, 97, 131, 3), dtype=torch.uint8)
    out = list(eng.predict_stream([raw, img.cpu()], T))
    torch.cuda.synchronize()
    print("fold", fold, float(maps.abs().max()), float(scores.mean()), float(emb.abs().max()), out[0][0].shape)
    eng.close()
m, s = ops.anomaly_head([t.to(torch.bfloat16) for t in seg], T, cfg.image_size, ops.HEAD_TEST_MEDICAL, det=det)
tr, _ = ops.anomaly_head(seg, T, cfg.image_size, ops.HEAD_TRAIN_SOFTMAX)
ex = fu.map_extrema(m)
u8 = ops.resize_bicubic_u8(torch.randint(0, 256, (2, 50, 333, 3), dtype=torch.uint8, device="cuda"), 224)
torch.cuda.synchronize()
print("ok", ex.shape, tr.shape, u8.shape)
