"""One stage-1 feature extraction (train.py:74-85) at train.py's defaults - 518 px, batch 2, DAPM_layer 20 - for ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/surgery_target.py"""
import sys

import torch

sys.path.insert(0, ".")
from aaclip_b200 import synth  # noqa: E402
from aaclip_b200.clip import CLIP  # noqa: E402
from aaclip_b200.surgery import CLIPImageEncoder, surgery_patch_features  # noqa: E402

cfg = synth.ModelCfg(image_size=518)
model = CLIP(cfg, text=False)
model.load_state_dict(synth.clip_state_dict(cfg, 0, text=False), strict=False)
model = model.cuda()
enc = CLIPImageEncoder(model, [6, 12, 18, 24], surgery_until_layer=20, max_batch=2)
plain = CLIPImageEncoder(model, [], max_batch=2)
img = synth.images(2, cfg, seed=3).cuda()
for _ in range(2):
    feats = surgery_patch_features(enc, plain, img)
torch.cuda.synchronize()
print("ok", [tuple(f.shape) for f in feats])
