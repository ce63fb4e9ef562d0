"""Drop-ins for the hot-path functions of the reference's forward_utils.py, backed by the CUDA head kernels.

  calculate_similarity_map     forward_utils.py:196-216   (same signature, same return shapes)
  class_text_embedding         forward_utils.py:153-161   (tokenised prompts -> [768, 2] anchor)
  get_predictions_batch        test.py:80-93               (fused: image batch -> summed maps + image scores)
  transform_x                  dataset/__init__.py:127-136 (Resize BICUBIC + ToTensor + Normalize, on the device)
  map_extrema / image_level_preds   forward_utils.py:241-254 (metrics_eval's min-max normalisation + pmax mix)

String work (prompt tables, BPE tokenizer: dataset/constants.py, model/tokenizer.py) stays with the caller:
`class_text_embedding` starts from token ids.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import ops

DOMAIN_MODE = {"Industrial": ops.HEAD_TEST_INDUSTRIAL, "Medical": ops.HEAD_TEST_MEDICAL}


def _mode(test: bool, domain: str) -> int:
    if not test:
        return ops.HEAD_TRAIN_SOFTMAX
    # forward_utils.py:205-206: anything that is not "Industrial" takes the 9x9 / sigma 1.5 branch
    return ops.HEAD_TEST_INDUSTRIAL if domain == "Industrial" else ops.HEAD_TEST_MEDICAL


@torch.no_grad()
def calculate_similarity_map(patch_features: torch.Tensor, epoch_text_feature: torch.Tensor, img_size: int,
                             test: bool = False, domain: str = "Medical") -> torch.Tensor:
    """[B,L,768] x ([768,C] | [B,768,C]) -> test: [B,1,img,img]; train: [B,C,img,img] softmaxed (C == 2)."""
    if epoch_text_feature.shape[-1] != 2:
        # the reference asserts C == 2 in test mode (forward_utils.py:204); the kernels are built for C == 2
        raise AssertionError("calculate_similarity_map: C must be 2")
    pf = patch_features if patch_features.dtype in (torch.float32, torch.bfloat16) else patch_features.float()
    maps, _ = ops.anomaly_head([pf.contiguous()], epoch_text_feature.float().contiguous(), int(img_size),
                               _mode(test, domain))
    if test:
        return maps.unsqueeze(1)
    return maps[0]


@torch.no_grad()
def similarity_maps_summed(patch_features: Sequence[torch.Tensor], epoch_text_feature: torch.Tensor, img_size: int,
                           domain: str = "Industrial", det_feature: torch.Tensor = None, with_extrema: bool = False):
    """test.py:83-93 in one call: torch.cat([calculate_similarity_map(f, ..., test=True) ...], 1).sum(1) and
    (optionally) the image score ((det @ T)[:, 1] + 1) / 2 - one launch of the streaming head kernel that reads every
    level's tokens (fp32 as the reference's model returns them, or bf16) once.  `with_extrema` adds the per-image
    (min, max) [B,2] of the maps (metrics_eval's normalisation input, forward_utils.py:241-252)."""
    feats = [(f if f.dtype in (torch.float32, torch.bfloat16) else f.float()).contiguous() for f in patch_features]
    return ops.anomaly_head(feats, epoch_text_feature.float().contiguous(), int(img_size), _mode(True, domain),
                            det=None if det_feature is None else det_feature.float().contiguous(),
                            want_extrema=with_extrema)


@torch.no_grad()
def class_text_embedding(model, tokens_normal: torch.Tensor, tokens_abnormal: torch.Tensor) -> torch.Tensor:
    """forward_utils.py:147-161 from token ids: encode_text -> row L2 norm -> mean -> L2 norm, stacked [768,2]."""
    from ._lib import check, cur_stream, load, ptr
    embs = [model.encode_text(t) for t in (tokens_normal, tokens_abnormal)]
    out = torch.empty(embs[0].shape[1], 2, device=embs[0].device, dtype=torch.float32)
    lib = load()
    for col, e in enumerate(embs):
        check(lib.aaclip_text_anchor(ptr(e), e.shape[0], e.shape[1], ptr(out), col, cur_stream(e.device)))
    return out


@torch.no_grad()
def get_predictions_batch(model, image: torch.Tensor, epoch_text_feature: torch.Tensor, domain: str = "Industrial"):
    """One iteration of test.py:get_predictions (lines 80-93) on the fused path: (maps [B,S,S], scores [B])."""
    return model.predict(image, epoch_text_feature, domain)


@torch.no_grad()
def transform_x(images_u8: torch.Tensor, img_size: int) -> torch.Tensor:
    """The loader's image transform (dataset/__init__.py:127-136: transforms.Resize((S,S), Image.BICUBIC), ToTensor,
    Normalize with the CLIP statistics) for a batch of equally sized raw RGB images uint8 [B,H0,W0,3] on the device;
    returns float32 [B,3,S,S], bit-exact with what the reference's dataset returns."""
    return ops.preprocess_u8(images_u8.contiguous(), int(img_size))


@torch.no_grad()
def map_extrema(maps: torch.Tensor) -> torch.Tensor:
    """Per-image (min, max) of a batch of anomaly maps [B,S,S] on the device -> fp32 [B,2].  With these, the
    image-level half of metrics_eval (forward_utils.py:241-254) needs no pixel data on the host."""
    return ops.map_minmax(maps.contiguous())


def image_level_preds(extrema, image_preds, domain: str = "Industrial"):
    """metrics_eval's image score (forward_utils.py:241-254) from per-image map extrema [N,2] (all images of the
    class, any number of batches) and the raw image scores [N]:

        pixel_preds = (p - p.min()) / (p.max() - p.min())   unless p.max() == 1      (:241-244)
        image_preds = (s - s.min()) / (s.max() - s.min())   unless s.max() == 1      (:245-248)
        pmax = pixel_preds.max(axis=(1, 2));  image = 0.5 pmax + 0.5 image_preds     (:250-252; Medical: pmax)

    min / max commute with the affine normalisation, so only the extrema are needed.  Host arithmetic on N numbers,
    in numpy float32 like the reference."""
    import numpy as np
    ex = np.asarray(extrema.cpu() if hasattr(extrema, "cpu") else extrema, dtype=np.float32)
    s = np.asarray(image_preds.cpu() if hasattr(image_preds, "cpu") else image_preds, dtype=np.float32)
    pmax = ex[:, 1].copy()
    gmin, gmax = ex[:, 0].min(), ex[:, 1].max()
    if gmax != 1:
        pmax = (pmax - gmin) / (gmax - gmin)
    if s.max() != 1:
        s = (s - s.min()) / (s.max() - s.min())
    return pmax * 0.5 + s * 0.5 if domain != "Medical" else pmax
