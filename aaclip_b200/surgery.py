"""Stage-1 feature extraction of the reference's train.py on the B200 engine (forward only - the
reference runs it under `torch.no_grad()`, train.py:74-85).

    clip_surgery = create_model(...); clip_surgery.visual.DAPM_replace(DPAM_layer=20)        train.py:235-243
    _, patch_features = clip_surgery.encode_image(image, [6, 12, 18, 24])                    train.py:75
    cls_token, _ = adapted_model.clipmodel.encode_image(image, [])                           train.py:76
    ... ln_post -> @ visual.proj -> / norm -> + cls_token                                    train.py:77-85

becomes

    surgery = CLIPImageEncoder(clip_model, [6, 12, 18, 24], surgery_until_layer=20)
    plain   = CLIPImageEncoder(clip_model, [])
    patch_features = surgery_patch_features(surgery, plain, image)

`DAPM_replace` swaps the attention of the last DPAM_layer - 1 visual blocks for `Attention`
(model/transformer.py:102-152), whose forward reads the block's [L, batch, D] tensor as (B, N, C): the softmax runs
over the images of the batch for every token position and head.  That is reproduced faithfully (csrc/vv_attn.cu), so
features depend on the batch composition exactly as they do in the reference; a batch is never split.

No CPU / PyTorch fallback: everything below runs in libaaclip_b200.so.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops
from .adapter import _infer_cfg, effective_levels
from .engine import Engine


class CLIPImageEncoder:
    """`CLIP.encode_image` (model/model.py:185-188) of a CLIP's visual tower; `surgery_until_layer` = the DPAM_layer of
    `VisionTransformer.DAPM_replace` (model/transformer.py:406-425; train.py:187 default 20).

    clip_model: the reference's `CLIP` (or aaclip_b200.clip.CLIP) - anything with `.visual` carrying the reference's
    parameter names; its parameters are re-uploaded when they change (tracked by tensor version, like AdaptedCLIP)."""

    def __init__(self, clip_model, out_layers: Sequence[int] = (6, 12, 18, 24), surgery_until_layer: Optional[int] = None,
                 max_batch: int = 8):
        self.clipmodel = clip_model
        self.out_layers = list(out_layers)
        self.surgery_until_layer = surgery_until_layer
        self.max_batch = max_batch
        self.out_layers_effective = False
        self._engine: Optional[Engine] = None
        self._versions = {}
        # a reference model whose visual tower already went through DAPM_replace carries `attn.qkv` / `attn.proj`
        # Linears in the replaced blocks: same tensors under other names, and the depth can be read off the keys
        replaced = sum(1 for k in clip_model.state_dict().keys()
                       if k.startswith("visual.transformer.resblocks.") and k.endswith(".attn.qkv.weight"))
        if replaced and surgery_until_layer is None:
            self.surgery_until_layer = replaced + 1

    # ---------------------------------------------------------------------------------- reference surface
    def DAPM_replace(self, DPAM_layer: Optional[int]) -> None:
        """model/transformer.py:406-425."""
        self.surgery_until_layer = DPAM_layer
        if self._engine is not None:
            self._engine.dapm_replace(DPAM_layer)

    def encode_image(self, image: torch.Tensor, out_layers: Optional[Sequence[int]] = None, normalize: bool = False):
        """-> (pooled [B, E], [tokens [B, L, width] after each block in out_layers]) like model/model.py:185-188.
        out_layers defaults to the constructor's; a different list must be a subset of it."""
        eng = self._sync(image.device)
        mine = eng.cfg.levels if self.out_layers_effective else []
        want = mine if out_layers is None else effective_levels(out_layers, eng.cfg.layers)
        missing = [l for l in want if l not in mine]
        if missing:
            raise ValueError(f"out_layers {missing} were not requested at construction ({self.out_layers})")
        pooled, toks = eng.encode_image(image, want_tokens=bool(want), want_pooled=True, normalize=normalize)
        return pooled, [t for l, t in zip(eng.cfg.levels, toks) if l in want]

    def patch_features(self, image: torch.Tensor) -> List[torch.Tensor]:
        """train.py:78-84: normalize(ln_post(tokens[:, 1:]) @ visual.proj) for every out_layer -> list of [B, P, E]."""
        eng = self._sync(image.device)
        if not self.out_layers_effective:
            raise ValueError("this encoder was built without out_layers")
        seg, _ = eng.visual_forward(image, want_seg=True, want_det=False)
        return seg

    # ---------------------------------------------------------------------------------- engine plumbing
    def _sync(self, device: torch.device) -> Engine:
        if device.type != "cuda":
            raise RuntimeError("aaclip_b200.CLIPImageEncoder runs on a B200 only (no CPU / PyTorch fallback)")
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        if self._engine is None or self._engine.device != dev_index:
            cfg = _infer_cfg(self.clipmodel, [], 0, 0, 0.0, 0.0, False)
            cfg.embed_dim = int(self.clipmodel.visual.proj.shape[1])
            cfg.t_layers = 0
            lv = effective_levels(self.out_layers, cfg.layers)
            self.out_layers_effective = bool(lv)
            cfg.levels = lv or [cfg.layers]          # `[]` (train.py:76): only the pooled feature is wanted
            self._engine = Engine(cfg, device=dev_index, max_batch=self.max_batch, text=False)
            self._engine.dapm_replace(self.surgery_until_layer)
            self._versions = {}
        eng = self._engine
        dirty = False
        seen = set()
        for k, t in self.clipmodel.state_dict(keep_vars=True).items():
            key = "clip." + _canonical_key(k)
            seen.add(key)
            sig = (t.data_ptr(), t._version)
            if k == "visual.proj":
                if self._versions.get(key) != sig:
                    w = t.detach().float().t().contiguous()          # nn.Linear layout [E, width]
                    for i in range(len(eng.cfg.levels)):
                        eng.set_weight(f"image_adapter.seg_proj.{i}.fc.weight", w)
                    self._versions[key] = sig
                    dirty = True
                continue
            if key not in eng._wmap:
                continue
            if self._versions.get(key) != sig:
                eng.set_weight(key, t)
                self._versions[key] = sig
                dirty = True
        if dirty:
            torch.cuda.synchronize(dev_index)
        missing = [k for k in eng._wmap if k.startswith("clip.visual.") and k not in seen]
        if missing or "clip.visual.proj" not in seen:
            raise KeyError(f"clip_model.state_dict() lacks {missing[:4] or ['visual.proj']} (+{max(len(missing) - 4, 0)} more)")
        return eng


def _canonical_key(k: str) -> str:
    """`Attention` (model/transformer.py:102-121) names of a DAPM-replaced block -> nn.MultiheadAttention names."""
    if k.startswith("visual.transformer.resblocks."):
        for a, b in ((".attn.qkv.weight", ".attn.in_proj_weight"), (".attn.qkv.bias", ".attn.in_proj_bias"),
                     (".attn.proj.weight", ".attn.out_proj.weight"), (".attn.proj.bias", ".attn.out_proj.bias")):
            if k.endswith(a):
                return k[: -len(a)] + b
    return k


def surgery_patch_features(clip_surgery: CLIPImageEncoder, clip_plain: CLIPImageEncoder,
                           image: torch.Tensor) -> List[torch.Tensor]:
    """train.py:74-85 in one call: the surgery tower's projected, normalised patch tokens at its out_layers, plus the
    normalised pooled class feature of the unmodified CLIP, broadcast over the patches."""
    feats = clip_surgery.patch_features(image)                                  # :75, :78-84
    cls, _ = clip_plain.encode_image(image, [], normalize=True)                 # :76-77
    return [ops.add_image_vector(t, cls) for t in feats]                        # :85
