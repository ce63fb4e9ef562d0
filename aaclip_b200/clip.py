"""A parameter-only stand-in for the reference `CLIP` container (model/model.py:149-212).

The reference's `AdaptedCLIP(clip_model, ...)` receives a fully built CLIP module and reads parameters off it
by attribute (`clip_model.visual.conv1`, `.transformer.resblocks[i].attn.in_proj_weight`, ...).  Our drop-in
`AdaptedCLIP` accepts either the reference's own `CLIP` object or this stand-in, which reproduces the same
attribute tree and state_dict keys so checkpoints load unchanged (`load_state_dict`) - it just carries no
forward code, because the arithmetic lives in the CUDA engine.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch
from torch import nn

from .synth import ModelCfg, VIT_L_14_336


class QuickGELU(nn.Module):
    """Marker with the reference's class name (model/transformer.py:46-49); selects the QuickGELU epilogue."""


class _Block(nn.Module):
    """Names of ResidualAttentionBlock (model/transformer.py:183-224)."""

    def __init__(self, d_model: int, n_head: int, mlp_width: int, quick_gelu: bool):
        super().__init__()
        self.ln_1 = nn.LayerNorm(d_model)
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_2 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, mlp_width)),
            ("gelu", QuickGELU() if quick_gelu else nn.GELU()),
            ("c_proj", nn.Linear(mlp_width, d_model)),
        ]))


class _Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, mlp_width: int, quick_gelu: bool):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.ModuleList([_Block(width, heads, mlp_width, quick_gelu) for _ in range(layers)])

    def get_cast_dtype(self) -> torch.dtype:
        return self.resblocks[0].mlp.c_fc.weight.dtype


class _Visual(nn.Module):
    """Names of VisionTransformer (model/transformer.py:320-404)."""

    def __init__(self, cfg: ModelCfg):
        super().__init__()
        w = cfg.width
        self.image_size = (cfg.image_size, cfg.image_size)
        self.patch_size = (cfg.patch_size, cfg.patch_size)
        self.grid_size = (cfg.grid, cfg.grid)
        self.output_dim = cfg.embed_dim
        self.conv1 = nn.Conv2d(3, w, kernel_size=cfg.patch_size, stride=cfg.patch_size, bias=False)
        scale = w ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(w))
        self.positional_embedding = nn.Parameter(scale * torch.randn(cfg.tokens, w))
        self.patch_dropout = nn.Identity()
        self.ln_pre = nn.LayerNorm(w)
        self.transformer = _Transformer(w, cfg.layers, cfg.heads, cfg.mlp_width, cfg.quick_gelu)
        self.ln_post = nn.LayerNorm(w)
        self.proj = nn.Parameter(scale * torch.randn(w, cfg.embed_dim))


class CLIP(nn.Module):
    def __init__(self, cfg: ModelCfg = VIT_L_14_336, text: bool = True):
        super().__init__()
        self.cfg = cfg
        self.visual = _Visual(cfg)
        if text:
            tw = cfg.t_width
            self.transformer = _Transformer(tw, cfg.t_layers, cfg.t_heads, 4 * tw, cfg.quick_gelu)
            self.context_length = cfg.t_context
            self.vocab_size = cfg.t_vocab
            self.token_embedding = nn.Embedding(cfg.t_vocab, tw)
            self.positional_embedding = nn.Parameter(torch.empty(cfg.t_context, tw).normal_(std=0.01))
            self.ln_final = nn.LayerNorm(tw)
            self.text_projection = nn.Parameter(torch.empty(tw, cfg.embed_dim).normal_(std=tw ** -0.5))
            self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592600345611572)
            mask = torch.full((cfg.t_context, cfg.t_context), float("-inf")).triu_(1)
            self.register_buffer("attn_mask", mask, persistent=False)

    def encode_text(self, text, normalize: bool = False):
        raise NotImplementedError("un-adapted CLIP.encode_text is outside the accelerated hot path (SURVEY 8)")

    def encode_image(self, image, *a, **k):
        raise NotImplementedError("un-adapted CLIP.encode_image is outside the accelerated hot path (SURVEY 8)")


def create_model(model_name: str = "ViT-L-14-336", img_size: int = 336, pretrained: Optional[str] = None,
                 text: bool = True, **_) -> CLIP:
    """Shape-compatible subset of model/clip.py:create_model for the single shipped config (random init)."""
    if model_name.replace("/", "-") != "ViT-L-14-336":
        raise RuntimeError(f"Model config for {model_name} not found.")
    if pretrained:
        raise RuntimeError("pretrained weights are not available offline; load a state_dict instead")
    cfg = ModelCfg(image_size=img_size)
    return CLIP(cfg, text=text)


def resize_pos_embed(state_dict, grid_size: int, interpolation: str = "bicubic", antialias: bool = True) -> bool:
    """Load-time rescale of `visual.positional_embedding` when a checkpoint trained at one resolution is used at
    another (the paper's 518-px setting from the 336-px OpenAI weights).  Semantics of model/model.py:395-426: the
    class-token row is kept, the patch grid is resampled with F.interpolate(bicubic, antialias=True,
    align_corners=False).  Mutates `state_dict`; returns True if a resize happened.  Host-side load-time plumbing."""
    import math

    import torch.nn.functional as F
    old = state_dict.get("visual.positional_embedding", None)
    if old is None:
        return False
    new_len = grid_size * grid_size + 1
    if new_len == old.shape[0]:
        return False
    tok, img = old[:1], old[1:]
    g0 = int(math.sqrt(img.shape[0]))
    if g0 * g0 != img.shape[0]:
        raise ValueError(f"positional embedding of {old.shape[0]} rows is not 1 + a square grid")
    img = img.reshape(1, g0, g0, -1).permute(0, 3, 1, 2)
    img = F.interpolate(img.float(), size=(grid_size, grid_size), mode=interpolation, antialias=antialias,
                        align_corners=False)
    img = img.permute(0, 2, 3, 1).reshape(grid_size * grid_size, -1).to(old.dtype)
    state_dict["visual.positional_embedding"] = torch.cat([tok, img], dim=0)
    return True


def load_checkpoint(model: CLIP, state_dict, strict: bool = True):
    """model/clip.py:60-82 (load_checkpoint) for the container: drops the keys the reference drops, rescales the
    positional embedding to the model's grid, then load_state_dict.  `state_dict` may be an OpenAI / open_clip
    ViT-L/14 state dict (model/openai.py:17-83 converts to the same keys) or ours."""
    sd = dict(state_dict)
    for k in ("input_resolution", "context_length", "vocab_size"):
        sd.pop(k, None)
    resize_pos_embed(sd, model.visual.grid_size[0])
    own = model.state_dict()
    if not strict:
        sd = {k: v for k, v in sd.items() if k in own}
    return model.load_state_dict(sd, strict=strict)
