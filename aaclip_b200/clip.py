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
