"""Deterministic synthetic weights and inputs (there are no checkpoints offline; BASELINE.json pins random init).

Every tensor is drawn from its own CPU generator seeded by crc32(key) ^ seed, so the same state dict is
reproduced bit-for-bit wherever it is needed (golden generation next to the reference, the oracle, the
CUDA engine, the benchmark) without shipping 1.2 GB of weights.  Keys and shapes are the reference's
state_dict keys (CLIP: model/model.py:149-212; adapters: model/adapter.py:27-44); scales follow the
reference's default initialisers, except that LayerNorm affine parameters and attention biases are
perturbed away from their (1, 0) defaults so that parity tests exercise them.
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass, field
from typing import Dict, List

import torch


@dataclass
class ModelCfg:
    """model/model_configs/ViT-L-14-336.json + AdaptedCLIP ctor defaults (model/adapter.py:7-17)."""
    image_size: int = 336
    patch_size: int = 14
    width: int = 1024
    layers: int = 24
    heads: int = 16
    mlp_ratio: float = 4.0
    embed_dim: int = 768
    quick_gelu: bool = False
    t_context: int = 77
    t_vocab: int = 49408
    t_width: int = 768
    t_heads: int = 12
    t_layers: int = 12
    text_adapt_weight: float = 0.1
    image_adapt_weight: float = 0.1
    text_adapt_until: int = 3
    image_adapt_until: int = 6
    levels: List[int] = field(default_factory=lambda: [6, 12, 18, 24])
    relu: bool = False

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def patches(self) -> int:
        return self.grid * self.grid

    @property
    def tokens(self) -> int:
        return self.patches + 1

    @property
    def mlp_width(self) -> int:
        return int(self.width * self.mlp_ratio)


VIT_L_14_336 = ModelCfg()


def tiny_cfg(**kw) -> ModelCfg:
    """A small configuration with the same structure (for fast host-side tests)."""
    base = dict(image_size=56, patch_size=14, width=256, layers=4, heads=4, embed_dim=256, t_context=16,
                t_vocab=512, t_width=256, t_heads=4, t_layers=3, text_adapt_until=2, image_adapt_until=2,
                levels=[1, 2, 3, 4])
    base.update(kw)
    return ModelCfg(**base)


def _draw(key: str, shape, std: float, seed: int, mean: float = 0.0) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    t = torch.randn(*shape, generator=g, dtype=torch.float32)
    return t * std + mean


def _block(sd: Dict[str, torch.Tensor], prefix: str, w: int, ff: int, seed: int, attn_std: float, proj_std: float,
           fc_std: float) -> None:
    d = lambda k, shape, std, mean=0.0: sd.__setitem__(prefix + k, _draw(prefix + k, shape, std, seed, mean))
    d("ln_1.weight", (w,), 0.1, 1.0); d("ln_1.bias", (w,), 0.05)
    d("attn.in_proj_weight", (3 * w, w), attn_std); d("attn.in_proj_bias", (3 * w,), 0.02)
    d("attn.out_proj.weight", (w, w), proj_std); d("attn.out_proj.bias", (w,), 0.02)
    d("ln_2.weight", (w,), 0.1, 1.0); d("ln_2.bias", (w,), 0.05)
    d("mlp.c_fc.weight", (ff, w), fc_std); d("mlp.c_fc.bias", (ff,), w ** -0.5 / 3 ** 0.5)
    d("mlp.c_proj.weight", (w, ff), ff ** -0.5 / 3 ** 0.5); d("mlp.c_proj.bias", (w,), ff ** -0.5 / 3 ** 0.5)


def clip_state_dict(cfg: ModelCfg = VIT_L_14_336, seed: int = 0, text: bool = True) -> Dict[str, torch.Tensor]:
    """State dict with the reference CLIP's keys (the subset the hot path reads, plus visual.proj /
    text_projection / logit_scale so that a strict load into the reference module succeeds)."""
    sd: Dict[str, torch.Tensor] = {}
    w, ff = cfg.width, cfg.mlp_width
    d = lambda k, shape, std, mean=0.0: sd.__setitem__(k, _draw(k, shape, std, seed, mean))
    d("visual.class_embedding", (w,), w ** -0.5)
    d("visual.positional_embedding", (cfg.tokens, w), w ** -0.5)
    d("visual.proj", (w, cfg.embed_dim), w ** -0.5)
    fan_in = 3 * cfg.patch_size ** 2
    d("visual.conv1.weight", (w, 3, cfg.patch_size, cfg.patch_size), fan_in ** -0.5 / 3 ** 0.5)
    d("visual.ln_pre.weight", (w,), 0.1, 1.0); d("visual.ln_pre.bias", (w,), 0.05)
    d("visual.ln_post.weight", (w,), 0.1, 1.0); d("visual.ln_post.bias", (w,), 0.05)
    for i in range(cfg.layers):
        _block(sd, f"visual.transformer.resblocks.{i}.", w, ff, seed, attn_std=(2.0 / (4 * w)) ** 0.5,
               proj_std=w ** -0.5 / 3 ** 0.5, fc_std=w ** -0.5 / 3 ** 0.5)
    if text:
        tw, tff = cfg.t_width, 4 * cfg.t_width
        d("positional_embedding", (cfg.t_context, tw), 0.01)
        d("text_projection", (tw, cfg.embed_dim), tw ** -0.5)
        sd["logit_scale"] = torch.tensor(2.6592600345611572)
        d("token_embedding.weight", (cfg.t_vocab, tw), 0.02)
        d("ln_final.weight", (tw,), 0.1, 1.0); d("ln_final.bias", (tw,), 0.05)
        attn_std = tw ** -0.5
        proj_std = tw ** -0.5 * (2 * cfg.t_layers) ** -0.5
        fc_std = (2 * tw) ** -0.5
        for i in range(cfg.t_layers):
            _block(sd, f"transformer.resblocks.{i}.", tw, tff, seed, attn_std=attn_std, proj_std=proj_std,
                   fc_std=fc_std)
    return sd


def image_adapter_state_dict(cfg: ModelCfg = VIT_L_14_336, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Keys of AdaptedCLIP.image_adapter.state_dict(); xavier-uniform-like scale."""
    sd: Dict[str, torch.Tensor] = {}
    w, e = cfg.width, cfg.embed_dim
    fc = "fc.0.weight" if cfg.relu else "fc.weight"
    for i in range(cfg.image_adapt_until):
        k = f"layer_adapters.{i}.fc.0.weight"
        sd[k] = _draw("image_adapter." + k, (w, w), (2.0 / (w + w)) ** 0.5, seed)
    for i in range(len(cfg.levels)):
        k = f"seg_proj.{i}.{fc}"
        sd[k] = _draw("image_adapter." + k, (e, w), (2.0 / (w + e)) ** 0.5, seed)
    k = f"det_proj.{fc}"
    sd[k] = _draw("image_adapter." + k, (e, w), (2.0 / (w + e)) ** 0.5, seed)
    return sd


def text_adapter_state_dict(cfg: ModelCfg = VIT_L_14_336, seed: int = 0) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    tw = cfg.t_width
    for i in range(cfg.text_adapt_until + 1):
        k = f"{i}.fc.0.weight"
        sd[k] = _draw("text_adapter." + k, (tw, tw), (1.0 / tw) ** 0.5, seed)
    return sd


def images(batch: int, cfg: ModelCfg = VIT_L_14_336, seed: int = 1) -> torch.Tensor:
    """Synthetic CLIP-normalised images: N(0,1) like the post-normalisation pixel distribution."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(batch, 3, cfg.image_size, cfg.image_size, generator=g, dtype=torch.float32)


def anchors(cfg: ModelCfg = VIT_L_14_336, seed: int = 1) -> torch.Tensor:
    """Synthetic text anchors [E, 2] with unit-norm columns (col 0 normal, col 1 abnormal)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + 7919)
    t = torch.randn(cfg.embed_dim, 2, generator=g, dtype=torch.float32)
    return t / t.norm(dim=0, keepdim=True)


def tokens(n: int, cfg: ModelCfg = VIT_L_14_336, seed: int = 2) -> torch.Tensor:
    """Synthetic CLIP token rows: <sot> body... <eot> 0 0 ...  (model/tokenizer.py:150-185 layout)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    out = torch.zeros(n, cfg.t_context, dtype=torch.int32)
    sot, eot = cfg.t_vocab - 2, cfg.t_vocab - 1
    for i in range(n):
        body = int(torch.randint(3, cfg.t_context - 2, (1,), generator=g))
        out[i, 0] = sot
        out[i, 1:1 + body] = torch.randint(1, cfg.t_vocab - 2, (body,), generator=g, dtype=torch.int32)
        out[i, 1 + body] = eot
    return out


def head_inputs(batch: int, grid: int, embed_dim: int = 768, levels: int = 4, seed: int = 5):
    """Synthetic inputs of the anomaly-map head alone: L2-normalised patch tokens per level, per-sample
    anchors [B,E,2] (train-mode form, train.py:69-72) and a det token.  Returns (feats, anchors_batched, det)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    norm = torch.nn.functional.normalize
    feats = [norm(torch.randn(batch, grid * grid, embed_dim, generator=g), dim=-1) for _ in range(levels)]
    tb = norm(torch.randn(batch, embed_dim, 2, generator=g), dim=1)
    det = norm(torch.randn(batch, embed_dim, generator=g), dim=-1) * 0.3
    return feats, tb, det


def openai_style_state_dict(cfg: ModelCfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    """clip_state_dict(cfg, seed) in the shape of an OpenAI JIT archive's state dict: the tensors
    convert_weights_to_lp touches (Linear / Conv weights and biases, the attention in_proj tensors, `proj`,
    `text_projection`: model/model.py:265-286) in fp16, LayerNorm and embeddings in fp32, plus the three metadata
    scalars that build_model_from_openai_state_dict drops (model/model.py:362-363).  For the checkpoint-ingestion tests."""
    sd = dict(clip_state_dict(cfg, seed))
    lp = ("conv1.weight", "in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias", "c_fc.weight", "c_fc.bias",
          "c_proj.weight", "c_proj.bias")
    for k in list(sd):
        if k.endswith(lp) or k in ("visual.proj", "text_projection"):
            sd[k] = sd[k].half()
    sd["input_resolution"] = torch.tensor(cfg.image_size)
    sd["context_length"] = torch.tensor(cfg.t_context)
    sd["vocab_size"] = torch.tensor(cfg.t_vocab)
    return sd


OPENAI_TINY = dict(image_size=56, patch_size=14, width=128, layers=2, heads=2, embed_dim=32, t_context=77, t_vocab=128,
                   t_width=64, t_heads=1, t_layers=2)
