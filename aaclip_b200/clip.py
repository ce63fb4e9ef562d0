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

    def forward(self, x, attn_mask=None):
        """The reference block returns (x, head-averaged attention weights) and is called per layer by AdaptedCLIP.forward
        (model/transformer.py:239-258, model/adapter.py:91).  Here the per-layer loop lives inside the engine (the
        [L, L] attention matrix is never materialised: flash-style kernel), so a block of this container cannot be
        called on its own - pass the reference's own CLIP to AdaptedCLIP if module-level calls are needed elsewhere."""
        raise NotImplementedError(
            "aaclip_b200.clip._Block holds parameters only: the block arithmetic runs inside the CUDA engine through "
            "AdaptedCLIP.forward / encode_text (no per-block entry, no attention-weight output)")


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
        self.dpam_layer: Optional[int] = None

    @torch.no_grad()
    def DAPM_replace(self, DPAM_layer):
        """model/transformer.py:406-425 (train.py:243): the last DPAM_layer - 1 blocks run the v-v `Attention` with their
        own in_proj / out_proj weights.  The container keeps its parameter names; the engine switches kernels
        (`CLIP.encode_image` below, aaclip_b200/surgery.py)."""
        if DPAM_layer is not None and DPAM_layer - 1 > len(self.transformer.resblocks):
            raise IndexError(f"DAPM_replace({DPAM_layer}): the tower has {len(self.transformer.resblocks)} blocks")
        self.dpam_layer = DPAM_layer


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

    def encode_text(self, text, normalize: bool = False, max_text: int = 256):
        """The un-adapted `CLIP.encode_text` (model/model.py:190-200): token + positional embedding, the causal text
        tower, ln_final, EOT row `@ text_projection`; int tokens [n, context] -> fp32 [n, embed_dim].  test.py:197-200 builds
        the class anchors from it when no text adapter is used (`get_adapted_text_embedding(clip_model, ...)`).  Runs on a
        text-only engine context (no adapters, plain final projection); parameters are re-uploaded when they change."""
        import torch.nn.functional as F
        from .engine import Engine
        if not hasattr(self, "token_embedding"):
            raise RuntimeError("this CLIP was built without a text tower (text=False)")
        if not text.is_cuda:
            raise RuntimeError("aaclip_b200.CLIP.encode_text runs on a B200 only (no CPU / PyTorch fallback)")
        tw = self.token_embedding.weight.shape[1]
        if tuple(self.text_projection.shape) != (tw, tw):
            raise ValueError(f"text_projection {tuple(self.text_projection.shape)}: only square projections are supported")
        dev = text.device.index if text.device.index is not None else torch.cuda.current_device()
        st = self.__dict__.setdefault("_text_engine", {})
        if st.get("engine") is None or st["engine"].device != dev:
            blocks = self.transformer.resblocks
            cfg = ModelCfg(image_size=14, patch_size=14, width=256, layers=1, heads=4, embed_dim=256, levels=[1],
                           image_adapt_until=0, text_adapt_until=0, relu=False,
                           quick_gelu=type(blocks[0].mlp.gelu).__name__ == "QuickGELU",
                           t_context=self.positional_embedding.shape[0], t_vocab=self.token_embedding.weight.shape[0],
                           t_width=tw, t_heads=blocks[0].attn.num_heads, t_layers=len(blocks))
            st["engine"] = Engine(cfg, device=dev, max_batch=1, max_text=max_text, text=True)
            st["engine"].set_text_final(leaky=False)
            st["versions"] = {}
        eng, versions = st["engine"], st["versions"]
        dirty = False
        for k, t in self.state_dict(keep_vars=True).items():
            if k.startswith("visual."):
                continue
            key = "text_adapter.0.fc.0.weight" if k == "text_projection" else "clip." + k
            if key not in eng._wmap:
                continue
            sig = (t.data_ptr(), t._version)
            if versions.get(key) != sig:
                eng.set_weight(key, t.detach().float().t().contiguous() if k == "text_projection" else t)
                versions[key] = sig
                dirty = True
        if dirty:
            torch.cuda.synchronize(dev)
        out = eng.text_forward(text)
        return F.normalize(out, dim=-1) if normalize else out

    def encode_image(self, image, out_layers, normalize: bool = False, max_batch: int = 8):
        """model/model.py:185-188: (pooled [B, E], [tokens [B, L, width] after each block in out_layers]) on the CUDA
        engine - what train.py:75-76 calls on the surgery model and on the un-modified one.  One engine context per
        distinct set of out_layers (each holds the tower's weights); `visual.DAPM_replace` is honoured."""
        from .adapter import effective_levels
        from .surgery import CLIPImageEncoder
        key = tuple(effective_levels(out_layers, len(self.visual.transformer.resblocks)))
        encoders = self.__dict__.setdefault("_image_encoders", {})
        enc = encoders.get(key)
        if enc is None or enc.max_batch < image.shape[0]:
            if enc is not None and enc._engine is not None:
                enc._engine.close()
            enc = CLIPImageEncoder(self, list(key), surgery_until_layer=self.visual.dpam_layer,
                                   max_batch=max(max_batch, int(image.shape[0])))
            encoders[key] = enc
        if enc.surgery_until_layer != self.visual.dpam_layer:
            enc.DAPM_replace(self.visual.dpam_layer)
        return enc.encode_image(image.float().contiguous(), None, normalize)


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


# ---------------------------------------------------------------------------------------------- OpenAI checkpoints
def _unwrap_openai(state_dict) -> dict:
    """model/openai.py:58-72: a checkpoint is either the JIT archive's state dict itself or {"state_dict": {"module.x": ...}}."""
    sd = state_dict
    if isinstance(sd, dict) and "state_dict" in sd and "visual.conv1.weight" not in sd:
        sd = {k[7:] if k.startswith("module.") else k: v for k, v in sd["state_dict"].items()}
    return dict(sd)


def cfg_from_openai_state_dict(state_dict, img_size: Optional[int] = None, quick_gelu: bool = False) -> ModelCfg:
    """Architecture of an OpenAI CLIP ViT state dict from its tensor shapes, as model/model.py:317-343 infers it
    (vision width / layers / patch size / grid from conv1, the in_proj weights and the positional embedding; text
    width, heads = width // 64, layers, context, vocabulary).  `img_size` overrides the checkpoint's resolution (the
    reference builds CLIP(**json, image_size=img_size) and resamples the positional embedding: model/clip.py:112-131);
    `quick_gelu` stays False as in the shipped ViT-L-14-336.json (the model the reference actually runs uses nn.GELU,
    model/model.py:84, whatever activation the checkpoint was trained with)."""
    sd = _unwrap_openai(state_dict)
    if "visual.proj" not in sd:
        raise ValueError("not a ViT CLIP state dict (visual.proj missing): ResNet towers are outside the hot path")
    width = sd["visual.conv1.weight"].shape[0]
    layers = len([k for k in sd if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
    patch = sd["visual.conv1.weight"].shape[-1]
    grid = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    if grid * grid + 1 != sd["visual.positional_embedding"].shape[0]:
        raise ValueError("visual.positional_embedding is not 1 + a square grid")
    mlp = sd["visual.transformer.resblocks.0.mlp.c_fc.weight"].shape[0]
    t_width = sd["ln_final.weight"].shape[0]
    t_layers = len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks")})
    return ModelCfg(image_size=int(img_size) if img_size else patch * grid, patch_size=patch, width=width, layers=layers,
                    heads=max(1, width // 64), mlp_ratio=mlp / width, embed_dim=sd["text_projection"].shape[1],
                    quick_gelu=quick_gelu, t_context=sd["positional_embedding"].shape[0],
                    t_vocab=sd["token_embedding.weight"].shape[0], t_width=t_width, t_heads=max(1, t_width // 64),
                    t_layers=t_layers)


def load_openai_state_dict(state_dict, img_size: Optional[int] = None, quick_gelu: bool = False) -> CLIP:
    """An OpenAI CLIP ViT checkpoint (the JIT archive's state dict: fp16 Linear / Conv / attention / projection
    tensors, fp32 LayerNorm and embeddings, plus the three metadata scalars) -> a loaded `CLIP` container, the way the
    reference's create_model(pretrained="openai") gets there:

        load_openai_model -> build_model_from_openai_state_dict -> .float()      model/openai.py:66-77, model/model.py:311-368
        state_dict = model_pre.state_dict(); CLIP(**json, image_size=img_size)   model/clip.py:112-126
        resize_pos_embed(state_dict, model); load_state_dict(strict=True)        model/clip.py:130-131, model/model.py:395-426

    i.e. metadata keys dropped, every tensor widened to fp32 (exact), the positional embedding resampled (bicubic,
    antialias, align_corners=False) when `img_size` differs from the checkpoint's, strict load.  Wrap the result in
    AdaptedCLIP: the engine packs GEMM operands to bf16 from these fp32 values."""
    sd = _unwrap_openai(state_dict)
    cfg = cfg_from_openai_state_dict(sd, img_size, quick_gelu)
    for k in ("input_resolution", "context_length", "vocab_size"):   # model/model.py:362-363
        sd.pop(k, None)
    sd = {k: (v.float() if torch.is_floating_point(v) else v) for k, v in sd.items()}
    model = CLIP(cfg, text=True)
    resize_pos_embed(sd, model.visual.grid_size[0])
    model.load_state_dict(sd, strict=True)
    return model.eval()
