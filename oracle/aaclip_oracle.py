"""ORACLE — test infrastructure only.  NOT part of the product path.

A CPU, fp32, plain-PyTorch restatement of the reference (wei-paul/AA-CLIP) inference hot path, written
against state dicts so that it runs where /root/reference does not exist (the GPU box).  Only tests/,
__graft_entry__.smoke() and bench.py's baseline legs (cpu_baseline, --impl reference, and torch_gpu_baseline - the
same op sequence on the B200 as the like-for-like PyTorch baseline of SURVEY 8(d)) may import this module; the
package aaclip_b200 never does.

Pinning: the reference ships no tests and no golden vectors, so the restatement is pinned
against the REAL reference modules imported from /root/reference by oracle/make_golden.py, which also
writes tests/golden/*.pt; tests/test_oracle_golden.py re-checks the oracle against those files wherever
it runs.  One boundary stays unpinned: `gaussian_blur2d` lives in the un-vendored third-party dependency
kornia==0.6.9 (requirements.txt:3, call site forward_utils.py:208-210), absent offline — it is restated
below from kornia 0.6.9's published algorithm and marked "parity unpinned (kornia)".

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------- blocks
def layer_norm(x: torch.Tensor, sd: SD, prefix: str, eps: float = 1e-5) -> torch.Tensor:
    """model/transformer.py:37-43 (LayerNorm.forward -> F.layer_norm, eps default 1e-5)."""
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + "weight"], sd[prefix + "bias"], eps)


def activation(x: torch.Tensor, quick_gelu: bool) -> torch.Tensor:
    """nn.GELU (exact erf) unless quick_gelu (model/model.py:84); QuickGELU model/transformer.py:46-49."""
    return x * torch.sigmoid(1.702 * x) if quick_gelu else F.gelu(x)


def multi_head_attention(x: torch.Tensor, sd: SD, prefix: str, heads: int,
                         attn_mask: Optional[torch.Tensor]) -> torch.Tensor:
    """nn.MultiheadAttention(q=k=v=x) as called at model/transformer.py:226-237: packed in-proj, q scaled by
    1/sqrt(head_dim), additive mask, softmax, out-proj.  x is [B, L, D] (batch-first restatement of the
    reference's LND layout; the arithmetic per token is identical)."""
    B, L, D = x.shape
    hd = D // heads
    qkv = F.linear(x, sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"])
    q, k, v = qkv.view(B, L, 3, heads, hd).permute(2, 0, 3, 1, 4)  # each [B, h, L, hd]
    s = (q * (1.0 / math.sqrt(hd))) @ k.transpose(-1, -2)
    if attn_mask is not None:
        s = s + attn_mask
    p = torch.softmax(s, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B, L, D)
    return F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def residual_attention_block(x: torch.Tensor, sd: SD, prefix: str, heads: int, quick_gelu: bool,
                             attn_mask: Optional[torch.Tensor]) -> torch.Tensor:
    """model/transformer.py:239-258 with ls_1 = ls_2 = Identity (:201-205, :220-224)."""
    x = x + multi_head_attention(layer_norm(x, sd, prefix + "ln_1."), sd, prefix + "attn.", heads, attn_mask)
    h = F.linear(layer_norm(x, sd, prefix + "ln_2."), sd[prefix + "mlp.c_fc.weight"], sd[prefix + "mlp.c_fc.bias"])
    h = activation(h, quick_gelu)
    return x + F.linear(h, sd[prefix + "mlp.c_proj.weight"], sd[prefix + "mlp.c_proj.bias"])


def adapter_mix(x: torch.Tensor, w_adapter: torch.Tensor, weight: float) -> torch.Tensor:
    """model/adapter.py:92-99: SimpleAdapter = Linear(bias=False) + LeakyReLU (model/adapter_modules.py:6-13),
    norm matched to x over the last dim (no epsilon), then i_w * a + (1 - i_w) * x."""
    a = F.leaky_relu(F.linear(x, w_adapter), 0.01)
    a = a * x.norm(dim=-1, keepdim=True) / a.norm(dim=-1, keepdim=True)
    return weight * a + (1 - weight) * x


def simple_proj(x: torch.Tensor, sd: SD, prefix: str) -> torch.Tensor:
    """model/adapter_modules.py:16-26: key `fc.0.weight` (Linear + LeakyReLU) iff relu else `fc.weight`."""
    if prefix + "fc.0.weight" in sd:
        return F.leaky_relu(F.linear(x, sd[prefix + "fc.0.weight"]), 0.01)
    return F.linear(x, sd[prefix + "fc.weight"])


# ----------------------------------------------------------------------------------------------- visual
def visual_forward(clip_sd: SD, image_adapter_sd: SD, image: torch.Tensor, *, patch_size: int = 14, heads: int = 16,
                   layers: int = 24, image_adapt_until: int = 6, image_adapt_weight: float = 0.1,
                   levels: Sequence[int] = (6, 12, 18, 24), quick_gelu: bool = False,
                   return_taps: bool = False):
    """AdaptedCLIP.forward, model/adapter.py:67-112.  Returns (seg_tokens: list of [B,P,E], det_token [B,E])."""
    sd = clip_sd
    x = F.conv2d(image, sd["visual.conv1.weight"], stride=patch_size)          # :68
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)                 # :69-70
    cls = sd["visual.class_embedding"] + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device)
    x = torch.cat([cls, x], dim=1)                                             # :72-81
    x = x + sd["visual.positional_embedding"]                                  # :82
    x = layer_norm(x, sd, "visual.ln_pre.")                                    # :84-85 (patch_dropout = identity)
    tokens = []
    for i in range(layers):                                                    # :90
        x = residual_attention_block(x, sd, f"visual.transformer.resblocks.{i}.", heads, quick_gelu, None)
        if i < image_adapt_until:                                              # :92-99
            x = adapter_mix(x, image_adapter_sd[f"layer_adapters.{i}.fc.0.weight"], image_adapt_weight)
        if i + 1 in levels:                                                    # :100-101
            tokens.append(x[:, 1:, :])
    taps = tokens
    tokens = [layer_norm(t, sd, "visual.ln_post.") for t in tokens]            # :105
    seg = [simple_proj(t, image_adapter_sd, f"seg_proj.{i}.") for i, t in enumerate(tokens)]   # :106-108
    seg = [F.normalize(t, dim=-1) for t in seg]                                # :109
    det = simple_proj(tokens[-1], image_adapter_sd, "det_proj.")               # :110
    det = F.normalize(det, dim=-1).mean(1)                                     # :111
    if return_taps:
        return seg, det, taps
    return seg, det


# ----------------------------------------------------------------------------------------------- surgery (train.py stage 1)
def vv_attention(x: torch.Tensor, sd: SD, prefix: str, heads: int) -> torch.Tensor:
    """`Attention.forward` (model/transformer.py:123-152) as installed by `DAPM_replace` (:406-425) and called from
    `ResidualAttentionBlock.attention` (:226-237) with the block's [L, batch, D] tensor.  The module reads the shape
    as (B, N, C) = (L, batch, D): its softmax therefore runs over the IMAGES OF THE BATCH for every token position
    and head.  q and k only feed `attn_ori` / `x_ori`, which are discarded (:133-136, :147, :150); the returned x is
    proj(softmax(v v^T * scale) v).  x here is [batch, L, D] (batch-first restatement); returns [batch, L, D]."""
    B, L, D = x.shape
    hd = D // heads
    qkv = F.linear(x, sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"])     # self.qkv(q_x), weights :412-417
    v = qkv[..., 2 * D:].reshape(B, L, heads, hd).permute(1, 2, 0, 3)                  # [L, heads, batch, hd]
    attn = torch.softmax((v @ v.transpose(-2, -1)) * hd ** -0.5, dim=-1)              # :142-145, [L, heads, batch, batch]
    o = (attn @ v).permute(2, 0, 1, 3).reshape(B, L, D)                                # :148
    return F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])  # :149, weights :418-423


def encode_image(clip_sd: SD, image: torch.Tensor, out_layers: Sequence[int], *, patch_size: int = 14, heads: int = 16,
                 layers: int = 24, surgery_until_layer: Optional[int] = None, quick_gelu: bool = False,
                 normalize: bool = False):
    """`CLIP.encode_image(image, out_layers, normalize)` (model/model.py:185-188) -> `VisionTransformer.forward`
    (model/transformer.py:490-551) -> `Transformer.forward` (:296-318).  `surgery_until_layer` = the DPAM_layer given
    to `DAPM_replace` (:406-425; train.py:243): blocks[-i] for i in 1..DPAM_layer-1 use `vv_attention`.
    Returns (pooled [B, E], [tokens [B, L, D] after each block in out_layers])."""
    sd = clip_sd
    x = F.conv2d(image, sd["visual.conv1.weight"], stride=patch_size)          # :507-509
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
    cls = sd["visual.class_embedding"] + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device)
    x = torch.cat([cls, x], dim=1)                                             # :512-521
    x = x + sd["visual.positional_embedding"]                                  # :522
    x = layer_norm(x, sd, "visual.ln_pre.")                                    # :525-526 (patch_dropout: identity in eval)
    n_vv = max(int(surgery_until_layer) - 1, 0) if surgery_until_layer is not None else 0
    tokens = []
    for i in range(layers):                                                    # Transformer.forward :304-316
        prefix = f"visual.transformer.resblocks.{i}."
        if i >= layers - n_vv:
            # ResidualAttentionBlock.forward :239-258 with self.attn = Attention
            x = x + vv_attention(layer_norm(x, sd, prefix + "ln_1."), sd, prefix + "attn.", heads)
            h = F.linear(layer_norm(x, sd, prefix + "ln_2."), sd[prefix + "mlp.c_fc.weight"], sd[prefix + "mlp.c_fc.bias"])
            x = x + F.linear(activation(h, quick_gelu), sd[prefix + "mlp.c_proj.weight"], sd[prefix + "mlp.c_proj.bias"])
        else:
            x = residual_attention_block(x, sd, prefix, heads, quick_gelu, None)
        if i + 1 in out_layers:
            tokens.append(x)
    pooled = layer_norm(x[:, 0], sd, "visual.ln_post.") @ sd["visual.proj"]    # :542-546 (_global_pool :484-488: class token)
    if normalize:
        pooled = F.normalize(pooled, dim=-1)
    return pooled, tokens


def surgery_patch_features(surgery_sd: SD, clip_sd: SD, image: torch.Tensor, *, levels: Sequence[int] = (6, 12, 18, 24),
                           surgery_until_layer: int = 20, patch_size: int = 14, heads: int = 16, layers: int = 24,
                           quick_gelu: bool = False) -> List[torch.Tensor]:
    """train.py:74-85, the frozen stage-1 feature extractor: patch tokens of the surgery CLIP at `levels` -> ln_post ->
    @ visual.proj -> / norm, plus the (normalised) pooled class feature of the unmodified CLIP.  -> list of [B, P, E]."""
    kw = dict(patch_size=patch_size, heads=heads, layers=layers, quick_gelu=quick_gelu)
    _, toks = encode_image(surgery_sd, image, levels, surgery_until_layer=surgery_until_layer, **kw)     # :75
    cls, _ = encode_image(clip_sd, image, [], **kw)                                                      # :76
    cls = cls / cls.norm(dim=-1, keepdim=True)                                                           # :77
    feats = [layer_norm(t[:, 1:, :], surgery_sd, "visual.ln_post.") for t in toks]                       # :78-80
    feats = [t @ surgery_sd["visual.proj"] for t in feats]                                               # :81
    feats = [t / t.norm(dim=-1, keepdim=True) for t in feats]                                            # :82-84
    return [t + cls.unsqueeze(1) for t in feats]                                                         # :85


# ----------------------------------------------------------------------------------------------- text
def causal_mask(n: int) -> torch.Tensor:
    """CLIP.attn_mask (model/model.py:172; build_attention_mask transformer.py:629-635)."""
    return torch.full((n, n), float("-inf")).triu_(1)


def encode_text(clip_sd: SD, text_adapter_sd: SD, tokens: torch.Tensor, *, heads: int = 12, layers: int = 12,
                text_adapt_until: int = 3, text_adapt_weight: float = 0.1, quick_gelu: bool = False) -> torch.Tensor:
    """AdaptedCLIP.encode_text(adapt_text=True), model/adapter.py:114-145.  tokens int [n, ctx] -> [n, width]."""
    sd = clip_sd
    x = F.embedding(tokens.long(), sd["token_embedding.weight"])               # :118
    x = x + sd["positional_embedding"]                                         # :122
    mask = causal_mask(tokens.shape[1]).to(x.device)
    for i in range(layers):                                                    # :125-136
        x = residual_attention_block(x, sd, f"transformer.resblocks.{i}.", heads, quick_gelu, mask)
        if i < text_adapt_until:
            x = adapter_mix(x, text_adapter_sd[f"{i}.fc.0.weight"], text_adapt_weight)
    x = layer_norm(x, sd, "ln_final.")                                         # :138
    eot = x[torch.arange(x.shape[0], device=x.device), tokens.argmax(dim=-1)]                   # :140
    return F.leaky_relu(F.linear(eot, text_adapter_sd[f"{text_adapt_until}.fc.0.weight"]), 0.01)


def clip_encode_text(clip_sd: SD, tokens: torch.Tensor, *, heads: int = 12, layers: int = 12, quick_gelu: bool = False,
                     normalize: bool = False) -> torch.Tensor:
    """The un-adapted `CLIP.encode_text` (model/model.py:189-200; what `AdaptedCLIP.encode_text(adapt_text=False)` delegates to,
    model/adapter.py:115-116, and what test.py:197-200 builds anchors from): tokens int [n, ctx] -> [n, embed_dim]."""
    sd = clip_sd
    x = F.embedding(tokens.long(), sd["token_embedding.weight"])               # :191
    x = x + sd["positional_embedding"]                                         # :193
    mask = causal_mask(tokens.shape[1]).to(x.device)
    for i in range(layers):                                                    # :195
        x = residual_attention_block(x, sd, f"transformer.resblocks.{i}.", heads, quick_gelu, mask)
    x = layer_norm(x, sd, "ln_final.")                                         # :197
    x = x[torch.arange(x.shape[0], device=x.device), tokens.argmax(dim=-1)] @ sd["text_projection"]   # :199
    return F.normalize(x, dim=-1) if normalize else x                          # :200


def class_text_anchor(emb_normal: torch.Tensor, emb_abnormal: torch.Tensor) -> torch.Tensor:
    """forward_utils.py:155-161: per-sentence L2 norm -> mean -> L2 norm, stacked to [width, 2]."""
    cols = []
    for e in (emb_normal, emb_abnormal):
        e = e / e.norm(dim=-1, keepdim=True)
        m = e.mean(dim=0)
        cols.append(m / m.norm())
    return torch.stack(cols, dim=1)


# ----------------------------------------------------------------------------------------------- head
def gaussian_kernel1d(ksize: int, sigma: float) -> torch.Tensor:
    """kornia 0.6.9 kornia/filters/kernels.py::gaussian: x = arange(k) - k//2 (+0.5 if k even),
    exp(-x^2 / (2 sigma^2)), normalised to sum 1.   [parity unpinned (kornia)]"""
    x = torch.arange(ksize, dtype=torch.float32) - ksize // 2
    if ksize % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * sigma ** 2))
    return g / g.sum()


def gaussian_blur2d(x: torch.Tensor, kernel_size: Tuple[int, int], sigma: Tuple[float, float]) -> torch.Tensor:
    """kornia 0.6.9 gaussian_blur2d(input, kernel_size, sigma, border_type='reflect', separable=True):
    reflect-pad by k//2 and correlate every channel with the (separable) normalised Gaussian.
    [parity unpinned (kornia): un-vendored dependency, restated from its published algorithm]"""
    ky, kx = kernel_size
    gy, gx = gaussian_kernel1d(ky, sigma[1]), gaussian_kernel1d(kx, sigma[0])
    b, c, h, w = x.shape
    xp = F.pad(x, (kx // 2, kx // 2, ky // 2, ky // 2), mode="reflect")
    k2d = torch.outer(gy, gx).to(device=x.device, dtype=x.dtype)[None, None].expand(c, 1, ky, kx)
    return F.conv2d(xp, k2d, groups=c)


def calculate_similarity_map(patch_features: torch.Tensor, text_feature: torch.Tensor, img_size: int,
                             test: bool = False, domain: str = "Medical") -> torch.Tensor:
    """forward_utils.py:196-216, line for line."""
    scores = 100.0 * torch.matmul(patch_features, text_feature)                # :199
    B, L, C = scores.shape
    H = int(math.sqrt(L))                                                      # :201
    pred = scores.permute(0, 2, 1).reshape(B, C, H, H)                         # :202
    if test:
        assert C == 2                                                          # :204
        sigma = 1 if domain == "Industrial" else 1.5
        ksize = 7 if domain == "Industrial" else 9
        pred = (pred[:, 1] + 1 - pred[:, 0]) / 2                               # :207
        pred = gaussian_blur2d(pred.unsqueeze(1), (ksize, ksize), (sigma, sigma))   # :208-210
    out = F.interpolate(pred, size=img_size, mode="bilinear", align_corners=True)   # :211-213
    if not test and C > 1:
        out = torch.softmax(out, dim=1)                                        # :214-215
    return out


def predict(seg_tokens: List[torch.Tensor], det_token: torch.Tensor, text_feature: torch.Tensor, img_size: int,
            domain: str = "Industrial") -> Tuple[torch.Tensor, torch.Tensor]:
    """Body of test.py:get_predictions for one batch (test.py:83-93): image score and level-summed map."""
    pred = det_token @ text_feature
    score = (pred[:, 1] + 1) / 2                                               # test.py:83-84
    maps = [calculate_similarity_map(f, text_feature, img_size, test=True, domain=domain) for f in seg_tokens]
    return torch.cat(maps, dim=1).sum(1), score                                # test.py:93


def minmax_normalise(maps: torch.Tensor) -> torch.Tensor:
    """forward_utils.py:241-244 (metrics_eval): (x - min) / (max - min) over the whole set."""
    lo, hi = maps.min(), maps.max()
    return (maps - lo) / (hi - lo)


def metrics_image_preds(pixel_preds, image_preds, domain: str = "Industrial"):
    """forward_utils.py:241-254 (the part of metrics_eval in front of the sklearn calls), numpy like the reference:
    global min-max normalisation of the pixel and image predictions unless their max is exactly 1, per-image pixel
    maximum, 50/50 mix (Medical: pixel maximum alone).  Returns (normalised pixel_preds, image_preds)."""
    import numpy as np
    pixel_preds = np.asarray(pixel_preds)
    image_preds = np.asarray(image_preds)
    if pixel_preds.max() != 1:
        pixel_preds = (pixel_preds - pixel_preds.min()) / (pixel_preds.max() - pixel_preds.min())
    if image_preds.max() != 1:
        image_preds = (image_preds - image_preds.min()) / (image_preds.max() - image_preds.min())
    pmax_pred = pixel_preds.max(axis=(1, 2))
    if domain != "Medical":
        image_preds = pmax_pred * 0.5 + image_preds * 0.5
    else:
        image_preds = pmax_pred
    return pixel_preds, image_preds
