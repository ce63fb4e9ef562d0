"""Torch-tensor front ends of the building-block entry points of the C ABI.

These only validate shapes/dtypes, allocate outputs with torch (device memory + current stream are the
plumbing torch provides) and pass raw pointers to libaaclip_b200.so.  No arithmetic happens in Python and
nothing here falls back to a torch implementation.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import (ACT_GELU_ERF, ACT_LEAKY, ACT_NONE, ACT_QUICK_GELU, HEAD_TEST_INDUSTRIAL, HEAD_TEST_MEDICAL,
                   HEAD_TRAIN_SOFTMAX, OUT_BF16, OUT_F32, OUT_F32_PATCH, OUT_F32_RESID, check, cur_stream, ptr)

__all__ = ["gemm", "layernorm", "attention", "adapter_mix", "anomaly_head", "preprocess_u8", "resize_bicubic_u8", "map_minmax", "gemm_resid_ln", "gemm_lnfold", "rowstats_cast", "fold_ln_weight",
           "ACT_NONE", "ACT_GELU_ERF", "ACT_QUICK_GELU", "ACT_LEAKY",
           "OUT_BF16", "OUT_F32", "OUT_F32_RESID", "OUT_F32_PATCH",
           "HEAD_TEST_INDUSTRIAL", "HEAD_TEST_MEDICAL", "HEAD_TRAIN_SOFTMAX"]


def _need(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (aaclip_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         out_mode: int = OUT_F32, out: Optional[torch.Tensor] = None, pos: Optional[torch.Tensor] = None,
         patches: int = 0, cta_group: int = 2) -> torch.Tensor:
    """out = epilogue(a[M,K] @ w[N,K]^T) on the tcgen05 GEMM.  a, w bf16; bias/pos fp32."""
    _need(a, torch.bfloat16, "a"); _need(w, torch.bfloat16, "w")
    M, K = a.shape
    N, K2 = w.shape
    if K != K2:
        raise ValueError(f"K mismatch: a {tuple(a.shape)} vs w {tuple(w.shape)}")
    if bias is not None:
        _need(bias, torch.float32, "bias")
    if out_mode == OUT_F32_PATCH:
        if pos is None or patches <= 0 or out is None:
            raise ValueError("OUT_F32_PATCH needs pos, patches and a preallocated out")
        _need(pos, torch.float32, "pos")
    if out is None:
        if out_mode == OUT_F32_RESID:
            raise ValueError("OUT_F32_RESID accumulates into `out`")
        out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16 if out_mode == OUT_BF16 else torch.float32)
    _need(out, torch.bfloat16 if out_mode == OUT_BF16 else torch.float32, "out")
    lib = _lib.load()
    check(lib.aaclip_gemm_bf16(ptr(a), K, ptr(w), K, M, N, K, ptr(bias), ptr(out), out.shape[-1], act, out_mode,
                               ptr(pos), patches, cta_group, cur_stream(a.device)))
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out_bf16: bool = True, out_f32: bool = False):
    _need(x, torch.float32, "x"); _need(gamma, torch.float32, "gamma"); _need(beta, torch.float32, "beta")
    rows, width = x.shape
    ob = torch.empty(rows, width, device=x.device, dtype=torch.bfloat16) if out_bf16 else None
    of = torch.empty_like(x) if out_f32 else None
    check(_lib.load().aaclip_layernorm(ptr(x), ptr(gamma), ptr(beta), eps, rows, width, ptr(ob), ptr(of),
                                       cur_stream(x.device)))
    return ob, of


def attention(qkv: torch.Tensor, B: int, L: int, heads: int, causal: bool = False) -> torch.Tensor:
    """qkv bf16 [B*L, 3*heads*64] -> bf16 [B*L, heads*64] = softmax(q k^T / 8 [+ causal mask]) v per head."""
    _need(qkv, torch.bfloat16, "qkv")
    if tuple(qkv.shape) != (B * L, 3 * heads * 64):
        raise ValueError(f"qkv shape {tuple(qkv.shape)} != ({B * L}, {3 * heads * 64})")
    out = torch.empty(B * L, heads * 64, device=qkv.device, dtype=torch.bfloat16)
    check(_lib.load().aaclip_attention(ptr(qkv), ptr(out), B, L, heads, int(causal), cur_stream(qkv.device)))
    return out


def adapter_mix(x: torch.Tensor, a: torch.Tensor, w: float) -> torch.Tensor:
    """In place: x <- w * a * |x| / |a| + (1 - w) * x, norms over the last dim."""
    _need(x, torch.float32, "x"); _need(a, torch.float32, "a")
    rows, width = x.shape
    check(_lib.load().aaclip_adapter_mix(ptr(x), ptr(a), w, rows, width, cur_stream(x.device)))
    return x


def anomaly_head(seg: Sequence[torch.Tensor], anchors: torch.Tensor, img_size: int, mode: int,
                 det: Optional[torch.Tensor] = None, want_maps: bool = True, want_extrema: bool = False):
    """seg: list of [B,P,E] (fp32 or bf16) normalised patch tokens; anchors fp32 [E,2] or [B,E,2].

    Returns (maps, scores): test modes maps fp32 [B,S,S] summed over levels; train mode [n_levels,B,2,S,S];
    scores fp32 [B] if det is given, else None.  With `want_extrema` (test modes, shared anchors) a third value
    fp32 [B,2] = per-image (min, max) of the maps, from the same kernel that writes them.
    """
    n = len(seg)
    if n == 0:
        raise ValueError("anomaly_head: no levels")
    B, P, E = seg[0].shape
    is_bf16 = seg[0].dtype == torch.bfloat16
    for t in seg:
        _need(t, torch.bfloat16 if is_bf16 else torch.float32, "seg level")
        if tuple(t.shape) != (B, P, E):
            raise ValueError("all levels must share one shape")
    _need(anchors, torch.float32, "anchors")
    batched = anchors.dim() == 3
    if tuple(anchors.shape) != ((B, E, 2) if batched else (E, 2)):
        raise ValueError(f"anchors shape {tuple(anchors.shape)}")
    dev = seg[0].device
    # the kernels read tokens with 16-byte vector / bulk loads: a view at an odd offset into a larger buffer is re-packed
    seg = [t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format) for t in seg]
    maps = None
    if want_maps:
        shape = (n, B, 2, img_size, img_size) if mode == HEAD_TRAIN_SOFTMAX else (B, img_size, img_size)
        maps = torch.empty(shape, device=dev, dtype=torch.float32)
    scores = None
    if det is not None:
        _need(det, torch.float32, "det")
        scores = torch.empty(B, device=dev, dtype=torch.float32)
    extrema = torch.empty(B, 2, device=dev, dtype=torch.float32) if want_extrema else None
    lib = _lib.load()
    # scratch comes from torch's caching allocator (stream-ordered, capturable); the library allocates nothing
    ws_bytes = int(lib.aaclip_anomaly_head_workspace_bytes(n, B, P)) if want_maps else 0
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8) if ws_bytes else None
    arr = (C.c_void_p * n)(*[t.data_ptr() for t in seg])
    check(lib.aaclip_anomaly_head(arr, n, int(is_bf16), ptr(anchors), int(batched), ptr(det), B, P, E, img_size, mode,
                                  ptr(maps), ptr(scores), ptr(extrema), ptr(ws), ws_bytes, cur_stream(dev)))
    if want_extrema:
        return maps, scores, extrema
    return maps, scores


def _check_u8(images: torch.Tensor) -> None:
    _need(images, torch.uint8, "images")
    if images.dim() != 4 or images.shape[3] != 3:
        raise ValueError(f"images must be uint8 [B,H0,W0,3], got {tuple(images.shape)}")


def preprocess_u8(images: torch.Tensor, size: int, mean=None, std=None) -> torch.Tensor:
    """dataset/__init__.py:127-136 on the device: uint8 [B,H0,W0,3] RGB -> PIL-bicubic resize to size x size ->
    /255 -> (x - mean) / std -> float32 [B,3,size,size], bit-exact with PIL + torchvision."""
    _check_u8(images)
    B, H0, W0, _ = images.shape
    lib = _lib.load()
    nbytes = int(lib.aaclip_preprocess_scratch_bytes(B, H0, W0, size))
    scratch = torch.empty(max(nbytes, 1), device=images.device, dtype=torch.uint8)
    out = torch.empty(B, 3, size, size, device=images.device, dtype=torch.float32)
    m = (C.c_float * 3)(*mean) if mean is not None else None
    s = (C.c_float * 3)(*std) if std is not None else None
    check(lib.aaclip_preprocess_u8(ptr(images), B, H0, W0, size, m, s, ptr(scratch), ptr(out), cur_stream(images.device)))
    return out


def resize_bicubic_u8(images: torch.Tensor, size: int) -> torch.Tensor:
    """PIL.Image.resize((size, size), Image.BICUBIC) on uint8 [B,H0,W0,3] -> uint8 [B,size,size,3], byte-exact."""
    _check_u8(images)
    B, H0, W0, _ = images.shape
    lib = _lib.load()
    nbytes = int(lib.aaclip_preprocess_scratch_bytes(B, H0, W0, size))
    scratch = torch.empty(max(nbytes, 1), device=images.device, dtype=torch.uint8)
    out = torch.empty(B, size, size, 3, device=images.device, dtype=torch.uint8)
    check(lib.aaclip_resize_bicubic_u8(ptr(images), B, H0, W0, size, ptr(scratch), ptr(out), cur_stream(images.device)))
    return out


def map_minmax(maps: torch.Tensor) -> torch.Tensor:
    """maps fp32 [B, ...] -> fp32 [B, 2] = per-image (min, max) over all trailing dims (exact)."""
    _need(maps, torch.float32, "maps")
    B = maps.shape[0]
    n_pix = maps.numel() // B if B else 0
    out = torch.empty(B, 2, device=maps.device, dtype=torch.float32)
    check(_lib.load().aaclip_map_minmax(ptr(maps), B, n_pix, ptr(out), cur_stream(maps.device)))
    return out


# ---- folded-LayerNorm building blocks (see include/aaclip_b200.h) -------------------------------------------------
def fold_ln_weight(w: torch.Tensor, bias: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor):
    """W fp32 [N,K], LayerNorm affine (gamma, beta) -> (Wf bf16 [N,K], colsum fp32 [N], bias_f fp32 [N])."""
    _need(w, torch.float32, "w"); _need(gamma, torch.float32, "gamma"); _need(beta, torch.float32, "beta")
    N, K = w.shape
    wf = torch.empty(N, K, device=w.device, dtype=torch.bfloat16)
    colsum = torch.empty(N, device=w.device, dtype=torch.float32)
    bias_f = torch.empty(N, device=w.device, dtype=torch.float32)
    check(_lib.load().aaclip_fold_ln_weight(ptr(w), ptr(bias), ptr(gamma), ptr(beta), N, K, ptr(wf), ptr(colsum),
                                            ptr(bias_f), cur_stream(w.device)))
    return wf, colsum, bias_f


def rowstats_cast(x: torch.Tensor, slices: int):
    """x fp32 [rows,width] -> (bf16 copy, part fp32 [rows,slices,2] with the whole-row (sum, sum sq) in slice 0)."""
    _need(x, torch.float32, "x")
    rows, width = x.shape
    xb = torch.empty(rows, width, device=x.device, dtype=torch.bfloat16)
    part = torch.empty(rows, slices, 2, device=x.device, dtype=torch.float32)
    check(_lib.load().aaclip_rowstats_cast(ptr(x), rows, width, ptr(xb), ptr(part), slices, cur_stream(x.device)))
    return xb, part


def gemm_resid_ln(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], x: torch.Tensor, cta_group: int = 2):
    """x += a @ w^T + bias in place (fp32); returns (bf16 copy of the new x, part fp32 [M, N/128, 2])."""
    _need(a, torch.bfloat16, "a"); _need(w, torch.bfloat16, "w"); _need(x, torch.float32, "x")
    M, K = a.shape
    N = w.shape[0]
    xb = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    part = torch.empty(M, N // 128, 2, device=a.device, dtype=torch.float32)
    check(_lib.load().aaclip_gemm_resid_ln(ptr(a), K, ptr(w), K, M, N, K, ptr(bias), ptr(x), N, ptr(xb), N, ptr(part),
                                           cta_group, cur_stream(a.device)))
    return xb, part


def gemm_lnfold(xb: torch.Tensor, wf: torch.Tensor, bias_f: torch.Tensor, colsum: torch.Tensor, part: torch.Tensor,
                eps: float = 1e-5, act: int = ACT_NONE, cta_group: int = 2) -> torch.Tensor:
    """bf16 [M,N] = act(LayerNorm(x) @ W^T + b) computed from the bf16 copy xb, the folded weight and the row statistics."""
    _need(xb, torch.bfloat16, "xb"); _need(wf, torch.bfloat16, "wf"); _need(part, torch.float32, "part")
    M, K = xb.shape
    N = wf.shape[0]
    out = torch.empty(M, N, device=xb.device, dtype=torch.bfloat16)
    check(_lib.load().aaclip_gemm_lnfold(ptr(xb), K, ptr(wf), K, M, N, K, ptr(bias_f), ptr(colsum), ptr(part),
                                         part.shape[1], eps, ptr(out), N, act, cta_group, cur_stream(xb.device)))
    return out


def vv_attention(v: torch.Tensor, B: int, L: int, heads: int) -> torch.Tensor:
    """The surgery extractor's v-v attention (model/transformer.py:123-152 under DAPM_replace): v bf16 [B*L, heads*64]
    (row = b*L + l) -> bf16 [B*L, heads*64]; for every (token l, head) the B images attend to each other:
    out[b] = sum_b' softmax_b'(<v[b], v[b']> / 8) v[b'].  Batch-coupled by construction; B <= 128."""
    _need(v, torch.bfloat16, "v")
    if v.dim() != 2 or v.shape[0] != B * L or v.shape[1] != heads * 64:
        raise ValueError(f"v shape {tuple(v.shape)} != ({B * L}, {heads * 64})")
    out = torch.empty(B * L, heads * 64, device=v.device, dtype=torch.bfloat16)
    check(_lib.load().aaclip_vv_attention(ptr(v), v.stride(0), ptr(out), heads * 64, B, L, heads, cur_stream(v.device)))
    return out


def add_image_vector(tokens: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    """In place: tokens fp32 [B, P, E] += vec fp32 [B, E] broadcast over the patches (train.py:85)."""
    _need(tokens, torch.float32, "tokens"); _need(vec, torch.float32, "vec")
    B, P, E = tokens.shape
    if tuple(vec.shape) != (B, E):
        raise ValueError(f"vec shape {tuple(vec.shape)} != ({B}, {E})")
    check(_lib.load().aaclip_add_image_vector(ptr(tokens), ptr(vec), B, P, E, cur_stream(tokens.device)))
    return tokens
