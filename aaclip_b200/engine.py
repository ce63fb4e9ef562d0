"""Python handle on an `aaclip_ctx` (include/aaclip_b200.h): creation from a ModelCfg, weight upload from
reference-keyed state dicts, and the forward entry points.  Pure plumbing: pointers in, pointers out."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import (ACT_GELU_ERF, ACT_QUICK_GELU, HEAD_TEST_INDUSTRIAL, HEAD_TEST_MEDICAL, W, AaclipCfg, check,
                   cur_stream, ptr)
from .synth import ModelCfg

DOMAIN_MODE = {"Industrial": HEAD_TEST_INDUSTRIAL, "Medical": HEAD_TEST_MEDICAL}


def _block_keys(prefix: str, tower: str):
    t = tower  # "V" or "T"
    return [
        (prefix + "ln_1.weight", W[f"{t}_LN1_G"]), (prefix + "ln_1.bias", W[f"{t}_LN1_B"]),
        (prefix + "attn.in_proj_weight", W[f"{t}_QKV_W"]), (prefix + "attn.in_proj_bias", W[f"{t}_QKV_B"]),
        (prefix + "attn.out_proj.weight", W[f"{t}_OUT_W"]), (prefix + "attn.out_proj.bias", W[f"{t}_OUT_B"]),
        (prefix + "ln_2.weight", W[f"{t}_LN2_G"]), (prefix + "ln_2.bias", W[f"{t}_LN2_B"]),
        (prefix + "mlp.c_fc.weight", W[f"{t}_FC_W"]), (prefix + "mlp.c_fc.bias", W[f"{t}_FC_B"]),
        (prefix + "mlp.c_proj.weight", W[f"{t}_PROJ_W"]), (prefix + "mlp.c_proj.bias", W[f"{t}_PROJ_B"]),
    ]


def weight_map(cfg: ModelCfg, text: bool = True) -> Dict[str, tuple]:
    """state-dict key -> (weight id, layer) for the three reference state dicts, with the prefixes
    'clip.', 'image_adapter.' and 'text_adapter.' telling them apart."""
    m: Dict[str, tuple] = {}
    m["clip.visual.conv1.weight"] = (W["V_CONV1"], 0)
    m["clip.visual.class_embedding"] = (W["V_CLS"], 0)
    m["clip.visual.positional_embedding"] = (W["V_POS"], 0)
    m["clip.visual.ln_pre.weight"] = (W["V_LN_PRE_G"], 0); m["clip.visual.ln_pre.bias"] = (W["V_LN_PRE_B"], 0)
    m["clip.visual.ln_post.weight"] = (W["V_LN_POST_G"], 0); m["clip.visual.ln_post.bias"] = (W["V_LN_POST_B"], 0)
    for i in range(cfg.layers):
        for k, wid in _block_keys(f"visual.transformer.resblocks.{i}.", "V"):
            m["clip." + k] = (wid, i)
    fc = "fc.0.weight" if cfg.relu else "fc.weight"
    for i in range(cfg.image_adapt_until):
        m[f"image_adapter.layer_adapters.{i}.fc.0.weight"] = (W["I_ADAPTER"], i)
    for i in range(len(cfg.levels)):
        m[f"image_adapter.seg_proj.{i}.{fc}"] = (W["I_SEG_PROJ"], i)
    m[f"image_adapter.det_proj.{fc}"] = (W["I_DET_PROJ"], 0)
    if text and cfg.t_layers > 0:
        m["clip.token_embedding.weight"] = (W["T_TOKEN_EMB"], 0)
        m["clip.positional_embedding"] = (W["T_POS"], 0)
        m["clip.ln_final.weight"] = (W["T_LN_FINAL_G"], 0); m["clip.ln_final.bias"] = (W["T_LN_FINAL_B"], 0)
        for i in range(cfg.t_layers):
            for k, wid in _block_keys(f"transformer.resblocks.{i}.", "T"):
                m["clip." + k] = (wid, i)
        for i in range(cfg.text_adapt_until):
            m[f"text_adapter.{i}.fc.0.weight"] = (W["T_ADAPTER"], i)
        m[f"text_adapter.{cfg.text_adapt_until}.fc.0.weight"] = (W["T_FINAL_PROJ"], 0)
    return m


class Engine:
    """One context on one B200.  Not thread-safe (neither is the reference's nn.Module)."""

    def __init__(self, cfg: ModelCfg, device: int = 0, max_batch: int = 64, max_text: int = 16,
                 text: bool = True, cta_group: int = 0, ln_fold: Optional[bool] = None):
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = device
        self.max_batch = max_batch
        self.has_text = bool(text and cfg.t_layers > 0)
        c = AaclipCfg()
        c.image_size, c.patch_size, c.width, c.heads = cfg.image_size, cfg.patch_size, cfg.width, cfg.heads
        c.layers, c.mlp_width, c.embed_dim = cfg.layers, cfg.mlp_width, cfg.embed_dim
        c.act = ACT_QUICK_GELU if cfg.quick_gelu else ACT_GELU_ERF
        c.image_adapt_until, c.image_adapt_weight = cfg.image_adapt_until, cfg.image_adapt_weight
        c.n_levels = len(cfg.levels)
        for i, l in enumerate(cfg.levels):
            c.levels[i] = int(l)
        c.proj_relu = int(cfg.relu)
        if self.has_text:
            c.t_context, c.t_vocab, c.t_width, c.t_heads, c.t_layers = (cfg.t_context, cfg.t_vocab, cfg.t_width,
                                                                        cfg.t_heads, cfg.t_layers)
            c.text_adapt_until, c.text_adapt_weight = cfg.text_adapt_until, cfg.text_adapt_weight
        c.max_batch, c.max_text, c.cta_group = max_batch, max_text, cta_group
        c.ln_fold = 0 if ln_fold is None else (1 if ln_fold else 2)
        self._ctx = C.c_void_p()
        check(self.lib.aaclip_create(C.byref(self._ctx), C.byref(c), device))
        self._wmap = weight_map(cfg, self.has_text)

    def close(self) -> None:
        if getattr(self, "_ctx", None) and self._ctx.value:
            self.lib.aaclip_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def set_weight(self, full_key: str, tensor: torch.Tensor) -> None:
        wid, layer = self._wmap[full_key]
        t = tensor.detach()
        if t.dtype != torch.float32:
            t = t.float()
        t = t.contiguous()
        on_host = not t.is_cuda
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(self.lib.aaclip_set_weight(self._ctx, wid, layer, t.data_ptr(), t.numel(), int(on_host), stream))

    def load_state_dicts(self, clip_sd: Optional[Dict[str, torch.Tensor]] = None,
                         image_adapter_sd: Optional[Dict[str, torch.Tensor]] = None,
                         text_adapter_sd: Optional[Dict[str, torch.Tensor]] = None) -> List[str]:
        """Upload every tensor the hot path reads; returns the keys that were consumed."""
        used = []
        for prefix, sd in (("clip.", clip_sd), ("image_adapter.", image_adapter_sd), ("text_adapter.", text_adapter_sd)):
            if sd is None:
                continue
            for k, v in sd.items():
                fk = prefix + k
                if fk in self._wmap:
                    self.set_weight(fk, v)
                    used.append(fk)
        torch.cuda.synchronize(self.device)
        return used

    @property
    def device_bytes(self) -> int:
        return int(self.lib.aaclip_device_bytes(self._ctx))

    @property
    def launch_count(self) -> int:
        return int(self.lib.aaclip_launch_count(self._ctx))

    def profile(self, on: bool) -> None:
        check(self.lib.aaclip_profile_enable(self._ctx, int(on)))

    def profile_read(self) -> Dict[str, tuple]:
        """{kernel class: (total ms, launches)} since the last read (CUDA events around every launch)."""
        n = len(_lib.PROFILE_CLASSES)
        ms = (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        check(self.lib.aaclip_profile_read(self._ctx, ms, cnt, n))
        self.profile_span_ms = float(self.lib.aaclip_profile_span_ms(self._ctx))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(_lib.PROFILE_CLASSES)}

    # ------------------------------------------------------------------ forward
    def _check_image(self, image: torch.Tensor) -> None:
        S = self.cfg.image_size
        if image.dim() != 4 or tuple(image.shape[1:]) != (3, S, S):
            raise ValueError(f"image must be [B,3,{S},{S}], got {tuple(image.shape)}")
        if not image.is_cuda or image.dtype != torch.float32 or not image.is_contiguous():
            raise ValueError("image must be a contiguous float32 CUDA tensor")

    def _stream(self) -> int:
        return cur_stream(self.device)

    def visual_forward(self, image: torch.Tensor, want_seg: bool = True, want_det: bool = True,
                       seg_dtype: torch.dtype = torch.float32):
        """AdaptedCLIP.forward: returns (list of [B,P,E] L2-normalised patch tokens, fp32 [B,E]).  `seg_dtype` is
        torch.float32 (what the reference returns) or torch.bfloat16 (half the bytes for the anomaly-map head)."""
        self._check_image(image)
        if seg_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("seg_dtype must be torch.float32 or torch.bfloat16")
        B = image.shape[0]
        cfg = self.cfg
        seg = [torch.empty(B, cfg.patches, cfg.embed_dim, device=image.device, dtype=seg_dtype)
               for _ in cfg.levels] if want_seg else []
        det = torch.empty(B, cfg.embed_dim, device=image.device, dtype=torch.float32) if want_det else None
        arr = (C.c_void_p * len(cfg.levels))(*[t.data_ptr() for t in seg]) if want_seg else None
        check(self.lib.aaclip_visual_forward(self._ctx, ptr(image), B, arr, int(seg_dtype == torch.bfloat16), ptr(det),
                                             self._stream()))
        return seg, det

    def set_text_final(self, leaky: bool) -> None:
        """text_adapter[-1] (Linear + LeakyReLU, the default) or a plain projection (CLIP's text_projection) after ln_final."""
        check(self.lib.aaclip_set_text_final(self._ctx, int(bool(leaky))))

    def dapm_replace(self, dpam_layer: Optional[int]) -> None:
        """VisionTransformer.DAPM_replace(DPAM_layer) (model/transformer.py:406-425): the last DPAM_layer - 1 visual
        blocks use the batch-coupled v-v attention (:123-152).  None / 0 / 1 restores ordinary attention."""
        check(self.lib.aaclip_dapm_replace(self._ctx, int(dpam_layer or 0)))

    def encode_image(self, image: torch.Tensor, want_tokens: bool = True, want_pooled: bool = True,
                     normalize: bool = False):
        """CLIP.encode_image(image, cfg.levels, normalize) (model/model.py:185-188): returns (pooled fp32 [B,E] or
        None, list of fp32 [B,L,width] residual-stream tokens after each block in cfg.levels).  The context must have
        been loaded with visual.proj^T in its seg_proj slots (aaclip_b200.surgery.CLIPImageEncoder does that)."""
        self._check_image(image)
        B = image.shape[0]
        cfg = self.cfg
        toks = [torch.empty(B, cfg.tokens, cfg.width, device=image.device, dtype=torch.float32)
                for _ in cfg.levels] if want_tokens else []
        pooled = torch.empty(B, cfg.embed_dim, device=image.device, dtype=torch.float32) if want_pooled else None
        arr = (C.c_void_p * len(cfg.levels))(*[t.data_ptr() for t in toks]) if want_tokens else None
        check(self.lib.aaclip_encode_image(self._ctx, ptr(image), B, arr, ptr(pooled), int(normalize), self._stream()))
        return pooled, toks

    def forward_fused(self, image: torch.Tensor, anchors: torch.Tensor, domain: str = "Industrial",
                      want_maps: bool = True, want_scores: bool = True, out=None, extrema: Optional[torch.Tensor] = None):
        """image -> (level-summed anomaly maps fp32 [B,S,S], image scores fp32 [B]) without seg tokens.

        `extrema`: optional float32 CUDA tensor [B,2] that receives every image's (min, max) over its map, written by
        the head's own epilogue (the pixel-side input of metrics_eval, forward_utils.py:241-252).

        `out = (maps, scores)` reuses caller-owned output tensors.  On a non-default stream a call whose pointers
        (image, anchors, outputs) and batch repeat is replayed as ONE CUDA graph launch (captured at its second
        occurrence), so steady-state loops should reuse their buffers."""
        self._check_image(image)
        if tuple(anchors.shape) != (self.cfg.embed_dim, 2) or anchors.dtype != torch.float32 or not anchors.is_cuda:
            raise ValueError("anchors must be a float32 CUDA tensor [E,2]")
        B, S = image.shape[0], self.cfg.image_size
        if out is not None:
            maps, scores = out
            if maps is not None and (tuple(maps.shape) != (B, S, S) or maps.dtype != torch.float32 or not maps.is_cuda
                                     or not maps.is_contiguous()):
                raise ValueError(f"out maps must be a contiguous float32 CUDA tensor [{B},{S},{S}]")
            if scores is not None and (tuple(scores.shape) != (B,) or scores.dtype != torch.float32 or not scores.is_cuda):
                raise ValueError(f"out scores must be a float32 CUDA tensor [{B}]")
        else:
            maps = torch.empty(B, S, S, device=image.device, dtype=torch.float32) if want_maps else None
            scores = torch.empty(B, device=image.device, dtype=torch.float32) if want_scores else None
        if extrema is not None and (tuple(extrema.shape) != (B, 2) or extrema.dtype != torch.float32 or not extrema.is_cuda
                                    or not extrema.is_contiguous() or maps is None):
            raise ValueError(f"extrema must be a contiguous float32 CUDA tensor [{B},2] (and maps must be requested)")
        check(self.lib.aaclip_forward_fused(self._ctx, ptr(image), B, ptr(anchors.contiguous()), DOMAIN_MODE[domain],
                                            ptr(maps), ptr(scores), ptr(extrema), self._stream()))
        return maps, scores

    @staticmethod
    def _host_f32(tensors, what: str) -> None:
        for t in tensors:
            if t is not None and (t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
                raise ValueError(f"{what} takes contiguous float32 CPU tensors")

    def forward_fused_host(self, image: torch.Tensor, anchors: torch.Tensor, maps_out: torch.Tensor,
                           scores_out: torch.Tensor, domain: str = "Industrial",
                           extrema_out: Optional[torch.Tensor] = None) -> None:
        """Host-buffer entry: `image`, `anchors`, `maps_out`, `scores_out` (and `extrema_out` [B,2]) are CPU tensors
        (pinned for speed); H2D, compute and D2H all happen inside the call, which returns after the results landed."""
        self._host_f32((image, anchors, maps_out, scores_out, extrema_out), "forward_fused_host")
        check(self.lib.aaclip_forward_fused_host(self._ctx, image.data_ptr(), image.shape[0], anchors.data_ptr(),
                                                 DOMAIN_MODE[domain], maps_out.data_ptr(), scores_out.data_ptr(),
                                                 ptr(extrema_out)))

    def submit_host(self, image: torch.Tensor, anchors: torch.Tensor, maps_out: torch.Tensor,
                    scores_out: torch.Tensor, domain: str = "Industrial",
                    extrema_out: Optional[torch.Tensor] = None) -> int:
        """Asynchronous host-buffer entry: enqueues H2D -> forward -> D2H for one batch (<= max_batch images) and
        returns a ticket; at most two tickets may be pending.  The CPU tensors must stay alive until wait_host."""
        self._host_f32((image, anchors, maps_out, scores_out, extrema_out), "submit_host")
        ticket = C.c_longlong(-1)
        check(self.lib.aaclip_submit_host(self._ctx, image.data_ptr(), image.shape[0], anchors.data_ptr(),
                                          DOMAIN_MODE[domain], maps_out.data_ptr(), scores_out.data_ptr(),
                                          ptr(extrema_out), C.byref(ticket)))
        return int(ticket.value)

    def submit_host_u8(self, image_u8: torch.Tensor, anchors: torch.Tensor, maps_out: torch.Tensor,
                       scores_out: torch.Tensor, domain: str = "Industrial",
                       extrema_out: Optional[torch.Tensor] = None) -> int:
        """submit_host for RAW images: `image_u8` is a CPU uint8 tensor [B,H0,W0,3] (RGB, as PIL decodes it); the
        reference's transform_x (dataset/__init__.py:127-136) runs on the device, bit-exact with PIL + torchvision."""
        if image_u8.is_cuda or image_u8.dtype != torch.uint8 or not image_u8.is_contiguous() or image_u8.dim() != 4 \
                or image_u8.shape[3] != 3:
            raise ValueError("submit_host_u8 takes a contiguous uint8 CPU tensor [B,H0,W0,3]")
        self._host_f32((anchors, maps_out, scores_out, extrema_out), "submit_host_u8 (anchors and outputs)")
        ticket = C.c_longlong(-1)
        B, H0, W0, _ = image_u8.shape
        check(self.lib.aaclip_submit_host_u8(self._ctx, image_u8.data_ptr(), B, H0, W0, anchors.data_ptr(),
                                             DOMAIN_MODE[domain], maps_out.data_ptr(), scores_out.data_ptr(),
                                             ptr(extrema_out), C.byref(ticket)))
        return int(ticket.value)

    def wait_host(self, ticket: int) -> None:
        check(self.lib.aaclip_wait_host(self._ctx, ticket))

    def predict_stream(self, batches, anchors: torch.Tensor, domain: str = "Industrial", with_extrema: bool = False):
        """The loop of test.py:get_predictions (test.py:53-99) over an iterable of CPU image batches, pipelined:
        while batch k computes, batch k+1 uploads and batch k-1 downloads.  Yields (maps [B,S,S], scores [B]) as
        pinned CPU tensors, in order; with `with_extrema` a third tensor [B,2] = per-image (min, max) of the maps
        (from the head's epilogue: metrics_eval's normalisation then needs no pass over the pixels).  A batch is either float32 [B,3,S,S] (already transformed, as the reference's
        DataLoader delivers it) or uint8 [B,H0,W0,3] raw RGB, in which case the loader's transform_x
        (dataset/__init__.py:127-136) also runs on the device.  Batches larger than `max_batch` are split into chunks."""
        S = self.cfg.image_size
        anchors = anchors.detach().float().cpu().contiguous()
        inflight = []   # tickets in submission order: (ticket, batch record); at most two
        ready = []      # batch records in input order: {"left": chunks not yet landed, "out": (maps, scores), "keep": ...}

        def drain_one():
            t, rec = inflight.pop(0)
            self.wait_host(t)
            rec["left"] -= 1

        for img in batches:
            img = img.detach()
            raw = img.dtype == torch.uint8   # [B,H0,W0,3] undecoded-size RGB bytes: transform_x runs on the device
            if raw:
                if img.is_cuda or not img.is_contiguous():
                    img = img.cpu().contiguous()
            elif img.is_cuda or img.dtype != torch.float32 or not img.is_contiguous():
                img = img.float().cpu().contiguous()
            if not img.is_pinned():
                img = img.pin_memory()
            n = img.shape[0]
            maps = torch.empty(n, S, S).pin_memory()
            scores = torch.empty(n).pin_memory()
            ext = torch.empty(n, 2).pin_memory() if with_extrema else None
            rec = {"left": 0, "out": (maps, scores, ext) if with_extrema else (maps, scores), "keep": img}
            ready.append(rec)
            submit = self.submit_host_u8 if raw else self.submit_host
            for b0 in range(0, n, self.max_batch):   # a batch larger than max_batch rides the two slots in chunks
                b1 = min(n, b0 + self.max_batch)
                if len(inflight) == 2:
                    drain_one()
                inflight.append((submit(img[b0:b1], anchors, maps[b0:b1], scores[b0:b1], domain,
                                        ext[b0:b1] if with_extrema else None), rec))
                rec["left"] += 1
            while ready and ready[0]["left"] == 0 and all(r is not ready[0] for _, r in inflight):
                yield ready.pop(0)["out"]
        while inflight:
            drain_one()
            while ready and ready[0]["left"] == 0 and all(r is not ready[0] for _, r in inflight):
                yield ready.pop(0)["out"]
        for rec in ready:   # empty batches
            yield rec["out"]

    def text_forward(self, tokens: torch.Tensor) -> torch.Tensor:
        """AdaptedCLIP.encode_text(adapt_text=True): int32 [n, ctx] -> fp32 [n, t_width]."""
        if not self.has_text:
            raise RuntimeError("engine was created without the text tower")
        if tokens.dim() != 2 or tokens.shape[1] != self.cfg.t_context:
            raise ValueError(f"tokens must be [n,{self.cfg.t_context}]")
        tk = tokens.to(device=f"cuda:{self.device}", dtype=torch.int32).contiguous()
        out = torch.empty(tk.shape[0], self.cfg.t_width, device=tk.device, dtype=torch.float32)
        check(self.lib.aaclip_text_forward(self._ctx, ptr(tk), tk.shape[0], ptr(out), self._stream()))
        return out
