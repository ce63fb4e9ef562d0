"""Drop-in `AdaptedCLIP` (reference: model/adapter.py:6-145) backed by the sm_100a CUDA engine.

Same constructor, attributes (`clipmodel`, `image_encoder`, `image_adapter`, `text_adapter`, `i_w`, `t_w`,
`levels`, ...), state_dict keys and return values as the reference class, so test.py can use it unchanged:

    model = AdaptedCLIP(clip_model=clip_model, text_adapt_weight=..., image_adapt_weight=...,
                        text_adapt_until=..., image_adapt_until=..., relu=args.relu).to(device)
    model.eval()
    model.image_adapter.load_state_dict(ckpt["image_adapter"])       # test.py:175-176
    patch_features, det_feature = model(image)                       # test.py:80

`clip_model` may be the reference's own `CLIP` instance or `aaclip_b200.clip.CLIP`; only parameters are read
from it.  Parameters are mirrored into the engine's packed device layout lazily and re-mirrored whenever a
parameter's version counter changes (i.e. after any load_state_dict / in-place update).

Inference only (the north star's scope): forward() runs under no_grad semantics and returns tensors that do
not require grad.  There is no PyTorch fallback - on a machine without a B200 the first forward raises.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from .adapter_modules import SimpleAdapter, SimpleProj
from .engine import Engine
from .synth import ModelCfg


def _infer_cfg(clip_model, levels, image_adapt_until, text_adapt_until, i_w, t_w, relu) -> ModelCfg:
    v = clip_model.visual
    w = v.conv1.out_channels
    ps = v.conv1.kernel_size[0]
    tokens = v.positional_embedding.shape[0]
    grid = int(round((tokens - 1) ** 0.5))
    if grid * grid + 1 != tokens:
        raise ValueError(f"positional_embedding has {tokens} rows: not a square patch grid + class token")
    blocks = v.transformer.resblocks
    heads = blocks[0].attn.num_heads
    gelu = blocks[0].mlp.gelu
    quick = type(gelu).__name__ == "QuickGELU"
    if not quick and not isinstance(gelu, nn.GELU):
        raise ValueError(f"unsupported MLP activation {type(gelu).__name__}")
    if isinstance(gelu, nn.GELU) and getattr(gelu, "approximate", "none") != "none":
        raise ValueError("only exact (erf) nn.GELU is supported")
    kw = dict(image_size=grid * ps, patch_size=ps, width=w, layers=len(blocks), heads=heads,
              mlp_ratio=blocks[0].mlp.c_fc.out_features / w, embed_dim=768, quick_gelu=quick,
              image_adapt_until=image_adapt_until, image_adapt_weight=i_w, levels=list(levels), relu=bool(relu),
              text_adapt_until=text_adapt_until, text_adapt_weight=t_w, t_layers=0)
    if hasattr(clip_model, "token_embedding"):
        tb = clip_model.transformer.resblocks
        kw.update(t_context=clip_model.positional_embedding.shape[0], t_vocab=clip_model.token_embedding.weight.shape[0],
                  t_width=clip_model.token_embedding.weight.shape[1], t_heads=tb[0].attn.num_heads, t_layers=len(tb))
    return ModelCfg(**kw)


def effective_levels(levels, layers: int) -> List[int]:
    """The taps the reference actually takes: block i is tapped when `i + 1 in self.levels` (model/adapter.py:100) -
    membership semantics, so duplicates fire once, order is irrelevant and a level outside [1, layers] never fires."""
    return [l for l in sorted({int(l) for l in levels}) if 1 <= l <= layers]


class AdaptedCLIP(nn.Module):
    def __init__(
        self,
        clip_model,
        text_adapt_weight: float = 0.1,
        image_adapt_weight: float = 0.1,
        text_adapt_until: int = 3,
        image_adapt_until: int = 6,
        levels: list = [6, 12, 18, 24],
        relu: bool = True,
        max_batch: int = 64,
        seg_dtype: torch.dtype = torch.float32,
        **kwargs,
    ):
        super().__init__()
        self.clipmodel = clip_model
        self.image_encoder = clip_model.visual
        self.text_adapt_until = text_adapt_until
        self.image_adapt_until = image_adapt_until
        self.t_w = text_adapt_weight
        self.i_w = image_adapt_weight
        self.levels = levels
        self.relu = relu
        self.max_batch = max_batch
        # dtype of the patch tokens forward() returns: float32 as the reference; bfloat16 halves what the head streams
        self.seg_dtype = seg_dtype
        width = clip_model.visual.conv1.out_channels
        t_width = clip_model.token_embedding.weight.shape[1] if hasattr(clip_model, "token_embedding") else 768

        layer_adapters = nn.ModuleList([SimpleAdapter(width, width) for _ in range(image_adapt_until)])
        seg_proj = nn.ModuleList([SimpleProj(width, 768, relu) for _ in range(len(levels))])
        det_proj = SimpleProj(width, 768, relu)
        self.image_adapter = nn.ModuleDict(
            {"layer_adapters": layer_adapters, "seg_proj": seg_proj, "det_proj": det_proj}
        )
        self.text_adapter = nn.ModuleList(
            [SimpleAdapter(t_width, t_width) for _ in range(text_adapt_until)] + [SimpleProj(t_width, t_width, relu=True)]
        )
        self._init_weights_()
        self._engine: Optional[Engine] = None
        self._engine_sig = None
        self._versions: Dict[str, Tuple[int, int]] = {}

    def _init_weights_(self):
        # model/adapter.py:47-53
        for p in self.image_adapter.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for p in self.text_adapter.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    # ---------------------------------------------------------------------------------- engine plumbing
    def _cfg_signature(self):
        return (tuple(int(l) for l in self.levels), int(self.image_adapt_until), int(self.text_adapt_until), float(self.i_w),
                float(self.t_w), bool(self.relu), int(self.max_batch))

    def _named_sources(self):
        for k, v in self.clipmodel.state_dict(keep_vars=True).items():
            yield "clip." + k, v
        for k, v in self.image_adapter.state_dict(keep_vars=True).items():
            yield "image_adapter." + k, v
        for k, v in self.text_adapter.state_dict(keep_vars=True).items():
            yield "text_adapter." + k, v

    def _sync_engine(self, device: torch.device) -> Engine:
        if device.type != "cuda":
            raise RuntimeError("aaclip_b200.AdaptedCLIP runs on a B200 only (no CPU / PyTorch fallback); "
                               "move the model and its inputs to a cuda device")
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        # the reference reads these attributes on every forward (model/adapter.py:90-101, 123-131): a change made after the
        # first forward must take effect here too, so the context (and its cached graphs) is rebuilt when one differs
        sig = self._cfg_signature()
        if self._engine is not None and (self._engine.device != dev_index or self._engine_sig != sig):
            self._engine.close()
            self._engine = None
        if self._engine is None:
            self._engine_sig = sig
            cfg = _infer_cfg(self.clipmodel, list(self.levels), self.image_adapt_until, self.text_adapt_until, self.i_w,
                             self.t_w, self.relu)
            cfg.levels = effective_levels(self.levels, cfg.layers)
            if not cfg.levels:
                raise ValueError(f"no level of {list(self.levels)} lies in [1, {cfg.layers}]: nothing to tap")
            self._engine = Engine(cfg, device=dev_index, max_batch=self.max_batch)
            self._versions = {}
        eng = self._engine
        dirty = False
        for key, t in self._named_sources():
            if key not in eng._wmap:
                continue
            sig = (t.data_ptr(), t._version)
            if self._versions.get(key) != sig:
                eng.set_weight(key, t)
                self._versions[key] = sig
                dirty = True
        if dirty:
            torch.cuda.synchronize(dev_index)
        return eng

    # ---------------------------------------------------------------------------------- reference surface
    def forward_original(self, x, modality="visual"):
        raise NotImplementedError("forward_original is dead code in the reference (model/adapter.py:55-65)")

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """model/adapter.py:67-112 -> (seg_tokens: list of [B,P,768] L2-normalised, det_token [B,768])."""
        eng = self._sync_engine(x.device)
        seg, det = eng.visual_forward(x.float().contiguous(), seg_dtype=self.seg_dtype)
        return seg, det

    @torch.no_grad()
    def encode_text(self, text: torch.Tensor, adapt_text: bool = True):
        """model/adapter.py:114-145."""
        if not adapt_text:
            return self.clipmodel.encode_text(text)
        eng = self._sync_engine(text.device if text.is_cuda else next(self.parameters()).device)
        return eng.text_forward(text)

    @torch.no_grad()
    def predict(self, image: torch.Tensor, text_feature: torch.Tensor, domain: str = "Industrial",
                with_extrema: bool = False):
        """Fused body of test.py:get_predictions for one batch (test.py:80-93): returns
        (anomaly maps [B,S,S] = sum over levels of calculate_similarity_map(test=True), image scores [B]) and, with
        `with_extrema`, the per-image (min, max) [B,2] of the maps (forward_utils.py:241-252) from the same kernel."""
        eng = self._sync_engine(image.device)
        ext = torch.empty(image.shape[0], 2, device=image.device, dtype=torch.float32) if with_extrema else None
        maps, scores = eng.forward_fused(image.float().contiguous(), text_feature.float().contiguous(), domain, extrema=ext)
        return (maps, scores, ext) if with_extrema else (maps, scores)

    @torch.no_grad()
    def predict_stream(self, batches, text_feature: torch.Tensor, domain: str = "Industrial", with_extrema: bool = False):
        """The batch loop of test.py:get_predictions (test.py:60-99) over an iterable of CPU image batches, pipelined
        (upload of batch k+1 and download of batch k-1 overlap the compute of batch k).  Yields
        (maps [B,S,S], scores [B]) as pinned CPU tensors, in order.  Batches are float32 [B,3,S,S] or raw uint8
        [B,H0,W0,3] (the loader's transform_x then runs on the device); larger ones than `max_batch` go in chunks."""
        eng = self._sync_engine(next(self.parameters()).device)
        yield from eng.predict_stream(batches, text_feature, domain, with_extrema=with_extrema)
