"""CPU: host logic — the C-ABI library loads and exports every symbol include/aaclip_b200.h declares, fails
loudly without a GPU (no fallback), the drop-in module tree has the reference's state_dict keys, and the
N>1 sharding / score gather works over gloo with world_size 2."""
import ctypes as C
import os
import re
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from aaclip_b200 import _lib, synth  # noqa: E402
from aaclip_b200.dist import gather_scores, shard_range  # noqa: E402


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "aaclip_b200.h")).read()
    declared = set(re.findall(r"\b(aaclip_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    lib = C.CDLL(str(_lib.LIB_PATH))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().aaclip_abi_version() == 2


def test_cfg_struct_matches_header():
    hdr = open(os.path.join(ROOT, "include", "aaclip_b200.h")).read()
    body = hdr[hdr.index("typedef struct {"):hdr.index("} aaclip_cfg;")]
    fields = re.findall(r"^\s*(?:int|float)\s+([a-z_]+)(?:\[\d+\])?;", body, flags=re.M)
    assert fields == [f[0] for f in _lib.AaclipCfg._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly():
    lib = _lib.load()
    ctx = C.c_void_p()
    cfg = _lib.AaclipCfg()
    rc = lib.aaclip_create(C.byref(ctx), C.byref(cfg), 0)
    assert rc == -3 and b"no CPU fallback" in lib.aaclip_last_error()
    from aaclip_b200 import ops
    with pytest.raises(ValueError, match="CUDA tensor"):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    from aaclip_b200.adapter import AdaptedCLIP
    from aaclip_b200.clip import CLIP
    m = AdaptedCLIP(CLIP(synth.tiny_cfg(), text=False), relu=False, image_adapt_until=2, levels=[1, 2]).eval()
    with pytest.raises(RuntimeError, match="no CPU"):
        m(torch.zeros(1, 3, 56, 56))


def test_dropin_state_dict_keys_match_reference():
    """Keys probed from the reference modules; checkpoints must load unchanged."""
    from aaclip_b200.adapter import AdaptedCLIP
    from aaclip_b200.clip import CLIP
    cfg = synth.tiny_cfg()
    for relu in (False, True):
        m = AdaptedCLIP(CLIP(cfg), text_adapt_until=3, image_adapt_until=6, relu=relu)
        ia = list(m.image_adapter.state_dict().keys())
        fc = "fc.0.weight" if relu else "fc.weight"
        assert ia == [f"layer_adapters.{i}.fc.0.weight" for i in range(6)] + \
            [f"seg_proj.{i}.{fc}" for i in range(4)] + [f"det_proj.{fc}"]
        assert list(m.text_adapter.state_dict().keys()) == [f"{i}.fc.0.weight" for i in range(4)]
    # the synthetic state dicts use exactly the container's keys
    full = synth.VIT_L_14_336
    clip_keys = set(CLIP(synth.tiny_cfg(layers=2, t_layers=2)).state_dict().keys())
    synth_keys = set(synth.clip_state_dict(synth.tiny_cfg(layers=2, t_layers=2)).keys())
    assert synth_keys == clip_keys, synth_keys ^ clip_keys
    assert synth.clip_state_dict(synth.tiny_cfg(), 0)["visual.conv1.weight"].shape == (256, 3, 14, 14)
    assert full.tokens == 577 and full.mlp_width == 4096


def test_pos_embed_resize_matches_reference_golden():
    """Checkpoint ingestion for the 518-px setting: 24x24 -> 37x37 positional-embedding rescale vs the output of the
    reference's resize_pos_embed (model/model.py:395-426) stored by oracle/make_golden.py."""
    from aaclip_b200.clip import CLIP, load_checkpoint, resize_pos_embed
    g = torch.load(os.path.join(ROOT, "tests", "golden", "pos_embed_24_to_37.pt"), weights_only=False)
    old = torch.randn(1 + g["old_grid"] ** 2, g["width"], generator=torch.Generator().manual_seed(g["seed"]))
    sd = {"visual.positional_embedding": old.clone()}
    assert resize_pos_embed(sd, g["new_grid"])
    assert sd["visual.positional_embedding"].shape == g["resized"].shape
    assert (sd["visual.positional_embedding"] - g["resized"]).abs().max().item() < 1e-6
    assert torch.equal(sd["visual.positional_embedding"][0], old[0])          # class-token row untouched
    assert not resize_pos_embed(sd, g["new_grid"])                            # already at the target grid
    # a 336-px state dict loads into a 518-px container through load_checkpoint
    cfg336, cfg518 = synth.tiny_cfg(image_size=56), synth.tiny_cfg(image_size=84)
    sd336 = synth.clip_state_dict(cfg336, 0)
    m518 = CLIP(cfg518)
    load_checkpoint(m518, sd336)
    assert m518.visual.positional_embedding.shape == (cfg518.tokens, cfg518.width)


def test_weight_map_covers_hot_path():
    from aaclip_b200.engine import weight_map
    cfg = synth.tiny_cfg()
    wm = weight_map(cfg)
    keys = ["clip." + k for k in synth.clip_state_dict(cfg)] + \
        ["image_adapter." + k for k in synth.image_adapter_state_dict(cfg)] + \
        ["text_adapter." + k for k in synth.text_adapter_state_dict(cfg)]
    unused = [k for k in keys if k not in wm]
    assert sorted(unused) == ["clip.logit_scale", "clip.text_projection", "clip.visual.proj"]  # not on the path
    assert set(wm) <= set(keys)
    assert len({v for v in wm.values()}) == len(wm)  # (id, layer) pairs are unique


def test_synth_is_deterministic():
    a = synth.clip_state_dict(synth.tiny_cfg(), 0)
    b = synth.clip_state_dict(synth.tiny_cfg(), 0)
    c = synth.clip_state_dict(synth.tiny_cfg(), 1)
    k = "visual.transformer.resblocks.1.mlp.c_fc.weight"
    assert torch.equal(a[k], b[k]) and not torch.equal(a[k], c[k])
    t = synth.tokens(5, synth.tiny_cfg())
    assert t.dtype == torch.int32 and (t.argmax(-1) > 0).all() and (t[:, 0] == 510).all()


@pytest.mark.parametrize("total,world", [(1024, 8), (10, 4), (3, 8), (0, 2), (64, 1)])
def test_shard_range_partitions(total, world):
    spans = [shard_range(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    sizes = [e - b for b, e in spans]
    assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, total, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(total, dtype=torch.float32) * 0.5 + 1
    b, e = shard_range(total, rank, world)
    out = gather_scores(full[b:e].clone(), total)
    ok = bool(torch.equal(out, full))
    if total % world == 0:   # even shards: one collective straight into a caller-owned buffer
        buf = torch.empty(total)
        ok = ok and gather_scores(full[b:e].clone(), total, out=buf) is buf and bool(torch.equal(buf, full))
    # per-image records ride the same collective: the map extrema [n, 2] and their global (min, max)
    from aaclip_b200.dist import gather_extrema
    ex_full = torch.stack([full - 3.0, full * 2.0], 1)
    allx, glob = gather_extrema(ex_full[b:e].clone(), total)
    ok = ok and bool(torch.equal(allx, ex_full)) and bool(torch.equal(glob, torch.stack([ex_full[:, 0].min(), ex_full[:, 1].max()])))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [16, 7])
def test_gather_scores_gloo_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611 + total
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, total, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_image_level_preds_matches_metrics_eval_restatement():
    """forward_utils.image_level_preds (from per-image map extrema) == the oracle's restatement of
    metrics_eval's normalisation + pmax mix (forward_utils.py:241-254) on the full pixel arrays."""
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import aaclip_oracle as orc
    from aaclip_b200 import forward_utils as fu
    rng = np.random.RandomState(3)
    maps = (rng.rand(9, 24, 24).astype(np.float32) * 7 + 1.5)
    scores = rng.rand(9).astype(np.float32)
    ex = np.stack([maps.min(axis=(1, 2)), maps.max(axis=(1, 2))], 1)
    for domain in ("Industrial", "Medical"):
        _, want = orc.metrics_image_preds(maps, scores, domain)
        got = fu.image_level_preds(ex, scores, domain)
        assert np.array_equal(got, want.astype(np.float32))
    # the reference's "already normalised" branch: max exactly 1 leaves the values alone
    maps1 = maps / maps.max()
    ex1 = np.stack([maps1.min(axis=(1, 2)), maps1.max(axis=(1, 2))], 1)
    _, want = orc.metrics_image_preds(maps1, scores, "Industrial")
    assert np.array_equal(fu.image_level_preds(ex1, scores, "Industrial"), want.astype(np.float32))


def test_new_entry_points_argument_checks_without_gpu():
    """Host-side validation of the loader-transform / extrema / folded-GEMM entry points: pure arithmetic helpers and
    the argument errors that are raised before any CUDA call (no compute without a GPU)."""
    import ctypes as C
    from aaclip_b200 import _lib
    lib = _lib.load()
    # scratch = uint8 result of the horizontal pass, [B, H0, S, 3]; none when the width does not change
    assert lib.aaclip_preprocess_scratch_bytes(64, 1024, 1024, 336) == 64 * 1024 * 336 * 3
    assert lib.aaclip_preprocess_scratch_bytes(2, 500, 336, 336) == 0
    assert lib.aaclip_preprocess_scratch_bytes(0, 10, 10, 5) == 0
    # empty batches are a no-op, null pointers / bad sizes are AACLIP_ERR_INVALID with a message
    assert lib.aaclip_preprocess_u8(None, 0, 8, 8, 4, None, None, None, None, None) == 0
    assert lib.aaclip_map_minmax(None, 0, 16, None, None) == 0
    assert lib.aaclip_preprocess_u8(None, 1, 8, 8, 4, None, None, None, None, None) == -1
    assert b"null" in lib.aaclip_last_error()
    buf = (C.c_uint8 * 16)()
    out = (C.c_float * 16)()
    assert lib.aaclip_preprocess_u8(buf, 1, 0, 8, 4, None, None, None, out, None) == -1      # H0 = 0
    assert lib.aaclip_preprocess_u8(buf, 1, 8, 20000, 4, None, None, buf, out, None) == -1  # row too wide for smem
    assert b"too wide" in lib.aaclip_last_error()
    assert lib.aaclip_preprocess_u8(buf, 1, 8, 8, 4, None, None, None, out, None) == -1      # scratch missing
    assert b"scratch" in lib.aaclip_last_error()
    assert lib.aaclip_map_minmax(None, 2, 16, None, None) == -1
    assert lib.aaclip_fold_ln_weight(None, None, None, None, 8, 8, None, None, None, None) == -1
    # folded-GEMM building blocks: shape rules are checked before the device is touched
    assert lib.aaclip_gemm_resid_ln(buf, 64, buf, 64, 128, 384, 64, None, out, 384, buf, 384, out, 2, None) == -1
    assert b"256" in lib.aaclip_last_error()                                                 # N % 256 != 0
    assert lib.aaclip_gemm_lnfold(buf, 64, buf, 64, 128, 256, 64, None, None, out, 8, 1e-5, buf, 256, 0, 2, None) == -1
    assert b"colsum" in lib.aaclip_last_error()


def test_head_workspace_contract_without_gpu():
    """aaclip_anomaly_head takes its scratch from the caller (no allocation, no global state inside the library): the
    size query and the argument errors are host-side and run without a device."""
    import ctypes as C
    from aaclip_b200 import _lib
    lib = _lib.load()
    need = lib.aaclip_anomaly_head_workspace_bytes(4, 64, 576)
    assert need >= 4 * 64 * 576 * 2 * 4                      # the general form's dots [levels][B*P][2]
    assert lib.aaclip_anomaly_head_workspace_bytes(1, 64, 576) >= (64 * 576 + 64) * 4   # streaming form: scalars + counters
    assert lib.aaclip_anomaly_head_workspace_bytes(0, 1, 1) == 0
    seg = (C.c_void_p * 1)(0x1000)
    f = (C.c_float * 4)()
    # empty batch: no-op;  maps without workspace, non-square grid, extrema without maps: ERR_INVALID
    assert lib.aaclip_anomaly_head(seg, 1, 0, f, 0, None, 0, 576, 768, 336, 0, f, None, None, None, 0, None) == 0
    assert lib.aaclip_anomaly_head(seg, 1, 0, f, 0, None, 2, 576, 768, 336, 0, f, None, None, None, 0, None) == -1
    assert b"workspace" in lib.aaclip_last_error()
    assert lib.aaclip_anomaly_head(seg, 1, 0, f, 0, None, 2, 577, 768, 336, 0, f, None, None, f, need, None) == -1
    assert b"square" in lib.aaclip_last_error()
    assert lib.aaclip_anomaly_head(seg, 1, 0, f, 0, None, 2, 576, 768, 336, 0, None, None, f, f, need, None) == -1
    assert b"extrema" in lib.aaclip_last_error()
    assert lib.aaclip_anomaly_head(seg, 1, 0, f, 0, None, 2, 576, 768, 336, 0, None, f, None, f, need, None) == -1
    assert b"det" in lib.aaclip_last_error()


def test_effective_levels_follow_the_reference_membership_test():
    """model/adapter.py:100 taps block i when `i + 1 in self.levels`: duplicates fire once, order does not matter and
    out-of-range levels never fire ([6, 6, 12] used to be rejected by aaclip_create)."""
    from aaclip_b200.adapter import effective_levels
    assert effective_levels([6, 12, 18, 24], 24) == [6, 12, 18, 24]
    assert effective_levels([12, 6, 6, 24, 18, 12], 24) == [6, 12, 18, 24]
    assert effective_levels([0, 6, 25, 24, -3], 24) == [6, 24]
    assert effective_levels([30], 24) == []


def test_openai_state_dict_ingestion_vs_reference_golden():
    """(f)3: an OpenAI-layout checkpoint (fp16 Linear / Conv / attention tensors, metadata scalars) through
    aaclip_b200.clip.load_openai_state_dict equals, tensor for tensor and bit for bit, what the reference's own chain
    (load_openai_model -> build_model_from_openai_state_dict -> .float() -> CLIP(**json, image_size) -> resize_pos_embed ->
    load_state_dict; golden written by oracle/make_golden.py from the real reference) leaves in the model."""
    import hashlib
    from aaclip_b200.clip import cfg_from_openai_state_dict, load_openai_state_dict
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "openai_ingest_tiny.pt"), weights_only=False)
    cfg = synth.ModelCfg(**g["cfg"])
    sd16 = synth.openai_style_state_dict(cfg, g["seed"])
    assert sd16["visual.transformer.resblocks.0.attn.in_proj_weight"].dtype == torch.float16
    assert sd16["visual.ln_pre.weight"].dtype == torch.float32 and "input_resolution" in sd16
    inferred = cfg_from_openai_state_dict(sd16)
    assert (inferred.width, inferred.layers, inferred.patch_size, inferred.image_size, inferred.heads) == (128, 2, 14, 56, 2)
    assert (inferred.t_width, inferred.t_heads, inferred.t_layers, inferred.t_vocab, inferred.embed_dim) == (64, 1, 2, 128, 32)
    model = load_openai_state_dict(sd16, img_size=g["target_image_size"])
    got = {k: v.detach().float().contiguous() for k, v in model.state_dict().items()}
    assert set(got) == set(g["sha256"])
    assert all(v.dtype == torch.float32 for v in model.state_dict().values() if torch.is_floating_point(v))
    bad = [k for k, v in got.items() if hashlib.sha256(v.numpy().tobytes()).hexdigest() != g["sha256"][k]]
    assert not bad, bad
    assert torch.equal(got["visual.positional_embedding"], g["visual.positional_embedding"])
    assert got["visual.positional_embedding"].shape[0] == (g["target_image_size"] // 14) ** 2 + 1
    # the wrapped form model/openai.py:70-72 also accepts: {"state_dict": {"module.<key>": tensor}}
    wrapped = {"state_dict": {"module." + k: v for k, v in sd16.items()}}
    again = load_openai_state_dict(wrapped, img_size=g["target_image_size"])
    assert all(torch.equal(a, b) for a, b in zip(again.state_dict().values(), model.state_dict().values()))
    with pytest.raises(ValueError):
        load_openai_state_dict({k: v for k, v in sd16.items() if k != "visual.proj"})


REFERENCE = os.environ.get("AACLIP_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "model")), reason="the reference tree is not on this machine")
def test_reference_clip_object_maps_onto_engine_weights():
    """Drop-in guard: the REAL reference CLIP(**ViT-L-14-336.json), built on the meta device (no
    memory, no init), wrapped in aaclip_b200.AdaptedCLIP: the architecture the engine infers is exactly VIT_L_14_336 and
    every tensor the engine wants (weight_map) is one the reference module tree actually has, with the expected shape."""
    import importlib
    import json
    import types
    for name in ("ipdb", "ftfy"):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["ftfy"].fix_text = getattr(sys.modules["ftfy"], "fix_text", lambda t: t)
    sys.path.insert(0, REFERENCE)
    try:
        from model.model import CLIP as RefCLIP
        jcfg = json.load(open(os.path.join(REFERENCE, "model/model_configs/ViT-L-14-336.json")))
        with torch.device("meta"):
            ref_clip = RefCLIP(**jcfg)
        from aaclip_b200.adapter import AdaptedCLIP, _infer_cfg, effective_levels
        model = AdaptedCLIP(clip_model=ref_clip, text_adapt_weight=0.1, image_adapt_weight=0.1, text_adapt_until=3,
                            image_adapt_until=6, levels=[6, 12, 18, 24], relu=False)
        cfg = _infer_cfg(ref_clip, effective_levels(model.levels, 24), 6, 3, 0.1, 0.1, False)
        want = synth.VIT_L_14_336
        for f in ("image_size", "patch_size", "width", "layers", "heads", "mlp_width", "embed_dim", "quick_gelu", "t_context",
                  "t_vocab", "t_width", "t_heads", "t_layers", "levels", "image_adapt_until", "text_adapt_until", "relu"):
            assert getattr(cfg, f) == getattr(want, f), f
        from aaclip_b200.engine import weight_map
        wm = weight_map(cfg)
        sources = dict(model._named_sources())
        missing = [k for k in wm if k not in sources]
        assert not missing, missing[:5]
        assert tuple(sources["clip.visual.conv1.weight"].shape) == (1024, 3, 14, 14)
        assert tuple(sources["clip.visual.positional_embedding"].shape) == (577, 1024)
        assert tuple(sources["clip.visual.transformer.resblocks.23.attn.in_proj_weight"].shape) == (3072, 1024)
        assert tuple(sources["clip.transformer.resblocks.11.mlp.c_fc.weight"].shape) == (3072, 768)
        assert tuple(sources["image_adapter.seg_proj.3.fc.weight"].shape) == (768, 1024)
        assert tuple(sources["text_adapter.3.fc.0.weight"].shape) == (768, 768)
    finally:
        sys.path.remove(REFERENCE)
        for m in [m for m in sys.modules if m == "model" or m.startswith("model.")]:
            del sys.modules[m]


def test_container_block_forward_raises_clearly():
    from aaclip_b200.clip import CLIP
    blk = CLIP(synth.tiny_cfg(), text=False).visual.transformer.resblocks[0]
    with pytest.raises(NotImplementedError, match="CUDA engine"):
        blk(torch.zeros(5, 1, 256))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "model")), reason="the reference tree is not on this machine")
def test_dapm_replaced_reference_clip_maps_onto_the_surgery_encoder():
    """(f)4 guard: a REAL reference CLIP whose visual tower went through DAPM_replace(20) (train.py:243) carries
    `attn.qkv` / `attn.proj` Linears in its last 19 blocks.  CLIPImageEncoder reads the surgery depth off those keys, and
    every visual tensor the engine wants is found under its canonical (nn.MultiheadAttention) name with the same shape."""
    import importlib
    import json
    import types
    for name in ("ipdb", "ftfy"):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["ftfy"].fix_text = getattr(sys.modules["ftfy"], "fix_text", lambda t: t)
    sys.path.insert(0, REFERENCE)
    try:
        from model.model import CLIP as RefCLIP
        jcfg = json.load(open(os.path.join(REFERENCE, "model/model_configs/ViT-L-14-336.json")))
        with torch.device("meta"):
            ref_clip = RefCLIP(**jcfg)
            ref_clip.visual.DAPM_replace(DPAM_layer=20)
        from aaclip_b200.adapter import _infer_cfg
        from aaclip_b200.engine import weight_map
        from aaclip_b200.surgery import CLIPImageEncoder, _canonical_key
        enc = CLIPImageEncoder(ref_clip, [6, 12, 18, 24])
        assert enc.surgery_until_layer == 20
        keys = list(ref_clip.state_dict().keys())
        assert "visual.transformer.resblocks.5.attn.qkv.weight" in keys          # first replaced block (0-based 5)
        assert "visual.transformer.resblocks.4.attn.in_proj_weight" in keys      # last ordinary one
        canon = {"clip." + _canonical_key(k): v for k, v in ref_clip.state_dict().items()}
        cfg = _infer_cfg(ref_clip, [6, 12, 18, 24], 0, 0, 0.0, 0.0, False)
        assert (cfg.width, cfg.layers, cfg.heads, cfg.image_size) == (1024, 24, 16, 336)
        wm = weight_map(cfg, text=False)
        missing = [k for k in wm if k.startswith("clip.visual.") and k not in canon]
        assert not missing, missing[:5]
        assert tuple(canon["clip.visual.transformer.resblocks.23.attn.in_proj_weight"].shape) == (3072, 1024)
        assert tuple(canon["clip.visual.transformer.resblocks.23.attn.out_proj.bias"].shape) == (1024,)
        assert tuple(canon["clip.visual.proj"].shape) == (1024, 768)
        # an untouched CLIP is not mistaken for a surgery model
        with torch.device("meta"):
            plain = RefCLIP(**jcfg)
        assert CLIPImageEncoder(plain, []).surgery_until_layer is None
    finally:
        sys.path.remove(REFERENCE)
        for m in [m for m in sys.modules if m == "model" or m.startswith("model.")]:
            del sys.modules[m]


def test_surgery_encoder_refuses_cpu_inputs():
    from aaclip_b200.clip import CLIP
    from aaclip_b200.surgery import CLIPImageEncoder
    enc = CLIPImageEncoder(CLIP(synth.tiny_cfg(), text=False), [2, 4], surgery_until_layer=3)
    with pytest.raises(RuntimeError, match="B200 only"):
        enc.encode_image(torch.zeros(1, 3, 56, 56))
    with pytest.raises(RuntimeError, match="B200 only"):
        enc.patch_features(torch.zeros(1, 3, 56, 56))
    with pytest.raises(RuntimeError, match="B200 only"):
        CLIP(synth.tiny_cfg(), text=False).encode_image(torch.zeros(1, 3, 56, 56), [2, 4])
