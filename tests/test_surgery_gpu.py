"""GPU parity of the stage-1 feature extractor of train.py: v-v attention kernel, encode_image, and the
train.py:74-85 sequence, through the C ABI, against the oracle and the golden vectors made from the real reference
(`DAPM_replace(20)` on the reference's CLIP, oracle/make_golden.py).

Tolerances (GEMM operands bf16, residual stream fp32 - the same budget as tests/test_model_gpu.py):
  * v-v attention kernel alone: bf16 output, max-abs <= 2^-8 * max|v| vs the fp32 formula on the same bf16 inputs
  * projected, normalised patch features (unit vectors):   max-abs <= 1.5e-2, cosine >= 0.999
  * residual-stream tokens: max-abs <= 2e-2 of the level's max |token|; pooled feature: 2e-2 of its max
"""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
GOLD = os.path.join(HERE, "golden")
FEAT_TOL, TOKEN_REL_TOL = 1.5e-2, 2e-2


def _vv_ref(v, B, L, heads):
    """softmax_over_images(v v^T / 8) v per (token, head), fp32, on [B*L, heads*64] rows (row = b*L + l)."""
    x = v.float().view(B, L, heads, 64).permute(1, 2, 0, 3)          # [L, h, B, 64]
    a = torch.softmax((x @ x.transpose(-1, -2)) * 0.125, dim=-1)
    return (a @ x).permute(2, 0, 1, 3).reshape(B * L, heads * 64)


@pytest.mark.parametrize("B", [1, 2, 3, 5, 8, 9, 33, 64, 128])
def test_vv_attention_kernel_vs_formula(B):
    from aaclip_b200 import ops
    L, heads = (577, 16) if B <= 3 else (29, 16)
    g = torch.Generator().manual_seed(100 + B)
    # correlated images: a shared component per (token, head) so that the off-diagonal attention weights matter
    base = torch.randn(1, L, heads * 64, generator=g)
    v = (0.8 * base + 0.6 * torch.randn(B, L, heads * 64, generator=g)).reshape(B * L, heads * 64)
    v = v.to(torch.bfloat16).cuda()
    out = ops.vv_attention(v, B, L, heads)
    torch.cuda.synchronize()
    ref = _vv_ref(v, B, L, heads)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2.0 ** -8 * v.float().abs().max().item(), err
    if B == 1:
        assert torch.equal(out, v)          # one image: softmax over a single key is 1, the output is v itself
    else:
        # the attention is not the identity here: the mixing term is what is being tested
        assert (ref - v.float()).abs().max().item() > 0.1


def test_vv_attention_couples_the_batch_and_is_permutation_equivariant():
    from aaclip_b200 import ops
    B, L, heads = 5, 40, 4
    g = torch.Generator().manual_seed(7)
    base = torch.randn(1, L, heads * 64, generator=g)
    v3 = (0.8 * base + 0.6 * torch.randn(B, L, heads * 64, generator=g)).to(torch.bfloat16).cuda()
    out = ops.vv_attention(v3.reshape(B * L, -1), B, L, heads).view(B, L, -1)
    perm = torch.tensor([3, 0, 4, 1, 2], device="cuda")
    out_p = ops.vv_attention(v3[perm].reshape(B * L, -1).contiguous(), B, L, heads).view(B, L, -1)
    assert (out_p.float() - out[perm].float()).abs().max().item() <= 2.0 ** -7 * v3.float().abs().max().item()
    # dropping an image changes the others: the operator is batch-coupled (model/transformer.py:123-152 reads [L, batch, D]
    # as (B, N, C))
    out_4 = ops.vv_attention(v3[:4].reshape(4 * L, -1).contiguous(), 4, L, heads).view(4, L, -1)
    assert (out_4.float() - out[:4].float()).abs().max().item() > 0.05


def test_vv_attention_rejects_what_it_cannot_do():
    from aaclip_b200 import ops
    v = torch.zeros(129 * 2, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError, match="couples the images"):
        ops.vv_attention(v, 129, 2, 1)
    with pytest.raises(ValueError):
        ops.vv_attention(v, 2, 2, 1)


def test_add_image_vector():
    from aaclip_b200 import ops
    g = torch.Generator().manual_seed(3)
    t = torch.randn(3, 50, 768, generator=g).cuda()
    c = torch.randn(3, 768, generator=g).cuda()
    want = t + c.unsqueeze(1)
    got = ops.add_image_vector(t.clone(), c)
    assert torch.equal(got, want)


# ------------------------------------------------------------------------------------------------ tiny configuration
def _tiny_models(seed=3, **kw):
    from aaclip_b200 import synth
    from aaclip_b200.clip import CLIP
    cfg = synth.tiny_cfg(**kw)
    sd = synth.clip_state_dict(cfg, seed, text=False)
    m = CLIP(cfg, text=False)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    return cfg, sd, m.cuda()


@pytest.mark.parametrize("dpam", [None, 2, 3, 5])
@pytest.mark.parametrize("B", [1, 2, 5])
def test_tiny_encode_image_and_features_vs_oracle(dpam, B):
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    from aaclip_b200.surgery import CLIPImageEncoder, surgery_patch_features
    cfg, sd, model = _tiny_models()
    levels = [1, 2, 4]
    enc = CLIPImageEncoder(model, levels, surgery_until_layer=dpam, max_batch=8)
    plain = CLIPImageEncoder(model, [], max_batch=8)
    img = synth.images(B, cfg, seed=11)
    kw = dict(patch_size=cfg.patch_size, heads=cfg.heads, layers=cfg.layers)
    with torch.no_grad():
        pooled_o, toks_o = orc.encode_image(sd, img, levels, surgery_until_layer=dpam, **kw)
        feats_o = orc.surgery_patch_features(sd, sd, img, levels=levels, surgery_until_layer=dpam or 0, **kw)
    pooled, toks = enc.encode_image(img.cuda())
    pooled_n, none = enc.encode_image(img.cuda(), [], normalize=True)
    feats = surgery_patch_features(enc, plain, img.cuda())
    torch.cuda.synchronize()
    assert none == [] and len(toks) == len(levels) == len(feats)
    for t, to in zip(toks, toks_o):
        assert tuple(t.shape) == (B, cfg.tokens, cfg.width)
        assert (t.cpu() - to).abs().max().item() <= TOKEN_REL_TOL * to.abs().max().item()
    assert (pooled.cpu() - pooled_o).abs().max().item() <= TOKEN_REL_TOL * pooled_o.abs().max().item()
    pn = torch.nn.functional.normalize(pooled_o, dim=-1)
    assert (pooled_n.cpu() - pn).abs().max().item() <= FEAT_TOL
    for f, fo in zip(feats, feats_o):
        assert (f.cpu() - fo).abs().max().item() <= 2 * FEAT_TOL      # two unit vectors added
    # a subset of the out_layers, in block order whatever the order asked for (Transformer.forward :304-316)
    _, sub = enc.encode_image(img.cuda(), [4, 1])
    assert len(sub) == 2 and torch.equal(sub[0], toks[0]) and torch.equal(sub[1], toks[2])
    with pytest.raises(ValueError, match="not requested"):
        enc.encode_image(img.cuda(), [3])


def test_tiny_surgery_both_layernorm_schedules_agree_with_the_oracle(monkeypatch):
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    from aaclip_b200.engine import Engine
    cfg, sd, _ = _tiny_models()
    cfg.levels, cfg.image_adapt_until, cfg.relu, cfg.t_layers = [2, 4], 0, False, 0
    img = synth.images(3, cfg, seed=12)
    with torch.no_grad():
        _, toks_o = orc.encode_image(sd, img, cfg.levels, surgery_until_layer=4, patch_size=cfg.patch_size, heads=cfg.heads,
                                     layers=cfg.layers)
    for fold in (True, False):
        eng = Engine(cfg, device=0, max_batch=4, text=False, ln_fold=fold)
        eng.load_state_dicts(sd, {f"seg_proj.{i}.fc.weight": sd["visual.proj"].t().contiguous() for i in range(2)}
                             | {"det_proj.fc.weight": sd["visual.proj"].t().contiguous()})
        eng.dapm_replace(4)
        _, toks = eng.encode_image(img.cuda(), want_pooled=False)
        torch.cuda.synchronize()
        for t, to in zip(toks, toks_o):
            assert (t.cpu() - to).abs().max().item() <= TOKEN_REL_TOL * to.abs().max().item(), fold
        eng.close()


def test_surgery_context_never_splits_a_batch_and_validates_depth():
    from aaclip_b200 import synth
    from aaclip_b200.surgery import CLIPImageEncoder
    cfg, sd, model = _tiny_models()
    enc = CLIPImageEncoder(model, [4], surgery_until_layer=3, max_batch=2)
    img = synth.images(3, cfg, seed=1).cuda()
    with pytest.raises(RuntimeError, match="couples the images"):
        enc.encode_image(img)
    with pytest.raises(RuntimeError, match="couples the images"):
        enc.patch_features(img)
    enc.DAPM_replace(None)                       # ordinary attention: chunks of max_batch are fine again
    pooled, toks = enc.encode_image(img)
    assert tuple(toks[0].shape) == (3, cfg.tokens, cfg.width)
    whole = CLIPImageEncoder(model, [4], max_batch=4)            # the same batch in one chunk: images are independent again
    pooled_w, toks_w = whole.encode_image(img)
    assert torch.equal(pooled, pooled_w) and torch.equal(toks[0], toks_w[0])
    feats, feats_w = enc.patch_features(img), whole.patch_features(img)
    assert torch.equal(feats[0], feats_w[0])
    p0, t0 = whole.encode_image(img[:0].contiguous())             # an empty batch passes through
    assert tuple(p0.shape) == (0, cfg.embed_dim) and tuple(t0[0].shape) == (0, cfg.tokens, cfg.width)
    whole.DAPM_replace(3)
    p0, t0 = whole.encode_image(img[:0].contiguous())
    assert tuple(p0.shape) == (0, cfg.embed_dim) and tuple(t0[0].shape) == (0, cfg.tokens, cfg.width)
    with pytest.raises(RuntimeError, match="reaches past"):
        enc.DAPM_replace(cfg.layers + 2)         # the reference indexes resblocks[-i] out of range here


# ------------------------------------------------------------------------------------------------ ViT-L/14-336, golden
def test_vitl_surgery_vs_golden_from_the_real_reference():
    from aaclip_b200 import synth
    from aaclip_b200.clip import CLIP
    from aaclip_b200.surgery import CLIPImageEncoder, surgery_patch_features
    g = torch.load(os.path.join(GOLD, "surgery_vitl336_b3.pt"), weights_only=False)
    cfg = synth.VIT_L_14_336
    sd = synth.clip_state_dict(cfg, g["seed"], text=False)
    model = CLIP(cfg, text=False)
    model.load_state_dict(sd, strict=False)
    model = model.cuda()
    enc = CLIPImageEncoder(model, g["levels"], surgery_until_layer=g["surgery_until_layer"], max_batch=4)
    plain = CLIPImageEncoder(model, [], max_batch=4)
    img = synth.images(g["batch"], cfg, seed=g["image_seed"]).cuda()
    pooled, toks = enc.encode_image(img)
    pooled_plain, _ = plain.encode_image(img, [])
    feats = surgery_patch_features(enc, plain, img)
    torch.cuda.synchronize()
    for lvl, (t, ref) in enumerate(zip(toks, g["tokens_sub"])):
        e = (t.cpu()[:, g["token_idx"]] - ref).abs().max().item()
        assert e <= TOKEN_REL_TOL * ref.abs().max().item(), (lvl, e, ref.abs().max().item())
    for p, ref in ((pooled, g["pooled_surgery"]), (pooled_plain, g["pooled_plain"])):
        assert (p.cpu() - ref).abs().max().item() <= TOKEN_REL_TOL * ref.abs().max().item()
    for lvl, (f, ref) in enumerate(zip(feats, g["features_sub"])):
        got = f.cpu()[:, g["patch_idx"]]
        assert (got - ref).abs().max().item() <= 2 * FEAT_TOL, lvl
        cos = torch.nn.functional.cosine_similarity(got, ref, dim=-1).min().item()
        assert cos >= 0.999, (lvl, cos)


def test_train_py_lines_75_to_85_run_unchanged_on_the_container():
    """The literal call sequence of train.py:75-85 (and :243) on aaclip_b200.clip.CLIP: `visual.DAPM_replace`,
    `encode_image(image, out_layers)` on the surgery model and on the un-modified one, the rest in the caller's torch ops."""
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    from aaclip_b200.clip import CLIP
    cfg, sd, clip_surgery = _tiny_models()
    clip_plain = CLIP(cfg, text=False)
    clip_plain.load_state_dict(sd, strict=False)
    clip_plain = clip_plain.cuda()
    clip_surgery.visual.DAPM_replace(DPAM_layer=3)                                       # train.py:243
    image = synth.images(2, cfg, seed=21).cuda()
    levels = [1, 2, 4]
    with torch.no_grad():
        _, patch_features = clip_surgery.encode_image(image, levels)                      # :75
        cls_token, _ = clip_plain.encode_image(image, [])                                 # :76
        cls_token = cls_token / cls_token.norm(dim=-1, keepdim=True)                      # :77
        patch_features = [clip_surgery.visual.ln_post(t[:, 1:, :]) for t in patch_features]   # :78-80
        patch_features = [t @ clip_surgery.visual.proj for t in patch_features]           # :81
        patch_features = [t / t.norm(dim=-1, keepdim=True) for t in patch_features]       # :82-84
        patch_features = [t + cls_token.unsqueeze(1) for t in patch_features]             # :85
        want = orc.surgery_patch_features(sd, sd, image.cpu(), levels=levels, surgery_until_layer=3, patch_size=cfg.patch_size,
                                          heads=cfg.heads, layers=cfg.layers)
    assert len(patch_features) == 3
    for f, fo in zip(patch_features, want):
        assert (f.cpu() - fo).abs().max().item() <= 2 * FEAT_TOL
    # the un-modified model did not pick the surgery up, and switching it off again restores ordinary attention
    with torch.no_grad():
        _, plain_tokens = clip_plain.encode_image(image, levels)
        clip_surgery.visual.DAPM_replace(None)
        _, back = clip_surgery.encode_image(image, levels)
    assert all(torch.equal(a, b) for a, b in zip(plain_tokens, back))
    with pytest.raises(IndexError):
        clip_surgery.visual.DAPM_replace(cfg.layers + 2)


def test_vitl_surgery_at_train_py_defaults_518px_batch_2_vs_oracle():
    """train.py's own operating point (--img_size 518, --image_batch_size 2, --surgery_until_layer 20; train.py:186-199):
    1370 tokens per image, v-v attention over a batch of two in the last 19 blocks, against the oracle on the host."""
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    from aaclip_b200.clip import CLIP
    from aaclip_b200.surgery import CLIPImageEncoder, surgery_patch_features
    cfg = synth.ModelCfg(image_size=518)
    sd = synth.clip_state_dict(cfg, 0, text=False)
    model = CLIP(cfg, text=False)
    model.load_state_dict(sd, strict=False)
    model = model.cuda()
    levels = [6, 12, 18, 24]
    enc = CLIPImageEncoder(model, levels, surgery_until_layer=20, max_batch=2)
    plain = CLIPImageEncoder(model, [], max_batch=2)
    img = synth.images(2, cfg, seed=6)
    feats = surgery_patch_features(enc, plain, img.cuda())
    torch.cuda.synchronize()
    with torch.no_grad():
        want = orc.surgery_patch_features(sd, sd, img, levels=levels, surgery_until_layer=20)
    for lvl, (f, fo) in enumerate(zip(feats, want)):
        assert tuple(f.shape) == (2, 37 * 37, 768)
        assert (f.cpu() - fo).abs().max().item() <= 2 * FEAT_TOL, lvl
        assert torch.nn.functional.cosine_similarity(f.cpu(), fo, dim=-1).min().item() >= 0.999, lvl
