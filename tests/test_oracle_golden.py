"""CPU: the oracle restatement (oracle/aaclip_oracle.py) against the golden vectors that
oracle/make_golden.py produced by running the REAL reference modules from /root/reference.

The reference ships no golden vectors of its own; these files are the pin.  Tolerance 2e-4 on
unit-norm features / O(1) maps: the only freedom is fp32 summation order inside the BLAS the host uses.
"""
import json
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import aaclip_oracle as orc  # noqa: E402
from aaclip_b200 import synth  # noqa: E402

GOLD = os.path.join(HERE, "golden")
TOL = 2e-4


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


@pytest.fixture(scope="module")
def weights():
    cfg = synth.VIT_L_14_336
    return cfg, synth.clip_state_dict(cfg, 0), synth.image_adapter_state_dict(cfg, 0), synth.text_adapter_state_dict(cfg, 0)


def test_reference_agreement_report():
    rep = json.load(open(os.path.join(GOLD, "oracle_vs_reference.json")))
    assert rep and all(v < TOL for v in rep.values()), rep


def test_visual_and_head_vs_reference_golden(weights):
    cfg, sd, ia, _ = weights
    g = _load("visual_vitl336_b2.pt")
    img = synth.images(g["batch"], cfg, seed=g["image_seed"])
    T = synth.anchors(cfg, seed=g["anchor_seed"])
    with torch.no_grad():
        seg, det = orc.visual_forward(sd, ia, img)
        maps_i, score = orc.predict(seg, det, T, cfg.image_size, "Industrial")
        maps_m, _ = orc.predict(seg, det, T, cfg.image_size, "Medical")
        train0 = orc.calculate_similarity_map(seg[0], T, cfg.image_size, test=False)
    for s, ref in zip(seg, g["seg_sub"]):
        assert (s[:, g["patch_idx"]] - ref).abs().max() < TOL
    assert (det - g["det"]).abs().max() < TOL
    assert (score - g["score"]).abs().max() < TOL
    assert (maps_i[:, ::8, ::8] - g["map_industrial_sub"]).abs().max() < 5e-4
    assert (maps_m[:, ::8, ::8] - g["map_medical_sub"]).abs().max() < 5e-4
    assert (train0[:, :, ::8, ::8] - g["train_l0_sub"]).abs().max() < 5e-4


@pytest.mark.parametrize("name", ["head_g24_s336", "head_g37_s518", "head_g16_s100"])
def test_head_vs_reference_golden(name):
    g = _load(name + ".pt")
    cfg = synth.VIT_L_14_336
    feats, Tb, det = synth.head_inputs(g["batch"], g["grid"], cfg.embed_dim, 4, seed=g["feat_seed"])
    T = synth.anchors(cfg, seed=g["anchor_seed"])
    for domain in ("Industrial", "Medical"):
        m = torch.cat([orc.calculate_similarity_map(f, T, g["size"], test=True, domain=domain) for f in feats], 1).sum(1)
        assert (m[:, ::7, ::7] - g["map_" + domain.lower()]).abs().max() < TOL
    tr = torch.stack([orc.calculate_similarity_map(f, Tb, g["size"], test=False) for f in feats], 0)
    assert (tr[:, :, :, ::7, ::7] - g["train_batched"]).abs().max() < TOL
    assert (((det @ T)[:, 1] + 1) / 2 - g["score"]).abs().max() < 1e-6


def test_text_vs_reference_golden(weights):
    cfg, sd, _, ta = weights
    g = _load("text_vitl336.pt")
    with torch.no_grad():
        emb = orc.encode_text(sd, ta, synth.tokens(6, cfg, seed=g["tok_synth_seed"]))
        assert (emb - g["emb_synth"]).abs().max() < TOL
        for cls, toks in g["prompt_tokens"].items():
            a = orc.class_text_anchor(orc.encode_text(sd, ta, toks[0]), orc.encode_text(sd, ta, toks[1]))
            assert a.shape == (768, 2)
            assert (a - g["anchors"][cls]).abs().max() < TOL


def test_gaussian_kernel_properties():
    """kornia restatement ("parity unpinned"): the properties its published algorithm guarantees."""
    for k, s in ((7, 1.0), (9, 1.5)):
        w = orc.gaussian_kernel1d(k, s)
        assert abs(float(w.sum()) - 1) < 1e-6 and torch.equal(w, w.flip(0)) and int(w.argmax()) == k // 2
    x = torch.full((1, 1, 24, 24), 3.25)
    assert (orc.gaussian_blur2d(x, (7, 7), (1.0, 1.0)) - 3.25).abs().max() < 1e-6  # constants are preserved


def test_gaussian_blur_against_independent_implementation():
    """kornia==0.6.9 is absent offline, so its gaussian_blur2d stays "parity unpinned" - but the restated algorithm
    (normalised exp(-x^2 / 2 sigma^2) window of k taps, reflect padding without edge repeat, separable) is exactly what
    scipy.ndimage.gaussian_filter(mode="mirror", radius=k//2) computes: an independent cross-check of the restatement
    for both operating points of forward_utils.py:205-210."""
    ndi = pytest.importorskip("scipy.ndimage")
    import numpy as np
    x = torch.randn(3, 1, 24, 24, generator=torch.Generator().manual_seed(5))
    for k, s in ((7, 1.0), (9, 1.5)):
        got = orc.gaussian_blur2d(x, (k, k), (s, s)).numpy()
        want = np.stack([ndi.gaussian_filter(x[i, 0].numpy().astype(np.float64), sigma=s, radius=k // 2, mode="mirror")
                         for i in range(3)])[:, None]
        assert np.abs(got - want).max() < 2e-6


def test_surgery_feature_extractor_vs_reference_golden(weights):
    """train.py:74-85 with DAPM_replace(20) on the real reference CLIP (oracle/make_golden.py) vs the restatement."""
    cfg, sd, _, _ = weights
    g = _load("surgery_vitl336_b3.pt")
    img = synth.images(g["batch"], cfg, seed=g["image_seed"])
    with torch.no_grad():
        pooled, toks = orc.encode_image(sd, img, g["levels"], surgery_until_layer=g["surgery_until_layer"])
        pooled_plain, none = orc.encode_image(sd, img, [])
        feats = orc.surgery_patch_features(sd, sd, img, levels=g["levels"], surgery_until_layer=g["surgery_until_layer"])
    assert none == []
    for t, ref in zip(toks, g["tokens_sub"]):
        assert (t[:, g["token_idx"]] - ref).abs().max() < TOL * ref.abs().max()
    assert (pooled - g["pooled_surgery"]).abs().max() < TOL * g["pooled_surgery"].abs().max()
    assert (pooled_plain - g["pooled_plain"]).abs().max() < TOL * g["pooled_plain"].abs().max()
    for f, ref in zip(feats, g["features_sub"]):
        assert (f[:, g["patch_idx"]] - ref).abs().max() < TOL


def test_vv_attention_properties():
    """The restated `Attention.forward` (model/transformer.py:123-152): one image -> the out-projected value itself;
    the images of a batch are coupled; permuting the batch permutes the result."""
    torch.manual_seed(0)
    D, heads, L = 128, 2, 5
    sd = {"a.in_proj_weight": torch.randn(3 * D, D) * D ** -0.5, "a.in_proj_bias": torch.randn(3 * D) * 0.1,
          "a.out_proj.weight": torch.randn(D, D) * D ** -0.5, "a.out_proj.bias": torch.randn(D) * 0.1}
    x = torch.randn(4, L, D)
    out = orc.vv_attention(x, sd, "a.", heads)
    v1 = torch.nn.functional.linear(x[:1], sd["a.in_proj_weight"], sd["a.in_proj_bias"])[..., 2 * D:]
    one = orc.vv_attention(x[:1], sd, "a.", heads)
    assert torch.allclose(one, torch.nn.functional.linear(v1, sd["a.out_proj.weight"], sd["a.out_proj.bias"]), atol=1e-5)
    assert (out[:1] - one).abs().max() > 1e-3
    perm = torch.tensor([2, 0, 3, 1])
    assert torch.allclose(orc.vv_attention(x[perm], sd, "a.", heads), out[perm], atol=1e-5)


def test_plain_clip_encode_text_vs_reference_golden(weights):
    """The un-adapted CLIP.encode_text (model/model.py:189-200) of the real reference vs the restatement."""
    cfg, sd, _, _ = weights
    g = _load("text_plain_vitl336.pt")
    with torch.no_grad():
        emb = orc.clip_encode_text(sd, synth.tokens(6, cfg, seed=g["tok_synth_seed"]))
    assert emb.shape == g["emb_plain"].shape == (6, 768)
    assert (emb - g["emb_plain"]).abs().max() < TOL * g["emb_plain"].abs().max().clamp_min(1.0)
