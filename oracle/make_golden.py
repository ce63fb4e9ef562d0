"""Generate tests/golden/*.pt by running the REAL reference (imported read-only from /root/reference) and
check the oracle restatement (oracle/aaclip_oracle.py) against it in the same run.

Run in the build container only (the reference cannot travel to the GPU box):

    python oracle/make_golden.py

Stubs (none of them on the arithmetic path we pin, except where stated):
  ipdb   : imported, never called (model/transformer.py:10, model/adapter_modules.py:2, forward_utils.py:9)
  ftfy   : fix_text = identity, exact for the ASCII prompt tables (model/tokenizer.py:18,63)
  kornia : un-vendored dependency, absent offline.  kornia.filters.gaussian_blur2d is bound to the oracle's
           restatement, so the reference's calculate_similarity_map runs its own code around it and the
           blur itself stays "parity unpinned (kornia)".
  cv2 / sklearn / pandas are imported by forward_utils.py for metrics/visualisation only; stubbed if absent.
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("AACLIP_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import aaclip_oracle as orc  # noqa: E402
from aaclip_b200 import synth  # noqa: E402


def _stub(name: str, **attrs):
    try:
        importlib.import_module(name)
        return
    except Exception:
        pass
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    _stub("ipdb")
    _stub("ftfy", fix_text=lambda t: t)
    k = types.ModuleType("kornia")
    kf = types.ModuleType("kornia.filters")
    kf.gaussian_blur2d = orc.gaussian_blur2d
    k.filters = kf
    sys.modules["kornia"] = k
    sys.modules["kornia.filters"] = kf
    _stub("cv2")
    sys.path.insert(0, REF)


def build_reference(cfg: synth.ModelCfg, seed: int):
    from model.adapter import AdaptedCLIP
    from model.model import CLIP

    jcfg = json.load(open(os.path.join(REF, "model/model_configs/ViT-L-14-336.json")))
    clip = CLIP(**jcfg)
    sd = synth.clip_state_dict(cfg, seed)
    missing, unexpected = clip.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(m == "attn_mask" for m in missing), missing
    model = AdaptedCLIP(clip, cfg.text_adapt_weight, cfg.image_adapt_weight, cfg.text_adapt_until,
                        cfg.image_adapt_until, list(cfg.levels), cfg.relu).eval()
    ia, ta = synth.image_adapter_state_dict(cfg, seed), synth.text_adapter_state_dict(cfg, seed)
    model.image_adapter.load_state_dict(ia, strict=True)
    model.text_adapter.load_state_dict(ta, strict=True)
    return model, sd, ia, ta


def maxdiff(a, b):
    return float((a - b).abs().max())


def main():
    torch.set_num_threads(os.cpu_count() or 4)
    install_stubs()
    import forward_utils as ref_fu
    from dataset.constants import CLASS_NAMES, PROMPTS, REAL_NAMES
    from model.tokenizer import tokenize

    cfg = synth.VIT_L_14_336
    seed = 0
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    model, sd, ia, ta = build_reference(cfg, seed)
    report = {}

    # ---------------------------------------------------------------- visual + head (config 1/2 shape, B=2)
    B = 2
    img = synth.images(B, cfg, seed=1)
    T = synth.anchors(cfg, seed=1)
    with torch.no_grad():
        seg_ref, det_ref = model(img)
        score_ref = ((det_ref @ T)[:, 1] + 1) / 2                                          # test.py:83-84
        maps_ref = {}
        for domain in ("Industrial", "Medical"):
            per_level = [ref_fu.calculate_similarity_map(f, T, cfg.image_size, test=True, domain=domain)
                         for f in seg_ref]
            maps_ref[domain] = torch.cat(per_level, dim=1).sum(1)                          # test.py:93
        train_ref = ref_fu.calculate_similarity_map(seg_ref[0], T, cfg.image_size, test=False)
        seg_o, det_o = orc.visual_forward(sd, ia, img)
        map_o, score_o = orc.predict(seg_o, det_o, T, cfg.image_size, "Industrial")
    report["visual_seg_maxdiff"] = max(maxdiff(a, b) for a, b in zip(seg_ref, seg_o))
    report["visual_det_maxdiff"] = maxdiff(det_ref, det_o)
    report["map_maxdiff"] = maxdiff(maps_ref["Industrial"], map_o)
    report["score_maxdiff"] = maxdiff(score_ref, score_o)
    patch_idx = torch.arange(0, cfg.patches, 37)
    torch.save({
        "cfg": "ViT-L-14-336", "seed": seed, "image_seed": 1, "anchor_seed": 1, "batch": B,
        "patch_idx": patch_idx,
        "seg_sub": [s[:, patch_idx].clone() for s in seg_ref],         # 4 x [2,16,768]
        "det": det_ref.clone(),                                          # [2,768]
        "score": score_ref.clone(),                                      # [2]
        "map_industrial_sub": maps_ref["Industrial"][:, ::8, ::8].clone(),   # [2,42,42]
        "map_medical_sub": maps_ref["Medical"][:, ::8, ::8].clone(),
        "map_industrial_stats": torch.stack([maps_ref["Industrial"].amin((1, 2)), maps_ref["Industrial"].amax((1, 2)),
                                              maps_ref["Industrial"].mean((1, 2))], 1),
        "train_l0_sub": train_ref[:, :, ::8, ::8].clone(),                # [2,2,42,42]
    }, os.path.join(out_dir, "visual_vitl336_b2.pt"))

    # ---------------------------------------------------------------- head alone on synthetic tokens
    for idx, (bh, grid, size, name) in enumerate([(3, 24, 336, "head_g24_s336"), (2, 37, 518, "head_g37_s518"),
                                                  (1, 16, 100, "head_g16_s100")]):
        feats, Tb, det = synth.head_inputs(bh, grid, cfg.embed_dim, 4, seed=5 + idx)
        rec = {"feat_seed": 5 + idx, "grid": grid, "size": size, "batch": bh, "anchor_seed": 1}
        with torch.no_grad():
            for domain in ("Industrial", "Medical"):
                m = torch.cat([ref_fu.calculate_similarity_map(f, T, size, test=True, domain=domain) for f in feats], 1).sum(1)
                mo = torch.cat([orc.calculate_similarity_map(f, T, size, test=True, domain=domain) for f in feats], 1).sum(1)
                report[f"{name}_{domain}_maxdiff"] = maxdiff(m, mo)
                rec["map_" + domain.lower()] = m[:, ::7, ::7].clone()
            tr = torch.stack([ref_fu.calculate_similarity_map(f, Tb, size, test=False) for f in feats], 0)
            rec["train_batched"] = tr[:, :, :, ::7, ::7].clone()
            rec["score"] = ((det @ T)[:, 1] + 1) / 2
        torch.save(rec, os.path.join(out_dir, name + ".pt"))

    # ---------------------------------------------------------------- text path (config 4)
    with torch.no_grad():
        tok_synth = synth.tokens(6, cfg, seed=2)
        emb_ref = model.encode_text(tok_synth)
        emb_o = orc.encode_text(sd, ta, tok_synth)
        report["text_synth_maxdiff"] = maxdiff(emb_ref, emb_o)
        # real MVTec prompts through the reference tokenizer (forward_utils.py:138-162)
        prompts = {}
        anchors_ref = {}
        for cls in ("bottle", "screw"):
            real = REAL_NAMES["MVTec"][cls]
            toks = []
            for states in (PROMPTS["prompt_normal"], PROMPTS["prompt_abnormal"]):
                sentences = [t.format(s.format(real)) for s in states for t in PROMPTS["prompt_templates"]]
                toks.append(tokenize(sentences))
            prompts[cls] = toks
            anchors_ref[cls] = ref_fu.get_adapted_single_class_text_embedding(model, "MVTec", cls, "cpu")
            a_o = orc.class_text_anchor(orc.encode_text(sd, ta, toks[0]), orc.encode_text(sd, ta, toks[1]))
            report[f"anchor_{cls}_maxdiff"] = maxdiff(anchors_ref[cls], a_o)
    # the un-adapted text path (model/model.py:189-200; test.py:197-200 when no text adapter is used)
    with torch.no_grad():
        plain_ref = model.clipmodel.encode_text(tok_synth)
        plain_o = orc.clip_encode_text(sd, tok_synth)
        assert torch.equal(model.encode_text(tok_synth, adapt_text=False), plain_ref)   # model/adapter.py:115-116
    report["text_plain_maxdiff"] = maxdiff(plain_ref, plain_o)
    torch.save({"tok_synth_seed": 2, "emb_plain": plain_ref.clone()}, os.path.join(out_dir, "text_plain_vitl336.pt"))
    torch.save({"tok_synth_seed": 2, "emb_synth": emb_ref.clone(),
                "prompt_tokens": prompts, "anchors": anchors_ref,
                "class_names_mvtec": list(CLASS_NAMES["MVTec"])},
               os.path.join(out_dir, "text_vitl336.pt"))

    # ---- checkpoint ingestion: positional-embedding rescale 24x24 -> 37x37 (model/model.py:395-426), the 518-px setting
    from model.model import resize_pos_embed as ref_resize
    from aaclip_b200.clip import resize_pos_embed as our_resize
    g = torch.Generator().manual_seed(123)
    old = torch.randn(1 + 24 * 24, 64, generator=g)

    class _V:
        grid_size = (37, 37)

    class _M:
        visual = _V()

    sd_ref, sd_our = {"visual.positional_embedding": old.clone()}, {"visual.positional_embedding": old.clone()}
    ref_resize(sd_ref, _M())
    our_resize(sd_our, 37)
    report["pos_embed_resize_maxdiff"] = maxdiff(sd_ref["visual.positional_embedding"], sd_our["visual.positional_embedding"])
    torch.save({"seed": 123, "old_grid": 24, "new_grid": 37, "width": 64,
                "resized": sd_ref["visual.positional_embedding"].clone()}, os.path.join(out_dir, "pos_embed_24_to_37.pt"))

    # ---- checkpoint ingestion: an OpenAI-layout fp16 state dict through the reference's own loading chain
    #      (model/openai.py:66-77 -> model/model.py:311-368 -> model/clip.py:112-131), at a different target resolution
    import hashlib
    from model.model import CLIP as RefCLIP, build_model_from_openai_state_dict
    from aaclip_b200.clip import load_openai_state_dict
    tiny = synth.ModelCfg(**synth.OPENAI_TINY)
    target = 98   # 7 x 7 grid from the checkpoint's 4 x 4
    pre = build_model_from_openai_state_dict(synth.openai_style_state_dict(tiny, 7)).float()   # load_openai_model, fp32
    sd_pre = pre.state_dict()
    ref_model = RefCLIP(embed_dim=tiny.embed_dim,
                        vision_cfg={"image_size": target, "layers": tiny.layers, "width": tiny.width, "patch_size": tiny.patch_size},
                        text_cfg={"context_length": tiny.t_context, "vocab_size": tiny.t_vocab, "width": tiny.t_width,
                                  "heads": tiny.t_heads, "layers": tiny.t_layers})
    ref_resize(sd_pre, ref_model)
    ref_model.load_state_dict(sd_pre, strict=True)
    ref_sd = {k: v.detach().float().contiguous() for k, v in ref_model.state_dict().items()}
    ours = load_openai_state_dict(synth.openai_style_state_dict(tiny, 7), img_size=target)
    our_sd = {k: v.detach().float().contiguous() for k, v in ours.state_dict().items()}
    assert set(ref_sd) == set(our_sd), set(ref_sd) ^ set(our_sd)
    report["openai_ingest_maxdiff"] = max(maxdiff(ref_sd[k], our_sd[k]) for k in ref_sd)
    torch.save({"cfg": synth.OPENAI_TINY, "seed": 7, "target_image_size": target,
                "sha256": {k: hashlib.sha256(v.numpy().tobytes()).hexdigest() for k, v in ref_sd.items()},
                "visual.positional_embedding": ref_sd["visual.positional_embedding"].clone()},
               os.path.join(out_dir, "openai_ingest_tiny.pt"))

    # ---- stage-1 feature extraction of train.py:74-85: the surgery CLIP (DAPM_replace, v-v attention
    #      over the batch for the last 19 blocks) + the class feature of the unmodified CLIP.  B = 3 so that the batch
    #      coupling is exercised with an odd batch.
    import copy
    Bs = 3
    img_s = synth.images(Bs, cfg, seed=4)
    levels = [6, 12, 18, 24]
    clip_plain = model.clipmodel
    clip_surgery = copy.deepcopy(clip_plain).eval()
    clip_surgery.visual.DAPM_replace(DPAM_layer=20)                                         # train.py:243
    with torch.no_grad():
        pooled_s, toks_s = clip_surgery.encode_image(img_s, levels)                          # train.py:75
        cls_token, _ = clip_plain.encode_image(img_s, [])                                    # :76
        cls_n = cls_token / cls_token.norm(dim=-1, keepdim=True)                             # :77
        feats = [clip_surgery.visual.ln_post(t[:, 1:, :]) for t in toks_s]                   # :78-80
        feats = [t @ clip_surgery.visual.proj for t in feats]                                # :81
        feats = [t / t.norm(dim=-1, keepdim=True) for t in feats]                            # :82-84
        feats = [t + cls_n.unsqueeze(1) for t in feats]                                      # :85
        pooled_o, toks_o = orc.encode_image(sd, img_s, levels, surgery_until_layer=20)
        cls_o, _ = orc.encode_image(sd, img_s, [])
        feats_o = orc.surgery_patch_features(sd, sd, img_s, levels=levels, surgery_until_layer=20)
    report["surgery_tokens_rel_maxdiff"] = max(maxdiff(a, b) / float(a.abs().max()) for a, b in zip(toks_s, toks_o))
    report["surgery_pooled_maxdiff"] = maxdiff(pooled_s, pooled_o) / float(pooled_s.abs().max())
    report["plain_pooled_maxdiff"] = maxdiff(cls_token, cls_o) / float(cls_token.abs().max())
    report["surgery_features_maxdiff"] = max(maxdiff(a, b) for a, b in zip(feats, feats_o))
    tok_idx = torch.arange(0, cfg.tokens, 41)
    torch.save({
        "cfg": "ViT-L-14-336", "seed": seed, "image_seed": 4, "batch": Bs, "levels": levels, "surgery_until_layer": 20,
        "patch_idx": patch_idx, "token_idx": tok_idx,
        "features_sub": [f[:, patch_idx].clone() for f in feats],         # 4 x [3,16,768]  (train.py:85)
        "tokens_sub": [t[:, tok_idx].clone() for t in toks_s],            # 4 x [3,15,1024] (encode_image out_tokens)
        "pooled_surgery": pooled_s.clone(),                               # [3,768]
        "pooled_plain": cls_token.clone(),                                # [3,768]
    }, os.path.join(out_dir, "surgery_vitl336_b3.pt"))

    json.dump(report, open(os.path.join(out_dir, "oracle_vs_reference.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(report, indent=1, sort_keys=True))
    bad = {k: v for k, v in report.items() if not v < 2e-4}
    if bad:
        raise SystemExit(f"oracle deviates from the reference: {bad}")


def F_normalize(t, dim=-1):
    return torch.nn.functional.normalize(t, dim=dim)


if __name__ == "__main__":
    main()
