"""GPU parity proper: the CUDA path, called through the C ABI (aaclip_b200.Engine / AdaptedCLIP / ops), against
the oracle on the same seeded inputs and against the golden vectors produced from the real reference.

Tolerances (north star + SURVEY D8; GEMM operands are bf16, residual stream / LN / softmax / head are fp32):
  * seg / det tokens are unit-scale cosines:          max-abs <= 1.5e-2 vs the fp32 oracle
  * min-max-normalised anomaly map (what metrics_eval consumes, forward_utils.py:241-244):  max-abs <= 1e-2
  * image-score ranking identical (argsort equality), score max-abs <= 2e-3
  * head kernels alone are fp32 end to end:           max-abs <= 2e-4 on O(1..10) maps
"""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
GOLD = os.path.join(HERE, "golden")

SEG_TOL, MAP_NORM_TOL, SCORE_TOL, HEAD_TOL = 1.5e-2, 1e-2, 2e-3, 2e-4


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def _mm(x):
    return (x - x.min()) / (x.max() - x.min())


@pytest.fixture(scope="module")
def full():
    """ViT-L/14-336 engine with the seed-0 synthetic weights the goldens were generated with."""
    from aaclip_b200 import synth
    from aaclip_b200.engine import Engine
    cfg = synth.VIT_L_14_336
    sd, ia, ta = synth.clip_state_dict(cfg, 0), synth.image_adapter_state_dict(cfg, 0), synth.text_adapter_state_dict(cfg, 0)
    eng = Engine(cfg, device=0, max_batch=4, max_text=16)
    used = eng.load_state_dicts(sd, ia, ta)
    assert len(used) == len(eng._wmap)
    yield cfg, eng, sd, ia, ta
    eng.close()


def test_visual_forward_vs_golden_and_oracle(full):
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    cfg, eng, sd, ia, _ = full
    g = _load("visual_vitl336_b2.pt")
    img = synth.images(g["batch"], cfg, seed=g["image_seed"])
    T = synth.anchors(cfg, seed=g["anchor_seed"])
    seg, det = eng.visual_forward(img.cuda())
    maps, scores = eng.forward_fused(img.cuda(), T.cuda(), "Industrial")
    maps_med, _ = eng.forward_fused(img.cuda(), T.cuda(), "Medical")
    torch.cuda.synchronize()
    # --- golden (real reference) subsamples
    for lvl, (s, ref) in enumerate(zip(seg, g["seg_sub"])):
        e = (s.cpu()[:, g["patch_idx"]] - ref).abs().max().item()
        print(f"[seg level {lvl}] max_abs_err vs reference golden = {e:.3e}")
        assert e < SEG_TOL
    e_det = (det.cpu() - g["det"]).abs().max().item()
    print(f"[det] max_abs_err = {e_det:.3e}")
    assert e_det < SEG_TOL
    assert (scores.cpu() - g["score"]).abs().max().item() < SCORE_TOL
    # --- full-resolution maps against the oracle (same weights, fp32 CPU)
    with torch.no_grad():
        seg_o, det_o = orc.visual_forward(sd, ia, img)
        map_o, score_o = orc.predict(seg_o, det_o, T, cfg.image_size, "Industrial")
        map_o_med, _ = orc.predict(seg_o, det_o, T, cfg.image_size, "Medical")
    raw = (maps.cpu() - map_o).abs().max().item()
    norm = (_mm(maps.cpu()) - _mm(map_o)).abs().max().item()
    norm_med = (_mm(maps_med.cpu()) - _mm(map_o_med)).abs().max().item()
    cos = max((a.cpu() - b).abs().max().item() for a, b in zip(seg, seg_o))
    print(f"[map] raw max_abs_err={raw:.3e} minmax-normalised={norm:.3e} (medical {norm_med:.3e}) cosine={cos:.3e} "
          f"score_err={(scores.cpu() - score_o).abs().max().item():.3e}")
    assert norm < MAP_NORM_TOL and norm_med < MAP_NORM_TOL
    assert (maps.cpu()[:, ::8, ::8] - g["map_industrial_sub"]).abs().max().item() <= raw + 1e-3
    assert torch.equal(scores.cpu().argsort(), score_o.argsort())


def test_fused_equals_dropin_path(full):
    """forward_fused (no seg materialisation) == AdaptedCLIP.forward + calculate_similarity_map x4 + sum."""
    from aaclip_b200 import ops, synth
    cfg, eng, *_ = full
    img = synth.images(3, cfg, seed=11).cuda()
    T = synth.anchors(cfg, seed=3).cuda()
    seg, det = eng.visual_forward(img)
    maps_a, scores_a = ops.anomaly_head(seg, T, cfg.image_size, ops.HEAD_TEST_INDUSTRIAL, det=det)
    maps_b, scores_b = eng.forward_fused(img, T, "Industrial")
    torch.cuda.synchronize()
    assert (maps_a - maps_b).abs().max().item() < 1e-4
    assert (scores_a - scores_b).abs().max().item() < 1e-6


def test_bf16_seg_tokens_and_fused_extrema(full):
    """aaclip_visual_forward's bf16 token output is the fp32 output rounded once (so the head can stream half the bytes),
    and the fused entry's extrema are exactly the (min, max) of the maps it wrote - on the device entry, the chunked
    path (B > max_batch) and the host pipeline."""
    from aaclip_b200 import ops, synth
    cfg, eng, *_ = full
    img = synth.images(6, cfg, seed=17).cuda()    # max_batch = 4: two chunks
    T = synth.anchors(cfg, seed=2).cuda()
    seg32, det32 = eng.visual_forward(img)
    seg16, det16 = eng.visual_forward(img, seg_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    assert all(t.dtype == torch.bfloat16 for t in seg16) and torch.equal(det32, det16)
    assert all(torch.equal(a.to(torch.bfloat16), b) for a, b in zip(seg32, seg16))
    ext = torch.empty(6, 2, device="cuda")
    maps, scores = eng.forward_fused(img, T, "Industrial", extrema=ext)
    torch.cuda.synchronize()
    flat = maps.flatten(1)
    assert torch.equal(ext[:, 0], flat.amin(1)) and torch.equal(ext[:, 1], flat.amax(1))
    # drop-in heads on both token dtypes agree with the fused entry (bf16 tokens: rounding of the cosines only)
    m32, s32 = ops.anomaly_head(seg32, T, cfg.image_size, ops.HEAD_TEST_INDUSTRIAL, det=det32)
    m16, _ = ops.anomaly_head(seg16, T, cfg.image_size, ops.HEAD_TEST_INDUSTRIAL)
    assert (m32 - maps).abs().max().item() < 1e-4 and (s32 - scores).abs().max().item() < 1e-6
    assert (_mm(m16) - _mm(maps)).abs().max().item() < MAP_NORM_TOL
    got = list(eng.predict_stream([img[:5].cpu(), img[5:].cpu()], T.cpu(), "Industrial", with_extrema=True))
    assert torch.equal(torch.cat([e for _, _, e in got]), ext.cpu())
    assert torch.equal(torch.cat([m for m, _, _ in got]), maps.cpu())


def test_batch_chunking_and_determinism(full):
    """B > max_batch is processed in chunks; results do not depend on the chunking or on batch neighbours."""
    from aaclip_b200 import synth
    cfg, eng, *_ = full
    img = synth.images(6, cfg, seed=21).cuda()   # max_batch = 4 -> chunks of 4 + 2
    T = synth.anchors(cfg, seed=1).cuda()
    m_all, s_all = eng.forward_fused(img, T)
    m_one, s_one = eng.forward_fused(img[4:5].contiguous(), T)
    m_again, _ = eng.forward_fused(img, T)
    torch.cuda.synchronize()
    assert torch.equal(m_all, m_again)
    assert (m_all[4:5] - m_one).abs().max().item() < 1e-4 and (s_all[4:5] - s_one).abs().max().item() < 1e-6
    m_empty, s_empty = eng.forward_fused(img[:0].contiguous(), T)
    assert m_empty.shape == (0, 336, 336) and s_empty.shape == (0,)


def test_host_buffer_entry(full):
    from aaclip_b200 import synth
    cfg, eng, *_ = full
    img = synth.images(2, cfg, seed=31).pin_memory()
    T = synth.anchors(cfg, seed=1)
    maps = torch.empty(2, 336, 336).pin_memory()
    scores = torch.empty(2).pin_memory()
    eng.forward_fused_host(img, T, maps, scores)
    m_dev, s_dev = eng.forward_fused(img.cuda(), T.cuda())
    torch.cuda.synchronize()
    assert torch.equal(maps, m_dev.cpu()) and torch.equal(scores, s_dev.cpu())


def test_host_pipeline_matches_device_entry(full):
    """submit_host / wait_host (two batches in flight), chunked forward_fused_host (B > max_batch) and
    predict_stream give bit-identical results to the device-pointer entry, in order."""
    from aaclip_b200 import synth
    cfg, eng, *_ = full
    T = synth.anchors(cfg, seed=1)
    batches = [synth.images(n, cfg, seed=40 + i) for i, n in enumerate((4, 1, 3, 4, 2))]
    ref = []
    for b in batches:
        m, s = eng.forward_fused(b.cuda(), T.cuda())
        ref.append((m.cpu(), s.cpu()))
    got = list(eng.predict_stream(batches, T))
    assert len(got) == len(ref)
    for (m, s), (mr, sr) in zip(got, ref):
        assert torch.equal(m, mr) and torch.equal(s, sr)
    # batches larger than max_batch (6 and 9 > 4) are split into chunks; an empty batch passes through
    big = [torch.cat([batches[0], batches[4]]), torch.cat([batches[3], batches[2], batches[4]]), batches[1][:0], batches[1]]
    got = list(eng.predict_stream(big, T))
    want = [(torch.cat([ref[0][0], ref[4][0]]), torch.cat([ref[0][1], ref[4][1]])),
            (torch.cat([ref[3][0], ref[2][0], ref[4][0]]), torch.cat([ref[3][1], ref[2][1], ref[4][1]])),
            (ref[1][0][:0], ref[1][1][:0]), ref[1]]
    assert len(got) == 4
    for (m, s), (mr, sr) in zip(got, want):
        assert m.shape == mr.shape and s.shape == sr.shape
    # chunk boundaries differ from the reference batches only where the composition differs: results are per image
    assert torch.equal(got[0][0], want[0][0]) and torch.equal(got[0][1], want[0][1])
    assert torch.equal(got[3][0], want[3][0])
    assert (got[1][0] - want[1][0]).abs().max() < 1e-5 and torch.equal(got[1][0][:4], want[1][0][:4])
    # a third submit without a wait is refused (two slots), then the pipeline drains normally
    bufs = [(b.pin_memory(), torch.empty(b.shape[0], 336, 336).pin_memory(), torch.empty(b.shape[0]).pin_memory())
            for b in batches[:3]]
    t0 = eng.submit_host(bufs[0][0], T, bufs[0][1], bufs[0][2])
    t1 = eng.submit_host(bufs[1][0], T, bufs[1][1], bufs[1][2])
    with pytest.raises(RuntimeError):
        eng.submit_host(bufs[2][0], T, bufs[2][1], bufs[2][2])
    eng.wait_host(t0)
    t2 = eng.submit_host(bufs[2][0], T, bufs[2][1], bufs[2][2])
    eng.wait_host(t1); eng.wait_host(t2)
    with pytest.raises(RuntimeError):
        eng.wait_host(t2)
    for (_, m, s), (mr, sr) in zip(bufs, ref):
        assert torch.equal(m, mr) and torch.equal(s, sr)
    # B = 6 > max_batch = 4 through the synchronous entry: chunks of 4 + 2 ride the same two slots
    big = torch.cat([batches[0], batches[4]]).pin_memory()
    maps, scores = torch.empty(6, 336, 336).pin_memory(), torch.empty(6).pin_memory()
    eng.forward_fused_host(big, T, maps, scores)
    assert torch.equal(maps, torch.cat([ref[0][0], ref[4][0]])) and torch.equal(scores, torch.cat([ref[0][1], ref[4][1]]))


def test_raw_image_pipeline_matches_reference_transform(full):
    """predict_stream fed RAW uint8 images [B,H0,W0,3]: the loader's transform_x (dataset/__init__.py:127-136) runs
    on the device inside the pipeline.  The result must be bit-identical to transforming the images with the oracle
    (pinned to PIL + torchvision) on the host and feeding the float batch - mixed raw sizes, mixed batch kinds."""
    import numpy as np
    import preprocess_oracle as po
    from aaclip_b200 import synth
    cfg, eng, *_ = full
    T = synth.anchors(cfg, seed=1)
    raw = [np.stack([po.synth_image(h, w, 70 + 10 * i + j) for j in range(n)])
           for i, (n, h, w) in enumerate([(2, 200, 260), (3, 400, 336), (1, 336, 336)])]
    floats = [torch.from_numpy(np.stack([po.transform_x(im, cfg.image_size) for im in b])).contiguous() for b in raw]
    ref = []
    for f in floats:
        m, s = eng.forward_fused(f.cuda(), T.cuda())
        ref.append((m.cpu(), s.cpu()))
    mixed = [torch.from_numpy(raw[0]), floats[1], torch.from_numpy(raw[2]), torch.from_numpy(raw[1])]
    want = [ref[0], ref[1], ref[2], ref[1]]
    got = list(eng.predict_stream(mixed, T))
    for (m, s), (mr, sr) in zip(got, want):
        assert torch.equal(m, mr) and torch.equal(s, sr)


def test_text_path_vs_golden(full):
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    from aaclip_b200._lib import check, cur_stream, load, ptr
    cfg, eng, sd, _, ta = full
    g = _load("text_vitl336.pt")
    emb = eng.text_forward(synth.tokens(6, cfg, seed=g["tok_synth_seed"]))
    torch.cuda.synchronize()
    ref = g["emb_synth"]
    rel = ((emb.cpu() - ref).abs().max() / ref.abs().max()).item()
    print(f"[text emb] rel_to_max err = {rel:.3e}")
    assert rel < 2e-2
    lib = load()
    for cls, toks in g["prompt_tokens"].items():
        anchor = torch.empty(768, 2, device="cuda")
        for col in (0, 1):
            e = eng.text_forward(toks[col])
            check(lib.aaclip_text_anchor(ptr(e), e.shape[0], e.shape[1], ptr(anchor), col, cur_stream()))
        torch.cuda.synchronize()
        err = (anchor.cpu() - g["anchors"][cls]).abs().max().item()
        cosine = (anchor.cpu() * g["anchors"][cls]).sum(0)
        print(f"[anchor {cls}] max_abs_err={err:.3e} cos={cosine.tolist()}")
        assert err < 5e-3 and (cosine > 0.999).all()


@pytest.mark.parametrize("name", ["head_g24_s336", "head_g37_s518", "head_g16_s100"])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_head_vs_golden(name, dtype):
    """aaclip_anomaly_head (the streaming kernel for the test modes, the general kernels for train mode) against the
    goldens made by the real reference (fp32 tokens) and, for bf16 tokens, against the oracle fed the SAME
    bf16-rounded tokens - both at the fp32 head tolerance: a wrong blur tap or edge weight cannot hide."""
    import aaclip_oracle as orc
    from aaclip_b200 import ops, synth
    g = _load(name + ".pt")
    cfg = synth.VIT_L_14_336
    feats, Tb, det = synth.head_inputs(g["batch"], g["grid"], cfg.embed_dim, 4, seed=g["feat_seed"])
    T = synth.anchors(cfg, seed=g["anchor_seed"])
    if dtype == "bf16":
        feats = [f.to(torch.bfloat16) for f in feats]
    fd = [f.cuda() for f in feats]
    for domain, mode in (("Industrial", ops.HEAD_TEST_INDUSTRIAL), ("Medical", ops.HEAD_TEST_MEDICAL)):
        maps, scores, ext = ops.anomaly_head(fd, T.cuda(), g["size"], mode, det=det.cuda(), want_extrema=True)
        torch.cuda.synchronize()
        if dtype == "f32":
            ref_sub = g["map_" + domain.lower()]
        else:   # the oracle on the same rounded tokens (forward_utils.py:196-216 restated, fp32 arithmetic)
            ref_sub = orc.predict([f.float() for f in feats], det, T, g["size"], domain)[0][:, ::7, ::7]
        err = (maps.cpu()[:, ::7, ::7] - ref_sub).abs().max().item()
        print(f"[head {name} {domain} {dtype}] max_abs_err={err:.3e}")
        assert err < HEAD_TOL
        assert (scores.cpu() - g["score"]).abs().max().item() < 1e-5
        # extrema come out of the kernel that writes the maps: exact against a reduction over what it wrote
        flat = maps.flatten(1)
        assert torch.equal(ext[:, 0], flat.amin(1)) and torch.equal(ext[:, 1], flat.amax(1))
        # the drop-in call sequence of test.py:89-93: one level per call, cat, sum
        per_level = [ops.anomaly_head([f], T.cuda(), g["size"], mode)[0] for f in fd]
        assert (torch.stack(per_level).sum(0) - maps).abs().max().item() < HEAD_TOL
    if dtype == "f32":
        tr, _ = ops.anomaly_head(fd, Tb.cuda(), g["size"], ops.HEAD_TRAIN_SOFTMAX)
        torch.cuda.synchronize()
        assert tr.shape == (4, g["batch"], 2, g["size"], g["size"])
        assert (tr.cpu()[:, :, :, ::7, ::7] - g["train_batched"]).abs().max().item() < HEAD_TOL
        assert (tr.sum(2) - 1).abs().max().item() < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,G,S,levels", [(1, 24, 336, 4), (3, 24, 336, 1), (67, 24, 336, 4), (5, 37, 518, 3),
                                          (2, 9, 63, 2), (300, 8, 32, 1)])
def test_head_stream_shapes_vs_oracle(dtype, B, G, S, levels):
    """The streaming head over ragged unit counts (P = 81 / 1369 are not multiples of the 64-patch unit), odd and
    even output widths (scalar / 8-byte / 16-byte store paths), one to four levels, fewer and (B = 300) far more
    units than resident CTAs - against the oracle on identical tokens."""
    import aaclip_oracle as orc
    from aaclip_b200 import ops
    gen = torch.Generator().manual_seed(1000 * B + G)
    feats = [torch.nn.functional.normalize(torch.randn(B, G * G, 768, generator=gen), dim=-1).to(dtype) for _ in range(levels)]
    det = torch.nn.functional.normalize(torch.randn(B, 768, generator=gen), dim=-1)
    T = torch.nn.functional.normalize(torch.randn(768, 2, generator=gen), dim=0)
    n_check = min(B, 6)
    pick = torch.linspace(0, B - 1, n_check).long()
    for domain, mode in (("Industrial", ops.HEAD_TEST_INDUSTRIAL), ("Medical", ops.HEAD_TEST_MEDICAL)):
        if G <= (3 if domain == "Industrial" else 4):
            continue
        maps, scores, ext = ops.anomaly_head([f.cuda() for f in feats], T.cuda(), S, mode, det=det.cuda(), want_extrema=True)
        torch.cuda.synchronize()
        map_o, score_o = orc.predict([f[pick].float() for f in feats], det[pick], T, S, domain)
        assert (maps.cpu()[pick] - map_o).abs().max().item() < HEAD_TOL
        assert (scores.cpu()[pick] - score_o).abs().max().item() < 1e-5
        flat = maps.flatten(1)
        assert torch.equal(ext[:, 0], flat.amin(1)) and torch.equal(ext[:, 1], flat.amax(1))
        # a second call on the same inputs is bit-identical (dynamic unit scheduling must not leak into the result)
        maps2, _ = ops.anomaly_head([f.cuda() for f in feats], T.cuda(), S, mode)
        assert torch.equal(maps, maps2)


def test_head_misaligned_and_unsupported_inputs_take_the_general_path():
    """Embed dims other than 768 cannot use the fixed-width streaming kernel: the general kernels serve them with the
    same results.  Level tensors that are not 16-byte aligned (a view into a larger buffer) are re-packed by the
    Python front end; the C entry refuses them with a message instead of faulting."""
    import aaclip_oracle as orc
    from aaclip_b200 import ops
    gen = torch.Generator().manual_seed(9)
    B, G, S = 2, 12, 84
    f = torch.nn.functional.normalize(torch.randn(B, G * G, 512, generator=gen), dim=-1)
    T = torch.nn.functional.normalize(torch.randn(512, 2, generator=gen), dim=0)
    maps, _, ext = ops.anomaly_head([f.cuda()], T.cuda(), S, ops.HEAD_TEST_INDUSTRIAL, want_extrema=True)
    map_o = orc.calculate_similarity_map(f, T, S, test=True, domain="Industrial")[:, 0]
    assert (maps.cpu() - map_o).abs().max().item() < HEAD_TOL
    assert torch.equal(ext[:, 1], maps.flatten(1).amax(1))
    f768 = torch.nn.functional.normalize(torch.randn(B, G * G, 768, generator=gen), dim=-1).to(torch.bfloat16)
    T768 = torch.nn.functional.normalize(torch.randn(768, 2, generator=gen), dim=0)
    big = torch.zeros(f768.numel() + 8, dtype=torch.bfloat16, device="cuda")
    view = big[1:1 + f768.numel()].view_as(f768)   # 2-byte offset: 16-byte alignment lost
    view.copy_(f768)
    assert view.data_ptr() % 16 != 0
    a, _ = ops.anomaly_head([view], T768.cuda(), S, ops.HEAD_TEST_MEDICAL)
    b, _ = ops.anomaly_head([f768.cuda()], T768.cuda(), S, ops.HEAD_TEST_MEDICAL)
    assert torch.equal(a, b)
    import ctypes as C
    from aaclip_b200 import _lib
    lib = _lib.load()
    arr = (C.c_void_p * 1)(view.data_ptr())
    ws = torch.empty(lib.aaclip_anomaly_head_workspace_bytes(1, B, G * G), dtype=torch.uint8, device="cuda")
    rc = lib.aaclip_anomaly_head(arr, 1, 1, T768.cuda().data_ptr(), 0, None, B, G * G, 768, S, 1, a.data_ptr(), None, None,
                                 ws.data_ptr(), ws.numel(), None)
    assert rc == -1 and b"aligned" in lib.aaclip_last_error()


def test_head_linearity_and_constant_property():
    """Size-independent properties at the full config: the test-mode map is affine in the per-patch dots, and a
    constant patch field maps to that constant (blur kernel sums to 1, reflect padding, bilinear partition)."""
    from aaclip_b200 import ops
    E, G, S, B = 768, 24, 336, 64
    T = torch.nn.functional.normalize(torch.randn(E, 2, generator=torch.Generator().manual_seed(1)), dim=0).cuda()
    f = torch.nn.functional.normalize(T[:, 1] - T[:, 0], dim=0)        # every patch identical
    feats = [f.expand(B, G * G, E).contiguous() for _ in range(4)]
    maps, _ = ops.anomaly_head(feats, T, S, ops.HEAD_TEST_INDUSTRIAL)
    torch.cuda.synchronize()
    d = (f @ T)
    expect = 4 * (100 * d[1] + 1 - 100 * d[0]) / 2
    assert (maps - expect).abs().max().item() < 1e-3 * abs(expect.item())


def test_dropin_module_surface():
    """The reference-shaped call sequence of test.py:153-176, 80-93 on the drop-in classes."""
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    from aaclip_b200.adapter import AdaptedCLIP
    from aaclip_b200.clip import CLIP
    from aaclip_b200.forward_utils import calculate_similarity_map, class_text_embedding
    cfg = synth.ModelCfg(layers=4, t_layers=2, image_adapt_until=2, text_adapt_until=1, levels=[1, 2, 3, 4])
    clip = CLIP(cfg)
    sd, ia, ta = synth.clip_state_dict(cfg, 3), synth.image_adapter_state_dict(cfg, 3), synth.text_adapter_state_dict(cfg, 3)
    clip.load_state_dict(sd, strict=True)
    model = AdaptedCLIP(clip_model=clip, text_adapt_weight=0.1, image_adapt_weight=0.1, text_adapt_until=1,
                        image_adapt_until=2, levels=[1, 2, 3, 4], relu=False, max_batch=2).to("cuda").eval()
    model.image_adapter.load_state_dict(ia)      # after construction, as test.py:175-176
    model.text_adapter.load_state_dict(ta)
    img = synth.images(3, cfg, seed=5)
    T = synth.anchors(cfg, seed=2)
    with torch.no_grad():
        patch_features, det_feature = model(img.cuda())
        pred = (det_feature @ T.cuda())
        per_level = [calculate_similarity_map(f, T.cuda(), 336, test=True, domain="Industrial") for f in patch_features]
        maps = torch.cat(per_level, dim=1).sum(1)
        seg_o, det_o = orc.visual_forward(sd, ia, img, layers=4, image_adapt_until=2, levels=(1, 2, 3, 4))
        map_o, _ = orc.predict(seg_o, det_o, T, 336, "Industrial")
    assert [tuple(t.shape) for t in patch_features] == [(3, 576, 768)] * 4 and det_feature.shape == (3, 768)
    assert per_level[0].shape == (3, 1, 336, 336)
    assert max((a.cpu() - b).abs().max().item() for a, b in zip(patch_features, seg_o)) < SEG_TOL
    assert (_mm(maps.cpu()) - _mm(map_o)).abs().max().item() < MAP_NORM_TOL
    # a re-loaded checkpoint must be picked up (version-counter resync)
    ia2 = synth.image_adapter_state_dict(cfg, 4)
    model.image_adapter.load_state_dict(ia2)
    with torch.no_grad():
        pf2, det2 = model(img.cuda())
        seg_o2, _ = orc.visual_forward(sd, ia2, img, layers=4, image_adapt_until=2, levels=(1, 2, 3, 4))
    assert max((a.cpu() - b).abs().max().item() for a, b in zip(pf2, seg_o2)) < SEG_TOL
    assert (pf2[0] - patch_features[0]).abs().max().item() > 1e-2
    # fused per-batch and pipelined-loop forms agree with the per-level form (same kernels up to the head entry)
    with torch.no_grad():
        maps2 = torch.cat([calculate_similarity_map(f, T.cuda(), 336, test=True, domain="Industrial") for f in pf2], 1).sum(1)
        m_f, s_f = model.predict(img.cuda(), T.cuda(), "Industrial")
        streamed = list(model.predict_stream([img[:2], img[2:]], T, "Industrial"))
    assert (m_f.cpu() - maps2.cpu()).abs().max().item() < 1e-3
    assert torch.equal(torch.cat([m for m, _ in streamed]), m_f.cpu())
    assert torch.equal(torch.cat([sc for _, sc in streamed]), s_f.cpu())
    assert (s_f.cpu() - ((det2 @ T.cuda())[:, 1].cpu() + 1) / 2).abs().max().item() < 1e-5
    # text anchors from token ids
    tn, tabn = synth.tokens(6, cfg, seed=7), synth.tokens(10, cfg, seed=8)
    anchor = class_text_embedding(model, tn.cuda(), tabn.cuda())
    with torch.no_grad():
        a_o = orc.class_text_anchor(orc.encode_text(sd, ta, tn, layers=2, text_adapt_until=1),
                                    orc.encode_text(sd, ta, tabn, layers=2, text_adapt_until=1))
    assert anchor.shape == (768, 2) and (anchor.cpu() - a_o).abs().max().item() < 5e-3


def test_full_batch64_properties_and_oracle_subset():
    """BASELINE.json configs[1] at full size (B=64, ViT-L/14-336, one chunk): size-independent properties - every
    image is an independent unit, so a permuted batch gives bit-identically permuted maps and scores - plus the
    oracle on a 3-image subset (the CPU oracle needs ~0.3 s per image) with the north-star tolerances."""
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    from aaclip_b200.engine import Engine
    cfg = synth.VIT_L_14_336
    sd, ia = synth.clip_state_dict(cfg, 0, text=False), synth.image_adapter_state_dict(cfg, 0)
    eng = Engine(cfg, device=0, max_batch=64, text=False)
    try:
        eng.load_state_dicts(sd, ia, None)
        img = synth.images(64, cfg, seed=77)
        T = synth.anchors(cfg, seed=1)
        maps, scores = eng.forward_fused(img.cuda(), T.cuda(), "Industrial")
        perm = torch.randperm(64, generator=torch.Generator().manual_seed(5))
        maps_p, scores_p = eng.forward_fused(img[perm].contiguous().cuda(), T.cuda(), "Industrial")
        torch.cuda.synchronize()
        assert torch.isfinite(maps).all() and torch.isfinite(scores).all()
        assert torch.equal(maps_p.cpu(), maps.cpu()[perm]) and torch.equal(scores_p.cpu(), scores.cpu()[perm])
        sub = [0, 31, 63]
        with torch.no_grad():
            seg_o, det_o = orc.visual_forward(sd, ia, img[sub])
            map_o, score_o = orc.predict(seg_o, det_o, T, cfg.image_size, "Industrial")
        m, s = maps.cpu()[sub], scores.cpu()[sub]
        err = max((_mm(m[i]) - _mm(map_o[i])).abs().max().item() for i in range(len(sub)))
        serr = (s - score_o).abs().max().item()
        print(f"[B=64] normalised-map max-abs = {err:.3e}  score max-abs = {serr:.3e}")
        assert err <= MAP_NORM_TOL and serr <= SCORE_TOL
        assert torch.equal(torch.argsort(s), torch.argsort(score_o))
    finally:
        eng.close()


def test_518px_operating_point_vs_oracle():
    """the paper's 518-px setting - 37x37+1 = 1370 tokens, head 37x37 -> 518x518 - with a 3-layer
    ViT-L-width model (L is a runtime parameter of every kernel; the full depth adds nothing new here)."""
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    from aaclip_b200.engine import Engine
    cfg = synth.ModelCfg(image_size=518, layers=3, t_layers=0, image_adapt_until=2, levels=[1, 2, 3])
    assert cfg.tokens == 1370
    sd, ia = synth.clip_state_dict(cfg, 0, text=False), synth.image_adapter_state_dict(cfg, 0)
    eng = Engine(cfg, device=0, max_batch=2, text=False)
    try:
        eng.load_state_dicts(sd, ia, None)
        img, T = synth.images(2, cfg, seed=9), synth.anchors(cfg, seed=1)
        seg, det = eng.visual_forward(img.cuda())
        maps, scores = eng.forward_fused(img.cuda(), T.cuda(), "Medical")
        torch.cuda.synchronize()
        with torch.no_grad():
            seg_o, det_o = orc.visual_forward(sd, ia, img, layers=3, image_adapt_until=2, levels=(1, 2, 3))
            map_o, score_o = orc.predict(seg_o, det_o, T, cfg.image_size, "Medical")
        e_seg = max((a.cpu() - b).abs().max().item() for a, b in zip(seg, seg_o))
        e_det = (det.cpu() - det_o).abs().max().item()
        e_map = max((_mm(maps.cpu()[i]) - _mm(map_o[i])).abs().max().item() for i in range(2))
        e_sc = (scores.cpu() - score_o).abs().max().item()
        print(f"[518px] seg {e_seg:.3e} det {e_det:.3e} normalised map {e_map:.3e} score {e_sc:.3e}")
        assert maps.shape == (2, 518, 518)
        assert e_seg <= SEG_TOL and e_det <= SEG_TOL and e_map <= MAP_NORM_TOL and e_sc <= SCORE_TOL
    finally:
        eng.close()


@pytest.mark.parametrize("relu,quick_gelu", [(True, False), (False, True)])
def test_relu_projections_and_quick_gelu_vs_oracle(relu, quick_gelu):
    """The two configuration switches the default bench does not exercise: AdaptedCLIP(relu=True) - the ctor default,
    LeakyReLU after seg_proj / det_proj (model/adapter_modules.py:16-26), in both the materialising and the fused
    (dots-in-the-epilogue) paths - and QuickGELU (model/transformer.py:46-49) in the c_fc epilogue."""
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    from aaclip_b200.engine import Engine
    cfg = synth.ModelCfg(layers=3, t_layers=0, image_adapt_until=2, levels=[1, 2, 3], relu=relu, quick_gelu=quick_gelu)
    sd, ia = synth.clip_state_dict(cfg, 5, text=False), synth.image_adapter_state_dict(cfg, 5)
    eng = Engine(cfg, device=0, max_batch=2, text=False)
    try:
        eng.load_state_dicts(sd, ia, None)
        img, T = synth.images(2, cfg, seed=13), synth.anchors(cfg, seed=4)
        seg, det = eng.visual_forward(img.cuda())
        maps, scores = eng.forward_fused(img.cuda(), T.cuda(), "Industrial")
        torch.cuda.synchronize()
        with torch.no_grad():
            seg_o, det_o = orc.visual_forward(sd, ia, img, layers=3, image_adapt_until=2, levels=(1, 2, 3),
                                              quick_gelu=quick_gelu)
            map_o, score_o = orc.predict(seg_o, det_o, T, cfg.image_size, "Industrial")
        e_seg = max((a.cpu() - b).abs().max().item() for a, b in zip(seg, seg_o))
        e_map = max((_mm(maps.cpu()[i]) - _mm(map_o[i])).abs().max().item() for i in range(2))
        e_sc = (scores.cpu() - score_o).abs().max().item()
        print(f"[relu={relu} quick_gelu={quick_gelu}] seg {e_seg:.3e} normalised map {e_map:.3e} score {e_sc:.3e}")
        assert e_seg <= SEG_TOL and (det.cpu() - det_o).abs().max().item() <= SEG_TOL
        assert e_map <= MAP_NORM_TOL and e_sc <= SCORE_TOL
    finally:
        eng.close()


def test_ln_fold_and_separate_layernorm_schedules_agree():
    """The two schedules of the visual tower - LayerNorm folded into the in_proj / c_fc GEMMs (default) and separate
    LayerNorm kernels - are the same function up to bf16 rounding: both must sit inside the oracle tolerances and
    within 2x the tolerance of each other; weights re-uploaded after creation must re-fold."""
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    from aaclip_b200.engine import Engine
    cfg = synth.VIT_L_14_336
    sd, ia = synth.clip_state_dict(cfg, 0), synth.image_adapter_state_dict(cfg, 0)
    img = synth.images(2, cfg, seed=77)
    T = synth.anchors(cfg, seed=1)
    with torch.no_grad():
        seg_o, det_o = orc.visual_forward(sd, ia, img)
        map_o, score_o = orc.predict(seg_o, det_o, T, cfg.image_size, "Industrial")
    outs = {}
    for fold in (True, False):
        eng = Engine(cfg, device=0, max_batch=2, text=False, ln_fold=fold)
        # first a DIFFERENT seed, then the real weights: the folded copies must follow the second upload
        eng.load_state_dicts(synth.clip_state_dict(cfg, 5), synth.image_adapter_state_dict(cfg, 5), None)
        eng.forward_fused(img.cuda(), T.cuda())
        eng.load_state_dicts(sd, ia, None)
        seg, det = eng.visual_forward(img.cuda())
        maps, scores = eng.forward_fused(img.cuda(), T.cuda())
        torch.cuda.synchronize()
        outs[fold] = (torch.stack(seg).cpu(), det.cpu(), maps.cpu(), scores.cpu())
        eng.close()
        e_seg = (outs[fold][0] - torch.stack(seg_o)).abs().max().item()
        e_map = (_mm(outs[fold][2]) - _mm(map_o)).abs().max().item()
        print(f"[ln_fold={fold}] seg {e_seg:.3e} normalised map {e_map:.3e} score {(outs[fold][3] - score_o).abs().max().item():.3e}")
        assert e_seg < SEG_TOL and e_map < MAP_NORM_TOL and (outs[fold][3] - score_o).abs().max() < SCORE_TOL
    assert (outs[True][0] - outs[False][0]).abs().max() < 2 * SEG_TOL
    assert (_mm(outs[True][2]) - _mm(outs[False][2])).abs().max() < 2 * MAP_NORM_TOL


def test_cuda_graph_replay_is_bit_identical_and_follows_weight_updates(full):
    """On a non-default stream a repeated (batch, pointers) call is captured into a CUDA graph at its second occurrence
    and replayed afterwards: replays must be bit-identical to the eager launches, keep counting launches, and pick up
    re-uploaded weights (the graph reads the same weight buffers; the fold kernels run outside it)."""
    from aaclip_b200 import synth
    cfg, eng, sd, ia, ta = full
    T = synth.anchors(cfg, seed=1).cuda()
    img = synth.images(3, cfg, seed=91).cuda()
    eager = eng.forward_fused(img, T)                       # default stream: always eager
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    out = (torch.empty(3, 336, 336, device="cuda"), torch.empty(3, device="cuda"))
    counts = []
    with torch.cuda.stream(side):
        for _ in range(5):                                   # eager, capture + launch, replay x3
            n0 = eng.launch_count
            eng.forward_fused(img, T, out=out)
            side.synchronize()
            counts.append(eng.launch_count - n0)
            assert torch.equal(out[0], eager[0]) and torch.equal(out[1], eager[1])
        assert len(set(counts)) == 1 and counts[0] > 50      # a replay counts the kernels of the graph
        # different weights through the same buffers: the replayed graph must compute with them
        eng.load_state_dicts(synth.clip_state_dict(cfg, 3), synth.image_adapter_state_dict(cfg, 3), None)
        eng.forward_fused(img, T, out=out)
        side.synchronize()
        changed = out[0].clone()
    eager2 = eng.forward_fused(img, T)
    torch.cuda.synchronize()
    assert torch.equal(changed, eager2[0]) and not torch.equal(changed, eager[0])
    eng.load_state_dicts(sd, ia, None)                       # restore for the other tests of the module
    with torch.cuda.stream(side):
        eng.forward_fused(img, T, out=out)
        side.synchronize()
    assert torch.equal(out[0], eager[0])


@pytest.mark.parametrize("offset", [3.0, -8.0])
def test_ln_fold_with_large_row_mean_vs_oracle(offset):
    """Stress for the folded LayerNorm: a large ln_pre bias gives every row of the residual stream a mean several times
    its spread, the case in which rstd * (x W'^T - mean * colsum) cancels two big terms.  Both schedules must stay inside
    the oracle tolerances (4-layer ViT-L-width model)."""
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    from aaclip_b200.engine import Engine
    cfg = synth.ModelCfg(layers=4, t_layers=0, image_adapt_until=2, levels=[1, 2, 3, 4])
    sd, ia = synth.clip_state_dict(cfg, 0, text=False), synth.image_adapter_state_dict(cfg, 0)
    sd = dict(sd)
    sd["visual.ln_pre.bias"] = sd["visual.ln_pre.bias"] + offset
    img, T = synth.images(2, cfg, seed=13), synth.anchors(cfg, seed=1)
    with torch.no_grad():
        seg_o, det_o = orc.visual_forward(sd, ia, img, layers=4, image_adapt_until=2, levels=(1, 2, 3, 4))
        map_o, score_o = orc.predict(seg_o, det_o, T, cfg.image_size, "Industrial")
    for fold in (True, False):
        eng = Engine(cfg, device=0, max_batch=2, text=False, ln_fold=fold)
        try:
            eng.load_state_dicts(sd, ia, None)
            seg, det = eng.visual_forward(img.cuda())
            maps, scores = eng.forward_fused(img.cuda(), T.cuda())
            torch.cuda.synchronize()
            e_seg = max((a.cpu() - b).abs().max().item() for a, b in zip(seg, seg_o))
            e_map = (_mm(maps.cpu()) - _mm(map_o)).abs().max().item()
            e_sc = (scores.cpu() - score_o).abs().max().item()
            print(f"[row mean {offset:+.0f}, ln_fold={fold}] seg {e_seg:.3e} normalised map {e_map:.3e} score {e_sc:.3e}")
            assert e_seg <= SEG_TOL and e_map <= MAP_NORM_TOL and e_sc <= SCORE_TOL
        finally:
            eng.close()


@pytest.mark.parametrize("scale", [60.0])
def test_outlier_channels_full_depth_both_ln_schedules_vs_oracle(scale):
    """Trained CLIP ViT-L carries a few 'massive activation' channels, tens of times larger than the rest of the residual
    stream from the first blocks to the last.  Random-init statistics do not: emulate them - three channels get an ln_pre
    gain of x`scale`, an ln_pre offset and a c_proj bias of +-`scale` in blocks 1 and 2 - and run the FULL 24-layer model.
    The folded schedule feeds bf16(x) (not bf16(LN(x))) to the tensor core and takes the variance as E[x^2] - mean^2 from
    per-slice sums; the separate schedule normalises first.  Both must hold the north-star tolerances against the fp32
    oracle, and the image ranking must be identical."""
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    from aaclip_b200.engine import Engine
    cfg = synth.VIT_L_14_336
    sd, ia = dict(synth.clip_state_dict(cfg, 0, text=False)), synth.image_adapter_state_dict(cfg, 0)
    ch = torch.tensor([7, 300, 777])
    sign = torch.tensor([1.0, -1.0, 1.0])
    for key in ("visual.ln_pre.weight", "visual.ln_pre.bias", "visual.transformer.resblocks.1.mlp.c_proj.bias",
                "visual.transformer.resblocks.2.mlp.c_proj.bias"):
        sd[key] = sd[key].clone()
    sd["visual.ln_pre.weight"][ch] *= scale
    sd["visual.ln_pre.bias"][ch] += 0.5 * scale * sign
    sd["visual.transformer.resblocks.1.mlp.c_proj.bias"][ch] += scale * sign
    sd["visual.transformer.resblocks.2.mlp.c_proj.bias"][ch] -= 0.5 * scale * sign
    img, T = synth.images(3, cfg, seed=29), synth.anchors(cfg, seed=1)
    with torch.no_grad():
        seg_o, det_o = orc.visual_forward(sd, ia, img)
        map_o, score_o = orc.predict(seg_o, det_o, T, cfg.image_size, "Industrial")
    for fold in (True, False):
        eng = Engine(cfg, device=0, max_batch=3, text=False, ln_fold=fold)
        try:
            eng.load_state_dicts(sd, ia, None)
            seg, det = eng.visual_forward(img.cuda())
            maps, scores = eng.forward_fused(img.cuda(), T.cuda())
            torch.cuda.synchronize()
            e_seg = max((a.cpu() - b).abs().max().item() for a, b in zip(seg, seg_o))
            e_det = (det.cpu() - det_o).abs().max().item()
            e_map = (_mm(maps.cpu()) - _mm(map_o)).abs().max().item()
            e_sc = (scores.cpu() - score_o).abs().max().item()
            print(f"[outlier channels x{scale:.0f}, ln_fold={fold}] seg {e_seg:.3e} det {e_det:.3e} normalised map {e_map:.3e} "
                  f"score {e_sc:.3e}")
            assert e_seg <= SEG_TOL and e_det <= SEG_TOL and e_map <= MAP_NORM_TOL and e_sc <= SCORE_TOL
            assert torch.equal(scores.cpu().argsort(), score_o.argsort())
        finally:
            eng.close()


def test_openai_checkpoint_through_the_drop_in(full):
    """(f)3 end to end: a ViT-L/14-336 checkpoint in the OpenAI archive's layout (fp16 tensors, metadata scalars) ->
    load_openai_state_dict -> AdaptedCLIP -> maps, against the engine loaded directly with the same values (the fp16
    tensors widened to fp32): bit-identical, because both paths hand the engine the same fp32 numbers."""
    from aaclip_b200 import synth
    from aaclip_b200.adapter import AdaptedCLIP
    from aaclip_b200.clip import load_openai_state_dict
    from aaclip_b200.engine import Engine
    cfg, _, _, ia, _ = full
    sd16 = synth.openai_style_state_dict(cfg, 5)
    clip = load_openai_state_dict(sd16)
    model = AdaptedCLIP(clip_model=clip, relu=False, max_batch=2).to("cuda").eval()
    model.image_adapter.load_state_dict(ia)
    img, T = synth.images(2, cfg, seed=3).cuda(), synth.anchors(cfg, seed=1).cuda()
    maps, scores = model.predict(img, T, "Industrial")
    eng = Engine(cfg, device=0, max_batch=2, text=False)
    try:
        eng.load_state_dicts({k: (v.float() if torch.is_floating_point(v) else v) for k, v in sd16.items()}, ia, None)
        m2, s2 = eng.forward_fused(img, T)
        torch.cuda.synchronize()
        assert torch.equal(maps, m2) and torch.equal(scores, s2)
    finally:
        eng.close()
        model._engine.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """Contexts on two GPUs of one process (the > 48 KB shared-memory opt-in is per device context, and a
    stream handle is only valid on its own device): same weights, same inputs, same results, with the thread's current
    device left where the caller had it."""
    from aaclip_b200 import ops, synth
    from aaclip_b200.engine import Engine
    cfg = synth.ModelCfg(layers=3, t_layers=0, image_adapt_until=2, levels=[1, 2, 3])
    sd, ia = synth.clip_state_dict(cfg, 0, text=False), synth.image_adapter_state_dict(cfg, 0)
    img, T = synth.images(2, cfg, seed=1), synth.anchors(cfg, seed=1)
    torch.cuda.set_device(0)
    outs = []
    for dev in (0, 1):
        eng = Engine(cfg, device=dev, max_batch=2, text=False)
        eng.load_state_dicts(sd, ia, None)
        assert torch.cuda.current_device() == 0
        seg, det = eng.visual_forward(img.to(f"cuda:{dev}"))
        maps, scores = eng.forward_fused(img.to(f"cuda:{dev}"), T.to(f"cuda:{dev}"))
        head, _ = ops.anomaly_head(seg, T.to(f"cuda:{dev}"), cfg.image_size, ops.HEAD_TEST_INDUSTRIAL)
        torch.cuda.synchronize(dev)
        assert torch.cuda.current_device() == 0
        outs.append((maps.cpu(), scores.cpu(), head.cpu()))
        eng.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])


def test_caller_side_stream_capture_joins_the_callers_graph(full):
    """when the CALLER is capturing the stream (torch.cuda.graph around the forward) the library must not
    start a nested capture of its own: its launches join the caller's graph, and replaying that graph reproduces the
    eager result bit for bit - for the fused entry (with extrema) and for the drop-in head."""
    from aaclip_b200 import ops, synth
    cfg, eng, *_ = full
    img = synth.images(2, cfg, seed=41).cuda()
    T = synth.anchors(cfg, seed=4).cuda()
    ext_e = torch.empty(2, 2, device="cuda")
    maps_e, scores_e = eng.forward_fused(img, T, extrema=ext_e)          # eager (also makes sure the weights are folded)
    seg, det = eng.visual_forward(img, seg_dtype=torch.bfloat16)
    head_e, hs_e = ops.anomaly_head(seg, T, cfg.image_size, ops.HEAD_TEST_MEDICAL, det=det)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    out = (torch.empty_like(maps_e), torch.empty_like(scores_e))
    ext = torch.empty_like(ext_e)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        eng.forward_fused(img, T, out=out, extrema=ext)                  # first sight of this key on this stream: eager
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            eng.forward_fused(img, T, out=out, extrema=ext)              # second sight: would be the library's own capture
            head_g, hs_g = ops.anomaly_head(seg, T, cfg.image_size, ops.HEAD_TEST_MEDICAL, det=det)
        out[0].zero_(); out[1].zero_(); ext.zero_(); head_g.zero_()
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out[0], maps_e) and torch.equal(out[1], scores_e) and torch.equal(ext, ext_e)
    assert torch.equal(head_g, head_e) and torch.equal(hs_g, hs_e)
    # the library's own graph cache still works afterwards on the same key
    with torch.cuda.stream(side):
        for _ in range(3):
            eng.forward_fused(img, T, out=out, extrema=ext)
        side.synchronize()
    assert torch.equal(out[0], maps_e)


def test_attributes_changed_after_the_first_forward_take_effect():
    """The reference reads `i_w`, `levels` and `image_adapt_until` on every forward (model/adapter.py:90-101): changing them
    on a model that has already run must change the result exactly as a freshly built model would (the context and its
    cached graphs are rebuilt)."""
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    from aaclip_b200.adapter import AdaptedCLIP
    from aaclip_b200.clip import CLIP
    cfg = synth.ModelCfg(layers=4, t_layers=0, image_adapt_until=2, levels=[1, 2, 3, 4])
    clip = CLIP(cfg, text=False)
    sd, ia = synth.clip_state_dict(cfg, 3, text=False), synth.image_adapter_state_dict(cfg, 3)
    clip.load_state_dict(sd, strict=False)
    model = AdaptedCLIP(clip_model=clip, image_adapt_until=2, levels=[1, 2, 3, 4], relu=False, max_batch=2).to("cuda").eval()
    model.image_adapter.load_state_dict(ia)
    img = synth.images(2, cfg, seed=5)
    T = synth.anchors(cfg, seed=2).cuda()
    with torch.no_grad():
        first, _ = model(img.cuda())
        for _ in range(3):
            model.predict(img.cuda(), T, "Industrial")          # primes the fused-forward graph cache
        model.i_w = 0.3
        model.levels = [4, 2]                                    # membership semantics: taps after blocks 2 and 4
        seg, det = model(img.cuda())
        maps, _ = model.predict(img.cuda(), T, "Industrial")
        ia_sub = dict(ia)
        seg_o, det_o = orc.visual_forward(sd, ia_sub, img, layers=4, image_adapt_until=2, image_adapt_weight=0.3, levels=(2, 4))
        map_o, _ = orc.predict(seg_o, det_o, T.cpu(), 336, "Industrial")
    assert len(seg) == 2
    assert max((a.cpu() - b).abs().max().item() for a, b in zip(seg, seg_o)) < SEG_TOL
    assert (det.cpu() - det_o).abs().max().item() < SEG_TOL
    assert (_mm(maps.cpu()) - _mm(map_o)).abs().max().item() < MAP_NORM_TOL
    assert (seg[0] - first[1]).abs().max().item() > 1e-3       # block 2's tap moved with the adapter weight


def test_plain_clip_encode_text_vs_golden_and_anchor_path():
    """`CLIP.encode_text` of the container = the un-adapted text path (model/model.py:189-200; what
    `AdaptedCLIP.encode_text(adapt_text=False)` delegates to and what test.py:197-200 builds anchors from), on a text-only
    engine context: against the golden of the real reference, normalised form, re-upload on a changed parameter."""
    import aaclip_oracle as orc
    from aaclip_b200 import synth
    from aaclip_b200.adapter import AdaptedCLIP
    from aaclip_b200.clip import CLIP
    cfg = synth.VIT_L_14_336
    g = _load("text_plain_vitl336.pt")
    sd = synth.clip_state_dict(cfg, 0)
    clip = CLIP(cfg)
    clip.load_state_dict(sd, strict=True)
    clip = clip.cuda()
    tok = synth.tokens(6, cfg, seed=g["tok_synth_seed"]).cuda()
    emb = clip.encode_text(tok)
    torch.cuda.synchronize()
    ref = g["emb_plain"]
    rel = ((emb.cpu() - ref).abs().max() / ref.abs().max()).item()
    print(f"[plain text emb] rel_to_max err = {rel:.3e}")
    assert tuple(emb.shape) == (6, 768) and rel < 2e-2
    n = clip.encode_text(tok, normalize=True)
    assert (n.norm(dim=-1) - 1).abs().max().item() < 1e-5
    # AdaptedCLIP.encode_text(adapt_text=False) delegates to it (model/adapter.py:115-116)
    model = AdaptedCLIP(clip_model=clip, relu=False).to("cuda").eval()
    assert torch.equal(model.encode_text(tok, adapt_text=False), emb)
    # a changed parameter is picked up
    with torch.no_grad():
        clip.text_projection.mul_(0.5)
        half = clip.encode_text(tok)
    assert (half - 0.5 * emb).abs().max().item() <= 2e-2 * emb.abs().max().item()
    with pytest.raises(RuntimeError, match="B200 only"):
        clip.encode_text(tok.cpu())
