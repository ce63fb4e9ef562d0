"""The loader-side image transform (dataset/__init__.py:127-136: Resize BICUBIC + ToTensor + Normalize).

CPU: oracle/preprocess_oracle.py (numpy restatement of Pillow's fixed-point resample) against the golden vectors
     oracle/make_preprocess_golden.py took from PIL + torchvision, and against PIL itself where it is importable.
GPU: the CUDA kernels through the C ABI against the oracle and the golden hashes - BIT-EXACT (integer work for the
     resize; the float normalisation is IEEE division / subtraction in torch's order).
"""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import preprocess_oracle as po  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "preprocess_pil.npz"))
CASES = [tuple(int(v) for v in c) for c in GOLD["cases"]]


def _sha(a: np.ndarray) -> np.ndarray:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_to_%d" % c[:3])
def test_oracle_vs_pil_golden(case):
    h, w, s, seed = case
    tag = f"{h}x{w}_{s}_{seed}"
    img = po.synth_image(h, w, seed)
    u8 = po.resize_bicubic_u8(img, s)
    f32 = po.transform_x(img, s)
    assert np.array_equal(_sha(u8), GOLD[tag + "_sha_u8"])
    assert np.array_equal(_sha(f32), GOLD[tag + "_sha_f32"])
    assert np.array_equal(f32[:, ::13, ::11].view(np.uint32), GOLD[tag + "_sub_f32"].view(np.uint32))
    if tag + "_u8" in GOLD:
        assert np.array_equal(u8, GOLD[tag + "_u8"])


def test_oracle_vs_pil_live():
    Image = pytest.importorskip("PIL.Image")
    for (h, w, s, seed) in [(97, 131, 50, 11), (64, 64, 200, 12), (300, 41, 41, 13)]:
        img = po.synth_image(h, w, seed)
        ref = np.asarray(Image.fromarray(img).resize((s, s), Image.BICUBIC))
        assert np.array_equal(po.resize_bicubic_u8(img, s), ref)


def test_coefficient_tables():
    # every row of fixed-point weights sums to 2^22 within rounding; windows stay inside the image
    for n_in, n_out in [(1024, 336), (90, 336), (336, 336 * 3), (700, 518)]:
        ksize, bounds, kk = po.precompute_coeffs(n_in, n_out)
        assert kk.shape == (n_out, ksize)
        assert np.abs(kk.sum(1) - (1 << po.PRECISION_BITS)).max() <= ksize
        assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= n_in).all() and (bounds[:, 1] >= 1).all()


def test_constant_image_is_preserved():
    img = np.full((50, 70, 3), 201, np.uint8)
    assert (po.resize_bicubic_u8(img, 33) == 201).all()


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_to_%d" % c[:3])
def test_gpu_transform_bit_exact(case):
    from aaclip_b200 import forward_utils, ops
    h, w, s, seed = case
    tag = f"{h}x{w}_{s}_{seed}"
    imgs = np.stack([po.synth_image(h, w, seed), po.synth_image(h, w, seed + 100)])
    d = torch.from_numpy(imgs).cuda()
    u8 = ops.resize_bicubic_u8(d, s).cpu().numpy()
    f32 = forward_utils.transform_x(d, s).cpu().numpy()
    assert np.array_equal(_sha(u8[0]), GOLD[tag + "_sha_u8"])              # PIL's bytes
    assert np.array_equal(_sha(f32[0]), GOLD[tag + "_sha_f32"])            # torchvision's floats, bit for bit
    if h * w <= 512 * 512:                                                 # second image against the oracle directly
        assert np.array_equal(u8[1], po.resize_bicubic_u8(imgs[1], s))
        assert np.array_equal(f32[1].view(np.uint32), po.transform_x(imgs[1], s).view(np.uint32))


@pytest.mark.gpu
def test_gpu_transform_edge_cases():
    from aaclip_b200 import ops
    # 1-pixel-wide / tall inputs, identity on both axes, prime sizes, extreme shrink
    for (h, w, s, seed) in [(1, 1, 14, 1), (1, 37, 28, 2), (53, 1, 28, 3), (56, 56, 56, 4), (211, 307, 14, 5),
                            (5, 3, 97, 6)]:
        img = po.synth_image(h, w, seed)
        got = ops.resize_bicubic_u8(torch.from_numpy(img[None]).cuda(), s).cpu().numpy()[0]
        assert np.array_equal(got, po.resize_bicubic_u8(img, s)), (h, w, s)
    # custom statistics + empty batch
    img = po.synth_image(40, 52, 6)
    got = ops.preprocess_u8(torch.from_numpy(img[None]).cuda(), 28, mean=(0.5, 0.5, 0.5), std=(0.25, 0.5, 1.0)).cpu().numpy()[0]
    r = po.resize_bicubic_u8(img, 28).transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    want = (r - np.float32(0.5)) / np.asarray((0.25, 0.5, 1.0), np.float32).reshape(3, 1, 1)
    assert np.array_equal(got.view(np.uint32), want.astype(np.float32).view(np.uint32))
    assert ops.preprocess_u8(torch.empty(0, 8, 8, 3, dtype=torch.uint8, device="cuda"), 14).shape == (0, 3, 14, 14)
    with pytest.raises(Exception):
        ops.preprocess_u8(torch.zeros(1, 8, 8, 4, dtype=torch.uint8, device="cuda"), 14)


@pytest.mark.gpu
def test_gpu_full_size_batch_properties():
    """BASELINE-size input (64 x 1024 x 1024 raw images): size-independent properties - every image of the batch equals
    the same image transformed alone (no cross-image coupling), and a constant image stays constant."""
    from aaclip_b200 import ops
    base = torch.from_numpy(po.synth_image(1024, 1024, 1)).cuda()
    batch = torch.stack([torch.roll(base, shifts=17 * i, dims=1) for i in range(64)])
    batch[63] = 77
    out = ops.resize_bicubic_u8(batch, 336)
    assert np.array_equal(_sha(out[0].cpu().numpy()), GOLD["1024x1024_336_1_sha_u8"])
    for i in (1, 31, 62):
        assert torch.equal(out[i], ops.resize_bicubic_u8(batch[i:i + 1].contiguous(), 336)[0])
    assert bool((out[63] == 77).all())
