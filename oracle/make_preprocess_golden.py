"""Pin oracle/preprocess_oracle.py against the reference's own transform stack (PIL + torchvision, exactly the
Compose of dataset/__init__.py:127-136) and write tests/golden/preprocess_*.npz.

    python oracle/make_preprocess_golden.py

Inputs are regenerated from seeds (preprocess_oracle.synth_image); the fixture keeps the full uint8 resize of the
small cases, a strided sample of the larger ones and a sha256 of every full result.
"""
import hashlib
import os
import sys

import numpy as np
import torch
from PIL import Image
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import preprocess_oracle as po  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

# (H0, W0, S, seed): shrink by 3.05 (MVTec 1024 -> 336), non-square shrink, enlarge, mixed, identity on one axis
CASES = [(1024, 1024, 336, 1), (700, 900, 336, 2), (120, 90, 336, 3), (200, 448, 224, 4), (336, 500, 336, 5),
         (40, 52, 28, 6), (1024, 1024, 518, 7)]


def reference_transform(img: np.ndarray, size: int):
    t = transforms.Compose([
        transforms.Resize((size, size), Image.BICUBIC),
        transforms.ToTensor(),
        transforms.Normalize(mean=(0.48145466, 0.4578275, 0.40821073), std=(0.26862954, 0.26130258, 0.27577711)),
    ])
    pil = Image.fromarray(img)
    return np.asarray(pil.resize((size, size), Image.BICUBIC)), t(pil).numpy()


def main():
    out = {}
    for (h, w, s, seed) in CASES:
        img = po.synth_image(h, w, seed)
        ref_u8, ref_f = reference_transform(img, s)
        mine_u8 = po.resize_bicubic_u8(img, s)
        mine_f = po.transform_x(img, s)
        assert np.array_equal(ref_u8, mine_u8), f"resize mismatch {h}x{w}->{s}: {np.abs(ref_u8.astype(int) - mine_u8).max()}"
        assert np.array_equal(ref_f.view(np.uint32), mine_f.view(np.uint32)), f"normalise mismatch {h}x{w}->{s}"
        tag = f"{h}x{w}_{s}_{seed}"
        out[tag + "_sha_u8"] = np.frombuffer(hashlib.sha256(ref_u8.tobytes()).digest(), np.uint8)
        out[tag + "_sha_f32"] = np.frombuffer(hashlib.sha256(ref_f.tobytes()).digest(), np.uint8)
        out[tag + "_sub_f32"] = ref_f[:, ::13, ::11].copy()
        if s <= 56:
            out[tag + "_u8"] = ref_u8
        print(f"{tag}: oracle == PIL/torchvision bit for bit ({ref_u8.shape}, {ref_f.shape})")
    out["cases"] = np.asarray(CASES, np.int32)
    np.savez_compressed(os.path.join(GOLD, "preprocess_pil.npz"), **out)
    print("wrote", os.path.join(GOLD, "preprocess_pil.npz"))


if __name__ == "__main__":
    main()
