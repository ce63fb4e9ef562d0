"""aaclip_b200 - B200-native (sm_100a) implementation of the AA-CLIP inference hot path.

Public surface mirrors the reference (wei-paul/AA-CLIP): `AdaptedCLIP` (model/adapter.py) and
`calculate_similarity_map` (forward_utils.py); everything computes in libaaclip_b200.so (hand-written CUDA,
C ABI in include/aaclip_b200.h).  Importing the package does not need a GPU; running it does.
"""
from .synth import ModelCfg, VIT_L_14_336  # noqa: F401

__all__ = ["AdaptedCLIP", "CLIP", "create_model", "calculate_similarity_map", "Engine", "ModelCfg", "VIT_L_14_336",
           "CLIPImageEncoder", "surgery_patch_features"]


def __getattr__(name):
    # lazy: keeps `import aaclip_b200.build` usable before the shared library exists
    if name == "AdaptedCLIP":
        from .adapter import AdaptedCLIP
        return AdaptedCLIP
    if name in ("CLIP", "create_model"):
        from . import clip
        return getattr(clip, name)
    if name == "calculate_similarity_map":
        from .forward_utils import calculate_similarity_map
        return calculate_similarity_map
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name in ("CLIPImageEncoder", "surgery_patch_features"):
        from . import surgery
        return getattr(surgery, name)
    raise AttributeError(name)
