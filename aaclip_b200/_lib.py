"""ctypes binding of libaaclip_b200.so (the C ABI declared in include/aaclip_b200.h).

The library is the only compute path of the package: if it is missing the import fails loudly, there is
no PyTorch / CPU fallback.  All pointer arguments are raw device (or, where named host_*, host) addresses.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libaaclip_b200.so"

# mirrors of the #defines in include/aaclip_b200.h
ACT_NONE, ACT_GELU_ERF, ACT_QUICK_GELU, ACT_LEAKY = 0, 1, 2, 3
OUT_BF16, OUT_F32, OUT_F32_RESID, OUT_F32_PATCH = 0, 1, 2, 3
HEAD_TEST_INDUSTRIAL, HEAD_TEST_MEDICAL, HEAD_TRAIN_SOFTMAX = 0, 1, 2

W = dict(
    V_CONV1=0, V_CLS=1, V_POS=2, V_LN_PRE_G=3, V_LN_PRE_B=4, V_LN_POST_G=5, V_LN_POST_B=6,
    V_LN1_G=10, V_LN1_B=11, V_QKV_W=12, V_QKV_B=13, V_OUT_W=14, V_OUT_B=15, V_LN2_G=16, V_LN2_B=17,
    V_FC_W=18, V_FC_B=19, V_PROJ_W=20, V_PROJ_B=21,
    I_ADAPTER=30, I_SEG_PROJ=31, I_DET_PROJ=32,
    T_TOKEN_EMB=40, T_POS=41, T_LN_FINAL_G=42, T_LN_FINAL_B=43,
    T_LN1_G=50, T_LN1_B=51, T_QKV_W=52, T_QKV_B=53, T_OUT_W=54, T_OUT_B=55, T_LN2_G=56, T_LN2_B=57,
    T_FC_W=58, T_FC_B=59, T_PROJ_W=60, T_PROJ_B=61,
    T_ADAPTER=70, T_FINAL_PROJ=71,
)


PROFILE_CLASSES = ["gemm_qkv", "gemm_out", "gemm_fc", "gemm_proj", "gemm_adapter", "gemm_segdet", "gemm_patch",
                   "attention", "layernorm", "adapter_mix", "cast", "l2norm", "det_mean", "stem_misc", "head_maps",
                   "other"]


class AaclipCfg(C.Structure):
    _fields_ = [
        ("image_size", C.c_int), ("patch_size", C.c_int), ("width", C.c_int), ("heads", C.c_int),
        ("layers", C.c_int), ("mlp_width", C.c_int), ("embed_dim", C.c_int), ("act", C.c_int),
        ("image_adapt_until", C.c_int), ("image_adapt_weight", C.c_float),
        ("n_levels", C.c_int), ("levels", C.c_int * 8), ("proj_relu", C.c_int),
        ("t_context", C.c_int), ("t_vocab", C.c_int), ("t_width", C.c_int), ("t_heads", C.c_int),
        ("t_layers", C.c_int), ("text_adapt_until", C.c_int), ("text_adapt_weight", C.c_float),
        ("max_batch", C.c_int), ("max_text", C.c_int), ("cta_group", C.c_int), ("ln_fold", C.c_int),
    ]


# name -> (restype, argtypes); every symbol include/aaclip_b200.h declares
_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
SIGNATURES = {
    "aaclip_last_error": (C.c_char_p, []),
    "aaclip_abi_version": (_i, []),
    "aaclip_create": (_i, [C.POINTER(_vp), C.POINTER(AaclipCfg), _i]),
    "aaclip_destroy": (None, [_vp]),
    "aaclip_set_weight": (_i, [_vp, _i, _i, _vp, _ll, _i, _vp]),
    "aaclip_device_bytes": (_ll, [_vp]),
    "aaclip_launch_count": (_ll, [_vp]),
    "aaclip_profile_enable": (_i, [_vp, _i]),
    "aaclip_profile_read": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(_ll), _i]),
    "aaclip_profile_span_ms": (C.c_double, [_vp]),
    "aaclip_visual_forward": (_i, [_vp, _vp, _i, C.POINTER(_vp), _i, _vp, _vp]),
    "aaclip_anomaly_head_workspace_bytes": (_ll, [_i, _i, _i]),
    "aaclip_anomaly_head": (_i, [C.POINTER(_vp), _i, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _ll, _vp]),
    "aaclip_map_minmax": (_i, [_vp, _i, _ll, _vp, _vp]),
    "aaclip_forward_fused": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "aaclip_forward_fused_host": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "aaclip_submit_host": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, C.POINTER(_ll)]),
    "aaclip_wait_host": (_i, [_vp, _ll]),
    "aaclip_preprocess_scratch_bytes": (_ll, [_i, _i, _i, _i]),
    "aaclip_preprocess_u8": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "aaclip_resize_bicubic_u8": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "aaclip_submit_host_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, C.POINTER(_ll)]),
    "aaclip_text_forward": (_i, [_vp, _vp, _i, _vp, _vp]),
    "aaclip_text_anchor": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "aaclip_gemm_bf16": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _i, _i, _vp]),
    "aaclip_gemm_resid_ln": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _i, _vp]),
    "aaclip_gemm_lnfold": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _f, _vp, _i, _i, _i, _vp]),
    "aaclip_rowstats_cast": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp]),
    "aaclip_fold_ln_weight": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "aaclip_layernorm": (_i, [_vp, _vp, _vp, _f, _i, _i, _vp, _vp, _vp]),
    "aaclip_attention": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "aaclip_attention_trace": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "aaclip_adapter_mix": (_i, [_vp, _vp, _f, _i, _i, _vp]),
    "aaclip_set_text_final": (_i, [_vp, _i]),
    "aaclip_dapm_replace": (_i, [_vp, _i]),
    "aaclip_encode_image": (_i, [_vp, _vp, _i, C.POINTER(_vp), _vp, _i, _vp]),
    "aaclip_vv_attention": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "aaclip_add_image_vector": (_i, [_vp, _vp, _i, _i, _i, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and declare every entry point's signature."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m aaclip_b200.build` (needs nvcc). "
            "aaclip_b200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header ever diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class AaclipError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        msg = load().aaclip_last_error()
        raise AaclipError(f"aaclip_b200 error {rc}: {msg.decode() if msg else '?'}")


def ptr(t) -> int:
    """Raw data pointer of a torch tensor (or None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def cur_stream(device=None) -> int:
    """Raw handle of torch's current stream ON `device` (a tensor's or an engine's device, not the thread's current
    one: a stream handle is only valid on the device it was created on)."""
    import torch

    return torch.cuda.current_stream(device).cuda_stream
