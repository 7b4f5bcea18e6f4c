"""ctypes binding of libovmr_b200.so (the C-ABI declared in include/ovmr_b200.h).

There is no fallback: if the shared library is missing or a CUDA device is absent, every compute
entry point raises.  Build the library with `python __graft_entry__.py` (or `make -C ovmr_b200/csrc`).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libovmr_b200.so")

c_void_p, c_int, c_ll, c_float, c_size_t = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t


class BlockWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "ln1_w", "ln1_b", "qkv_w", "qkv_b", "out_w", "out_b", "ln2_w", "ln2_b", "fc_w", "fc_b", "proj_w", "proj_b",
        "qkv_wf", "qkv_cs", "qkv_bf", "fc_wf", "fc_cs", "fc_bf")]


class Transformer(C.Structure):
    _fields_ = [("width", c_int), ("heads", c_int), ("layers", c_int), ("fp16", c_int),
                ("blocks", C.POINTER(BlockWeights))]


class Vit(C.Structure):
    _fields_ = [("resolution", c_int), ("patch", c_int), ("width", c_int), ("embed_dim", c_int), ("k_pad", c_int),
                ("conv_w", c_void_p), ("class_embedding", c_void_p), ("positional_embedding", c_void_p),
                ("ln_pre_w", c_void_p), ("ln_pre_b", c_void_p), ("ln_post_w", c_void_p), ("ln_post_b", c_void_p),
                ("proj_t", c_void_p), ("transformer", Transformer)]


class Text(C.Structure):
    _fields_ = [("width", c_int), ("embed_dim", c_int), ("context_length", c_int),
                ("positional_embedding", c_void_p), ("ln_final_w", c_void_p), ("ln_final_b", c_void_p),
                ("text_projection_t", c_void_p), ("transformer", Transformer)]


# name -> (restype, argtypes); must list every symbol declared in include/ovmr_b200.h
SIGNATURES = {
    "ovmr_abi_version": (c_int, []),
    "ovmr_last_error": (C.c_char_p, []),
    "ovmr_launch_count": (c_ll, []),
    "ovmr_profile_enable": (c_int, [c_int]),
    "ovmr_profile_summary": (c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(c_ll), c_int]),
    "ovmr_transformer_workspace_bytes": (c_size_t, [c_ll, c_int]),
    "ovmr_vit_workspace_bytes": (c_size_t, [C.POINTER(Vit), c_int]),
    "ovmr_text_workspace_bytes": (c_size_t, [C.POINTER(Text), c_int, c_int]),
    "ovmr_transformer_forward": (c_int, [C.POINTER(Transformer), c_void_p, c_int, c_int, c_int, c_void_p, c_size_t,
                                         c_void_p]),
    "ovmr_vit_forward": (c_int, [C.POINTER(Vit), c_void_p, c_int, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "ovmr_vit_forward_u8": (c_int, [C.POINTER(Vit), c_void_p, C.POINTER(c_float), c_int, c_void_p, c_int, c_void_p,
                                    c_size_t, c_void_p]),
    "ovmr_text_forward": (c_int, [C.POINTER(Text), c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p,
                                  c_size_t, c_void_p]),
    "ovmr_gemm_tn": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_ll,
                             c_void_p, c_ll, c_int, c_int, c_float, c_int, c_int, c_int, c_void_p]),
    "ovmr_gemm_tn_resid_ln": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_ll, c_void_p,
                                      c_ll, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "ovmr_gemm_ln_scratch_bytes": (c_size_t, [c_ll, c_int]),
    "ovmr_gemm_tn_resid_ln_gx": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_ll, c_void_p,
                                         c_ll, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p, c_size_t, C.c_uint, c_void_p]),
    "ovmr_layernorm": (c_int, [c_void_p, c_ll, c_int, c_int, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_ll,
                               c_void_p, c_ll, c_void_p, c_void_p, c_int, c_void_p]),
    "ovmr_attention": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ovmr_attention_impl": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ovmr_u8_normalization_is_exact": (c_int, [C.POINTER(c_float)]),
    "ovmr_patch_embed": (c_int, [c_void_p, c_int, C.POINTER(c_float), c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                 c_int, c_int, c_void_p]),
    "ovmr_patchify": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ovmr_patchify_u8": (c_int, [c_void_p, C.POINTER(c_float), c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ovmr_resample_coeffs": (c_int, [c_int, c_int, c_int, C.POINTER(c_int), C.POINTER(c_int), c_int]),
    "ovmr_resize_crop_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                    c_int, C.POINTER(c_int), c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p,
                                    c_void_p]),
    "ovmr_build_text_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                                     c_int, c_int, c_int, c_int, c_void_p]),
    "ovmr_agg_build": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ovmr_take_rows": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p]),
    "ovmr_l2norm": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_void_p, c_void_p]),
    "ovmr_segmented_mean": (c_int, [c_void_p, c_ll, c_int, c_int, c_void_p, c_int, c_void_p]),
    "ovmr_split_bf16": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_int, c_ll, c_void_p]),
    "ovmr_fusion_softmax_topk": (c_int, [c_void_p, c_ll, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_ll, c_int,
                                         c_void_p, c_void_p, c_void_p]),
    "ovmr_head_fused": (c_int, [c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_ll, c_int, c_void_p,
                                c_void_p, c_void_p]),
    "ovmr_head_fused_argmax": (c_int, [c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ovmr_argmax_segments": (c_int, [c_void_p, c_ll, c_ll, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ovmr_f1_counts": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p]),
    "ovmr_fusion_weights": (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "ovmr_layernorm_backward": (c_int, [c_void_p, c_int, c_int, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p]),
    "ovmr_quickgelu_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "ovmr_cast_16": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "ovmr_transpose_16": (c_int, [c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_ll, c_int, c_void_p]),
    "ovmr_colsum": (c_int, [c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_int, c_void_p]),
    "ovmr_l2norm_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "ovmr_cross_entropy": (c_int, [c_void_p, c_ll, c_void_p, c_int, c_int, c_void_p, c_void_p, c_ll, c_void_p]),
    "ovmr_attention_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                        C.c_uint, c_void_p]),
    "ovmr_attention_dropout_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, C.c_uint,
                                               c_void_p]),
    "ovmr_dropout_16": (c_int, [c_void_p, c_void_p, c_ll, c_float, C.c_uint, c_int, c_void_p]),
    "ovmr_dropout_add": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_float, C.c_uint, c_void_p]),
    "ovmr_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float, c_float, c_float,
                               c_int, c_void_p]),
}

_lib = None


class OvmrNativeError(RuntimeError):
    pass


def load():
    """dlopen the C-ABI library (no GPU needed just to load it and resolve symbols)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise OvmrNativeError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "ovmr_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.ovmr_abi_version() != 2:
        raise OvmrNativeError("libovmr_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def lib():
    l = load()
    if not torch.cuda.is_available():
        raise OvmrNativeError("ovmr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return l


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().ovmr_last_error().decode("utf-8", "replace")
        raise OvmrNativeError(f"{what or 'ovmr call'} failed (status {rc}): {msg}")


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().ovmr_launch_count())


PROFILE_CLASSES = ("gemm", "attention", "layernorm", "patchify", "head", "gemm_ln")


def profile_enable(on: bool):
    check(load().ovmr_profile_enable(int(on)), "ovmr_profile_enable")


def profile_summary() -> dict:
    """{class: {"ms": elapsed, "work": FLOPs or bytes, "launches": n}} for the launches recorded since
    profile_enable(True).  Synchronises."""
    n = len(PROFILE_CLASSES)
    ms, work, cnt = (C.c_double * n)(), (C.c_double * n)(), (c_ll * n)()
    check(load().ovmr_profile_summary(ms, work, cnt, n), "ovmr_profile_summary")
    return {name: {"ms": ms[i], "work": work[i], "launches": int(cnt[i])} for i, name in enumerate(PROFILE_CLASSES)}
