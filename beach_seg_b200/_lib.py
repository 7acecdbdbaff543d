"""ctypes binding of libbseg.so (C ABI in include/bseg.h).  There is no CPU fallback: if the library is missing
or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libbseg.so"
CSRC_DIR = PKG_DIR / "csrc"

_lib = None


class BsegError(RuntimeError):
    pass


class LayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_w", "ln1_b", "qkv_w", "qkv_b", "rel_pos_h", "rel_pos_w", "proj_w", "proj_b", "ln2_w", "ln2_b",
        "lin1_w", "lin1_b", "lin2_w", "lin2_b")]


class Weights(C.Structure):
    _fields_ = [
        ("image_size", C.c_int), ("num_layers", C.c_int), ("merge_index", C.c_int), ("intermediate_indices", C.c_int * 4),
        ("layer_norm_eps", C.c_float),
        ("patch_w", C.c_void_p), ("patch_b", C.c_void_p), ("mask_token", C.c_void_p),
        ("segment_token_input", C.c_void_p), ("segment_token_prompt", C.c_void_p),
        ("type_token_semantic", C.c_void_p), ("type_token_instance", C.c_void_p),
        ("position_embeddings", C.c_void_p),
        ("layers", C.POINTER(LayerWeights)),
        ("enc_ln_w", C.c_void_p), ("enc_ln_b", C.c_void_p),
        ("dec_embed_w", C.c_void_p), ("dec_embed_b", C.c_void_p),
        ("dec_conv_w", C.c_void_p), ("dec_conv_b", C.c_void_p),
        ("dec_ln_w", C.c_void_p), ("dec_ln_b", C.c_void_p),
        ("dec_head_w", C.c_void_p), ("dec_head_b", C.c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/bseg.h declares
_vp, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
_f3 = C.POINTER(C.c_float)
SIGNATURES = {
    "bseg_last_error": (C.c_char_p, []),
    "bseg_version": (_i, []),
    "bseg_launch_count": (_ll, []),
    "bseg_create": (_i, [C.POINTER(Weights), C.POINTER(_vp), _vp]),
    "bseg_destroy": (_i, [_vp]),
    "bseg_workspace_bytes": (_sz, [_vp, _i]),
    "bseg_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "bseg_forward_query_half": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "bseg_set_graph_batch_limit": (_i, [_vp, _i]),
    "bseg_enable_fp32": (_i, [_vp, C.POINTER(Weights), _vp]),
    "bseg_workspace_bytes_f32": (_sz, [_vp, _i]),
    "bseg_forward_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "bseg_train_prepare": (_i, [_vp, _vp]),
    "bseg_train_workspace_bytes": (_sz, [_vp, _i]),
    "bseg_forward_train": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _sz, _vp, _vp]),
    "bseg_backward_to_prompt": (_i, [_vp, _vp, _i, _vp, _sz, _vp, _vp]),
    "bseg_train_workspace_bytes_f32": (_sz, [_vp, _i]),
    "bseg_forward_train_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _sz, _vp, _vp]),
    "bseg_backward_to_prompt_f32": (_i, [_vp, _vp, _i, _vp, _sz, _vp, _vp]),
    "bseg_attention_fwd_lse": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "bseg_attention_bwd_scratch_bytes": (_sz, [_i]),
    "bseg_attention_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "bseg_layernorm1024_bwd": (_i, [_vp, _vp, _ll, _vp, _vp, _vp, _vp, _ll, _f, _vp]),
    "bseg_gemm_bf16_dgelu": (_i, [_vp, _ll, _vp, _ll, _i, _i, _vp, _vp, _ll, _vp]),
    "bseg_pack_conv_w9_dgrad": (_i, [_vp, _vp, _vp]),
    "bseg_decoder_head_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "bseg_scene_stats": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "bseg_ingest_u16x4": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _f3, _f3, _vp, _vp, _ll, _vp, _vp,
                               _vp]),
    "bseg_ingest_native_u16x4": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _f3, _f3, _vp, _vp, _vp, _vp]),
    "bseg_ingest_native_f32x4": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _f3, _f3, _vp, _vp, _vp, _vp]),
    "bseg_scene_stats_f32": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "bseg_scene_stats_rows": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "bseg_scene_stats_finalize": (_i, [_vp, _vp, _vp]),
    "bseg_ingest_f32x4": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _f3, _f3, _vp, _vp, _ll, _vp, _vp,
                               _vp]),
    "bseg_merge_mosaic": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "bseg_preprocess_u8": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _f3, _f3, _vp, _vp]),
    "bseg_colorize_resize_norm255": (_i, [_vp, _vp, _i, _f3, _f3, _vp, _vp, _i, _i, _i, _vp]),
    "bseg_postprocess_semantic": (_i, [_vp, _vp, _i, _f3, _f3, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "bseg_colorize_norm": (_i, [_vp, _vp, _i, _f3, _f3, _vp, _i, _i, _i, _vp]),
    "bseg_decode_palette": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "bseg_mean_over_prompts": (_i, [_vp, _vp, _i, _i, _ll, _vp]),
    "bseg_train_aug_fwd": (_i, [_vp, _vp, _vp, C.POINTER(C.c_int32), _vp, _f, _f, _f3, _f3, _vp, _vp, _vp, _i, _i, _i,
                                _vp]),
    "bseg_train_aug_bwd": (_i, [_vp, _vp, C.POINTER(C.c_int32), _f3, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "bseg_vote_accumulate": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _i, _vp]),
    "bseg_vote_argmax": (_i, [_vp, _vp, _ll, _vp]),
    "bseg_paste_tiles_u8": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp]),
    "bseg_overlay_prediction": (_i, [_vp, _vp, _vp, _i, _ll, _vp, _vp]),
    "bseg_loss_smoothl1_fwd_bwd": (_i, [_vp, _vp, _vp, _f, _i, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "bseg_gemm_set_cta_pairs": (_i, [_i]),
    "bseg_gemm_set_small_tiles": (_i, [_i]),
    "bseg_gemm_set_fused_ln": (_i, [_i]),
    "bseg_set_pdl": (_i, [_i]),
    "bseg_gemm_resid_ln_scratch_bytes": (C.c_size_t, [_ll]),
    "bseg_gemm_bf16_resid_ln": (_i, [_vp, _ll, _vp, _ll, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp]),
    "bseg_gemm_bf16": (_i, [_vp, _ll, _vp, _ll, _i, _i, _vp, _vp, _ll, _i, _i, _vp]),
    "bseg_layernorm1024": (_i, [_vp, _ll, _vp, _vp, _vp, _ll, _ll, _f, _vp]),
    "bseg_attention": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "bseg_attention_grid": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "bseg_relcat_rows": (_i, [_i, _i]),
    "bseg_pack_relcat_grid": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "bseg_pack_relcat": (_i, [_vp, _vp, _vp, _vp]),
    "bseg_decoder_head": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "bseg_pack_conv_w9": (_i, [_vp, _vp, _vp]),
    "bseg_f32_to_bf16": (_i, [_vp, _vp, _ll, _vp]),
    "bseg_profile_enable": (_i, [_i]),
    "bseg_profile_collect_gemm": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "bseg_profile_collect": (_i, [C.POINTER(C.c_double), C.POINTER(_ll), C.POINTER(C.c_double),
                                  C.POINTER(C.c_double)]),
}


def build(verbose: bool = False) -> Path:
    """Compile libbseg.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(CSRC_DIR), "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise BsegError("building libbseg.so failed")
    return LIB_PATH


def lib():
    """The loaded library (loads on first use; raises if it was never built)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise BsegError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().bseg_last_error().decode("utf-8", "replace")
        raise BsegError(f"{what or 'libbseg call'} failed (rc={rc}): {msg}")


def f3(vals):
    return (C.c_float * 3)(*[float(v) for v in vals])


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
