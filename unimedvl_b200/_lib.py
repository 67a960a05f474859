"""ctypes binding of libumv.so (include/umv.h).  Fails loudly when the library is missing: there
is no Python / CPU fallback for any compute call."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libumv.so")


class UmvError(RuntimeError):
    pass


class Dims(C.Structure):
    _fields_ = [
        ("hidden", C.c_int32), ("heads", C.c_int32), ("kv_heads", C.c_int32), ("inter", C.c_int32),
        ("layers", C.c_int32), ("vocab", C.c_int32), ("rope_theta", C.c_float), ("rms_eps", C.c_float),
        ("vit_hidden", C.c_int32), ("vit_heads", C.c_int32), ("vit_inter", C.c_int32), ("vit_layers", C.c_int32),
        ("vit_patch_dim", C.c_int32), ("vit_positions", C.c_int32), ("vit_eps", C.c_float),
        ("vit_pos_table", C.c_int32), ("latent_dim", C.c_int32), ("latent_pos_table", C.c_int32),
        ("max_tokens", C.c_int32), ("max_seqs", C.c_int32), ("kv_pages", C.c_int32),
        ("enable_vit", C.c_int32), ("enable_gen", C.c_int32), ("enable_vae", C.c_int32),
    ]


class DecodeRiders(C.Structure):
    _fields_ = [("n", C.c_int32), ("seqs", C.POINTER(C.c_int32)), ("tokens", C.POINTER(C.c_int64)), ("positions", C.POINTER(C.c_int32)),
                ("temperature", C.c_float), ("seed", C.c_uint64), ("next_tokens", C.c_void_p)]


class FlowArgs(C.Structure):
    _fields_ = [
        ("n_seqs", C.c_int32),
        ("seqs", C.POINTER(C.c_int32)), ("cfg_text_seqs", C.POINTER(C.c_int32)), ("cfg_img_seqs", C.POINTER(C.c_int32)),
        ("lat_lens", C.POINTER(C.c_int32)), ("positions", C.POINTER(C.c_int32)),
        ("cfg_text_positions", C.POINTER(C.c_int32)), ("cfg_img_positions", C.POINTER(C.c_int32)),
        ("marker_ids", C.POINTER(C.c_int64)), ("lat_pos_ids", C.c_void_p),
        ("timestep", C.c_float), ("cfg_text_scale", C.c_float), ("cfg_img_scale", C.c_float),
        ("cfg_renorm_min", C.c_float), ("renorm_type", C.c_int32),
    ]


UMV_OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM, ERR_STATE = 0, -1, -2, -3, -4, -5
BF16, F32, I64, I32, U8 = 0, 1, 2, 3, 4

# name -> (restype, argtypes); mirrors include/umv.h one to one
_P, _I, _F, _U64, _SZ = C.c_void_p, C.c_int32, C.c_float, C.c_uint64, C.c_size_t
_IP, _LP = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
SIGNATURES = {
    "umv_last_error": (C.c_char_p, []),
    "umv_abi_version": (C.c_int, []),
    "umv_create": (C.c_int, [C.POINTER(Dims), C.POINTER(_P)]),
    "umv_destroy": (C.c_int, [_P]),
    "umv_load_tensor": (C.c_int, [_P, C.c_char_p, _P, C.c_int, C.c_int, _LP, C.c_int]),
    "umv_export_tensor": (C.c_int, [_P, C.c_char_p, _P, _SZ]),
    "umv_fill_synthetic": (C.c_int, [_P, _U64]),
    "umv_finalize": (C.c_int, [_P]),
    "umv_weight_bytes": (C.c_int, [_P, _LP]),
    "umv_seq_new": (C.c_int, [_P, _IP]),
    "umv_seq_fork": (C.c_int, [_P, _I, _IP]),
    "umv_seq_free": (C.c_int, [_P, _I]),
    "umv_seq_len": (C.c_int, [_P, _I, _IP]),
    "umv_seq_truncate": (C.c_int, [_P, _I, _I]),
    "umv_seq_export": (C.c_int, [_P, _I, _I, _P, _P, _P]),
    "umv_pages_free": (C.c_int, [_P, _IP]),
    "umv_vit_embed": (C.c_int, [_P, _P, _P, _IP, _I, _P, _P]),
    "umv_patchify_u8": (C.c_int, [_P, _LP, _IP, _I, _I, _I, _P, _P, _P]),
    "umv_resize_bicubic_u8": (C.c_int, [_P, _I, _I, _P, _I, _I, _P, C.c_int64, _P]),
    "umv_resize_workspace_bytes": (C.c_int64, [_I, _I, _I, _I]),
    "umv_resize_coefficients": (C.c_int, [_I, _I, _IP, _IP, _IP]),
    "umv_embed_tokens": (C.c_int, [_P, _P, _I, _P, _P]),
    "umv_llm_forward": (C.c_int, [_P, _P, _I, _IP, _IP, _IP, C.POINTER(C.c_uint8), _I, _I, _P, _P]),
    "umv_lm_head": (C.c_int, [_P, _P, _I, _P, _P]),
    "umv_generate_text": (C.c_int, [_P, _I, _IP, _LP, _IP, _I, _F, _U64, _P, _P, _P, _P, _P]),
    "umv_flow_velocity": (C.c_int, [_P, C.POINTER(FlowArgs), _P, _P, _P]),
    "umv_flow_branches_last": (C.c_int, [_P, _IP]),
    "umv_flow_euler": (C.c_int, [_P, _P, _P, C.c_int64, _F, _I, _P]),
    "umv_latent_embed": (C.c_int, [_P, _P, _P, _I, _F, _P, _P]),
    "umv_vae_decode": (C.c_int, [_P, _P, _I, _I, _I, _P, _P]),
    "umv_vae_encode_moments": (C.c_int, [_P, _P, _I, _I, _I, _P, _P]),
    "umv_forward_cache_update_text": (C.c_int, [_P, _I, _IP, _IP, _LP, _IP, _P]),
    "umv_forward_cache_update_vit": (C.c_int, [_P, _I, _IP, _IP, _I, _LP, _IP, _P, _P, _I, _IP, _IP, _IP, _P]),
    "umv_forward_cache_update_text_riders": (C.c_int, [_P, _I, _IP, _IP, _LP, _IP, C.POINTER(DecodeRiders), _P]),
    "umv_forward_cache_update_vit_riders": (C.c_int, [_P, _I, _IP, _IP, _I, _LP, _IP, _P, _P, _I, _IP, _IP, _IP, C.POINTER(DecodeRiders), _P]),
    "umv_forward_cache_update_vit_prompt": (C.c_int, [_P, _I, _IP, _IP, _IP, _I, _LP, _IP, _P, _P, _I, _IP, _IP, _IP, C.POINTER(DecodeRiders), _P]),
    "umv_forward_cache_update_vae": (C.c_int, [_P, _I, _IP, _IP, _I, _LP, _IP, _P, _I, _I, _I, _IP, _I, _P, _IP, _F, _IP, _P]),
    "umv_vit_model": (C.c_int, [_P, _P, _P, _IP, _I, _P, _P]),
    "umv_connector": (C.c_int, [_P, _P, _I, _P, _P]),
    "umv_pos_embed": (C.c_int, [_P, _I, _P, _I, _P, _P]),
    "umv_vae2llm": (C.c_int, [_P, _P, _I, _P, _P]),
    "umv_llm2vae": (C.c_int, [_P, _P, _I, _P, _P]),
    "umv_time_embedder": (C.c_int, [_P, C.POINTER(C.c_float), _I, _P, _P]),
    "umv_vae_sample": (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "umv_decode_image_u8": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "umv_op_linear": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "umv_op_rmsnorm": (C.c_int, [_P, _P, _P, _I, _I, _F, _P]),
    "umv_op_layernorm": (C.c_int, [_P, _P, _P, _P, _I, _I, _F, _P]),
    "umv_op_attention": (C.c_int, [_P, _P, _P, _P, _I, _IP, _IP, _I, _I, _I, _I, _P]),
    "umv_op_attention_block": (C.c_int, [_P, _I, _P, _P, _I, _P, _I, _IP, _IP, _IP, C.POINTER(C.c_uint8), _I, _I, _P, _IP, _P]),
    "umv_op_vae_block": (C.c_int, [_P, C.c_char_p, _P, _I, _I, _I, _P, _IP, _P]),
    "umv_op_argmax": (C.c_int, [_P, _I, _I, _P, _P]),
    "umv_op_sample": (C.c_int, [_P, _I, _I, _F, _U64, _F, _P, _P]),
    "umv_bench_decode_linear": (C.c_int, [_P, _I, _I, _I, _LP, _P]),
    "umv_launch_count": (C.c_int64, []),
    "umv_trace_begin": (C.c_int, [_I]),
    "umv_trace_read": (C.c_int, [_P, _P, _I, _I, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load libumv.so (built by `python -m unimedvl_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UmvError(f"{LIB_PATH} is missing: build it with `python -m unimedvl_b200.build` "
                           "(there is no Python fallback for the CUDA engine)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError here == header / library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


_EXC = {ERR_INVALID: ValueError, ERR_UNSUPPORTED: NotImplementedError, ERR_CUDA: UmvError, ERR_NOMEM: MemoryError,
        ERR_STATE: AssertionError}


def check(rc: int) -> None:
    """Map a umv_status to the exception type the reference raises at the same boundary
    (ValueError inferencer.py:315,610; NotImplementedError bagel.py:1203; AssertionError bagel.py:1098)."""
    if rc != UMV_OK:
        msg = load().umv_last_error().decode(errors="replace")
        raise _EXC.get(rc, UmvError)(f"libumv: {msg} (status {rc})")


def i32_array(values):
    return (C.c_int32 * len(values))(*[int(v) for v in values])


def i64_array(values):
    return (C.c_int64 * len(values))(*[int(v) for v in values])
