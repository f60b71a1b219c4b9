"""ctypes binding of the C ABI in include/volta_b200.h (libvolta_b200.so, built in-tree by
`volta_b200.csrc.build`).  This is the only place the Python host side touches native code.

There is no fallback: if the library is missing the import of any compute entry point raises, and
every non-zero return code becomes a RuntimeError carrying `vb_last_error()`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvolta_b200.so")

F32, BF16 = 0, 1
GEMM_AUTO, GEMM_TC, GEMM_SIMT = 0, 1, 2
REGION_XENT, REGION_MSE, REGION_HUBER = 0, 1, 2
ATTN_AUTO, ATTN_MMA, ATTN_SIMT, ATTN_TC = 0, 1, 2, 3
(EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RELU, EPI_BIAS_DROP_RES, EPI_MUL, EPI_ADD, EPI_DRELU, EPI_ACCUM,
 EPI_BIAS_ADD) = range(10)
(ELT_COPY, ELT_ADD, ELT_MUL, ELT_RELU, ELT_DRELU, ELT_GELU, ELT_DGELU, ELT_SCALE, ELT_DROPOUT, ELT_AXPY, ELT_SWISH,
 ELT_DSWISH) = range(12)

_vp, _i32, _i64, _u64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float


class GemmArgs(C.Structure):
    _fields_ = [("A", _vp), ("B", _vp), ("D", _vp), ("M", _i64), ("N", _i64), ("K", _i64),
                ("lda", _i64), ("ldb", _i64), ("ldd", _i64), ("ab_dtype", _i32), ("d_dtype", _i32),
                ("a_kmajor", _i32), ("b_kmajor", _i32), ("epilogue", _i32), ("backend", _i32),
                ("bias", _vp), ("aux_in", _vp), ("aux_out", _vp), ("ld_aux", _i64), ("alpha", _f),
                ("dropout_p", _f), ("seed", _u64), ("offset", _u64), ("split_k", _i32), ("reserved", _i32),
                ("B2", _vp), ("bias2", _vp), ("split_M", _i64), ("m_limit", _vp), ("k_limit", _vp), ("m_limit2", _vp), ("limit_seg_M", _i64),
                ("D2", _vp), ("k_limit2", _vp), ("k_seg", _i64), ("m_hint", _i64)]


class AttnArgs(C.Structure):
    _fields_ = [("qkv_t", _vp), ("qkv_v", _vp), ("ld_t", _i64), ("ld_v", _i64), ("mask_t", _vp), ("mask_v", _vp),
                ("ctx_t", _vp), ("ctx_v", _vp), ("ld_ctx_t", _i64), ("ld_ctx_v", _i64), ("lse_t", _vp), ("lse_v", _vp),
                ("dctx_t", _vp), ("dctx_v", _vp), ("dqkv_t", _vp), ("dqkv_v", _vp),
                ("B", _i32), ("Lt", _i32), ("Lv", _i32), ("heads", _i32), ("d", _i32), ("pair_mask", _i32),
                ("dtype", _i32), ("scale", _f), ("dropout_p_t", _f), ("dropout_p_v", _f), ("seed", _u64),
                ("offset", _u64), ("backend", _i32), ("reserved", _i32), ("len_t", _vp), ("len_v", _vp), ("cu_t", _vp), ("cu_v", _vp)]


class RowSeqs(C.Structure):
    _fields_ = [("len_a", _vp), ("L_a", _i32), ("rows_a", _i64), ("len_b", _vp), ("L_b", _i32)]


class CastItem(C.Structure):
    _fields_ = [("src", _vp), ("dst", _vp), ("n", _i64)]


# name -> (restype, argtypes); every symbol declared in include/volta_b200.h
SIGNATURES = {
    "vb_version": (_i32, []),
    "vb_last_error": (C.c_char_p, []),
    "vb_device_check": (_i32, []),
    "vb_sm_count": (_i32, []),
    "vb_set_seed_pointer": (_i32, [_vp]),
    "vb_set_gemm_sm_limit": (_i32, [_i32]),
    "vb_gemm": (_i32, [C.POINTER(GemmArgs), _vp]),
    "vb_gemm_backend": (_i32, [C.POINTER(GemmArgs)]),
    "vb_gemm_bench": (_i32, [C.POINTER(GemmArgs), _vp, _i32, C.POINTER(C.c_float)]),
    "vb_layernorm_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f, _f, _i32, _vp, _vp, _vp, _i64, _i32, _vp]),
    "vb_layernorm_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f, _vp, _f, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "vb_attention_fwd": (_i32, [C.POINTER(AttnArgs), _vp]),
    "vb_attention_bwd": (_i32, [C.POINTER(AttnArgs), _vp]),
    "vb_attention_probs": (_i32, [C.POINTER(AttnArgs), _vp, _vp, _vp]),
    "vb_embed_gather_fwd": (_i32, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64), _i32, _vp, _i64, _i64, _i32, _i32, _i32, _vp]),
    "vb_embed_gather_bwd": (_i32, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i64), _i32, _vp, _i64, _i64, _i32, _i32, _vp]),
    "vb_colsum_by_key": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _i32, _vp, _vp]),
    "vb_box_project_fwd": (_i32, [_vp, _i32, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp]),
    "vb_box_project_bwd": (_i32, [_vp, _i32, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _vp]),
    "vb_xent_fwd": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _vp]),
    "vb_xent_bwd": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "vb_kl_fwd": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _i64, _i32, _f, _vp, _vp, _vp, _i32, _vp]),
    "vb_region_loss_fwd": (_i32, [_i32, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _f, _vp, _vp, _vp, _i32, _i32, _vp]),
    "vb_region_loss_bwd": (_i32, [_i32, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _f, _vp, _vp, _vp, _vp, _i32, _vp]),
    "vb_kl_bwd": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _i64, _i32, _f, _vp, _vp, _vp, _vp, _i32, _vp]),
    "vb_loss_finalize": (_i32, [_vp, _vp, _vp, _i32, _vp]),
    "vb_bce_logits_fwd": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _f, _vp, _i32, _vp]),
    "vb_bce_logits_bwd": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _f, _vp, _i64, _i32, _vp]),
    "vb_argmax_pick": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _vp, _i32, _vp]),
    "vb_eltwise": (_i32, [_i32, _vp, _i32, _i64, _vp, _i32, _i64, _vp, _i32, _i64, _i64, _i64, _f, _f, _u64, _u64, _vp]),
    "vb_colsum": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp]),
    "vb_zero_rows": (_i32, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _vp]),
    "vb_mask_lengths": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "vb_compact_plan": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "vb_compact_gather": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp]),
    "vb_compact_scatter": (_i32, [_vp, _i64, _vp, _vp, _i64, _i64, _i32, _i32, _vp]),
    "vb_gather_rows": (_i32, [_vp, _i64, _vp, _vp, _i64, _i64, _i64, _i32, _vp]),
    "vb_scatter_rows_add": (_i32, [_vp, _i64, _vp, _vp, _i64, _i64, _i64, _i32, _vp]),
    "vb_multi_cast_bf16": (_i32, [_vp, _i32, _i64, _vp]),
    "vb_box_sincos": (_i32, [_vp, _i32, _i32, _vp, _i64, _i64, _i32, _vp]),
    "vb_fill": (_i32, [_vp, _i64, _i32, _vp]),
    "vb_sumsq": (_i32, [_vp, _i64, _vp, _vp]),
    "vb_allreduce_f32": (_i32, [_vp, _u64, _vp, _i32, _i32, _i64, _i64, _f, _i32, _vp]),
    "vb_allreduce_pad_bytes": (_i32, []),
    "vb_adamw_flat": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, C.c_double, C.c_double, _f, _vp, _f, _vp]),
    "vb_radam_flat": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, C.c_double, C.c_double, _f, _vp, _f, _vp]),
    "vb_vlbert_prep": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp]),
    "vb_vlbert_index": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "vb_batched_colsum": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _vp]),
    "vb_retrieval_ranks": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _i32, _vp]),
    "vb_region_batch": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Loads libvolta_b200.so (raises with build instructions if it is absent)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "volta_b200: %s is missing - build it with `python -m volta_b200.csrc.build` "
                "(needs nvcc; there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def last_error():
    return load().vb_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise RuntimeError("volta_b200.%s failed (code %d): %s" % (what, rc, last_error()))


def call(name, *args):
    check(getattr(load(), name)(*args), name)
