"""ctypes binding of libvitb200.so (the C ABI declared in include/vitb200.h).

The library is the ONLY compute path: if it cannot be loaded, importing the package's ops raises —
there is no eager / CPU fallback anywhere in the product (the oracle lives under oracle/ and is
test infrastructure only).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvitb200.so")

F32, BF16 = 0, 1
EPI_NONE, EPI_GELU, EPI_GELU_BWD, EPI_GELU_DG, EPI_MUL_AUX = 0, 1, 2, 3, 4


class VitbError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise VitbError(
            "libvitb200.so is missing at %s — run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no fallback path)" % LIB_PATH
        )
    return C.CDLL(LIB_PATH)


_lib = _load()


class GemmParams(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_int32),
        ("M", C.c_int32),
        ("N", C.c_int32),
        ("num_segments", C.c_int32),
        ("A", C.c_void_p * 3),
        ("B", C.c_void_p * 3),
        ("lda", C.c_int64 * 3),
        ("ldb", C.c_int64 * 3),
        ("K", C.c_int32 * 3),
        ("a_mn_major", C.c_int32),
        ("b_mn_major", C.c_int32),
        ("split_k", C.c_int32),
        ("epilogue", C.c_int32),
        ("D", C.c_void_p),
        ("ldd", C.c_int64),
        ("d_dtype", C.c_int32),
        ("accumulate", C.c_int32),
        ("D2", C.c_void_p),
        ("ldd2", C.c_int64),
        ("bias", C.c_void_p),
        ("row_bias", C.c_void_p),
        ("row_bias_group", C.c_int32),
        ("row_remap_group", C.c_int32),
        ("residual", C.c_void_p),
        ("ldr", C.c_int64),
        ("r_dtype", C.c_int32),
        ("_pad0", C.c_int32),
        ("aux", C.c_void_p),
        ("ldaux", C.c_int64),
        ("colsum", C.c_void_p),
        ("n_groups", C.c_int32),
        ("_pad1", C.c_int32),
        ("b_group_stride", C.c_int64),
        ("d_group_stride", C.c_int64),
        ("m_dev", C.c_void_p),
    ]


def _sig(name, argtypes, restype=C.c_int):
    fn = getattr(_lib, name)
    fn.argtypes = argtypes
    fn.restype = restype
    return fn


vitb_version = _sig("vitb_version", [])
_vitb_last_error = _sig("vitb_last_error", [C.c_char_p, C.c_size_t])
vitb_device_check = _sig("vitb_device_check", [])
vitb_struct_size = _sig("vitb_struct_size", [C.c_int])
_vitb_gemm = _sig("vitb_gemm", [C.POINTER(GemmParams), C.c_void_p])


def last_error():
    buf = C.create_string_buffer(1024)
    _vitb_last_error(buf, 1024)
    return buf.value.decode(errors="replace")


LAUNCHES = [0]  # number of successful kernel-launching ABI calls (each launches exactly one kernel)


def check(status, what):
    if status != 0:
        raise VitbError("%s failed (%d): %s" % (what, status, last_error()))
    LAUNCHES[0] += 1


def check_host(status, what):
    """Status check of a host-only entry point (launches nothing)."""
    if status != 0:
        raise VitbError("%s failed (%d): %s" % (what, status, last_error()))


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise VitbError("unsupported dtype %s" % t.dtype)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise VitbError(
                "vit-of-pytorch_b200 ops run only on a B200 (sm_100a) CUDA device; got a %s tensor — "
                "there is no CPU path" % t.device
            )


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class AttnParams(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_int32),
        ("dtype", C.c_int32),
        ("B", C.c_int32),
        ("H", C.c_int32),
        ("Nq", C.c_int32),
        ("Nk", C.c_int32),
        ("head_dim", C.c_int32),
        ("_pad0", C.c_int32),
        ("q", C.c_void_p),
        ("k", C.c_void_p),
        ("v", C.c_void_p),
        ("o", C.c_void_p),
        ("lse", C.c_void_p),
        ("q_batch_stride", C.c_int64), ("q_row_stride", C.c_int64),
        ("k_batch_stride", C.c_int64), ("k_row_stride", C.c_int64),
        ("v_batch_stride", C.c_int64), ("v_row_stride", C.c_int64),
        ("o_batch_stride", C.c_int64), ("o_row_stride", C.c_int64),
        ("dout", C.c_void_p),
        ("do_batch_stride", C.c_int64), ("do_row_stride", C.c_int64),
        ("dq", C.c_void_p),
        ("dk", C.c_void_p),
        ("dv", C.c_void_p),
        ("dq_batch_stride", C.c_int64), ("dq_row_stride", C.c_int64),
        ("dk_batch_stride", C.c_int64), ("dk_row_stride", C.c_int64),
        ("dv_batch_stride", C.c_int64), ("dv_row_stride", C.c_int64),
    ]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_vitb_layernorm_fwd = _sig("vitb_layernorm_fwd", [_vp, _i, _i64, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp])
_vitb_layernorm_bwd = _sig("vitb_layernorm_bwd", [_vp, _i, _vp, _i64, _vp, _vp, _vp, _i, _i, _vp, _i64, _vp, _i64,
                                                   _vp, _vp, _vp, _vp, _vp, _vp])
_vitb_layernorm_bwd_sparse_res = _sig("vitb_layernorm_bwd_sparse_res", [_vp, _i, _vp, _i64, _vp, _vp, _vp, _i, _i, _vp, _i64, _i,
                                                                         _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp])
vitb_attn_supported_tc = _sig("vitb_attn_supported_tc", [_i, _i, _i])
vitb_attn_fwd_supported_tc = _sig("vitb_attn_fwd_supported_tc", [_i, _i, _i])
_vitb_attn_fwd_tc = _sig("vitb_attn_fwd_tc", [C.POINTER(AttnParams), _vp])
_vitb_attn_bwd_tc = _sig("vitb_attn_bwd_tc", [C.POINTER(AttnParams), _vp])
vitb_attn_ws_supported = _sig("vitb_attn_ws_supported", [_i, _i, _i, _i])
_vitb_attn_fwd_ws = _sig("vitb_attn_fwd_ws", [C.POINTER(AttnParams), _vp])         # persistent warp-specialised kernels
_vitb_attn_bwd_ws = _sig("vitb_attn_bwd_ws", [C.POINTER(AttnParams), _vp])
_vitb_attn_fwd_simt = _sig("vitb_attn_fwd_simt", [C.POINTER(AttnParams), _vp])
_vitb_attn_bwd_simt = _sig("vitb_attn_bwd_simt", [C.POINTER(AttnParams), _vp])
vitb_attn_bwd_long_supported = _sig("vitb_attn_bwd_long_supported", [_i, _i, _i])
_vitb_attn_bwd_tc_long = _sig("vitb_attn_bwd_tc_long", [C.POINTER(AttnParams), _vp, _vp])   # any token count, head_dim 64
vitb_attn_q1_supported = _sig("vitb_attn_q1_supported", [_i, _i])
_vitb_attn_q1_bwd = _sig("vitb_attn_q1_bwd", [C.POINTER(AttnParams), _vp])           # bf16 gradients, single query
_vitb_cast_split = _sig("vitb_cast_split", [_vp, _i64, _vp, _vp, _vp])
_vitb_im2col = _sig("vitb_im2col", [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp])
_vitb_cls_rows = _sig("vitb_cls_rows", [_vp, _i, _i, _i, _vp, _vp, _vp])
_vitb_embed_bwd = _sig("vitb_embed_bwd", [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp])
_vitb_gelu_bwd = _sig("vitb_gelu_bwd", [_vp, _vp, _vp, _i64, _i, _vp])
_vitb_dropout_fwd = _sig("vitb_dropout_fwd", [_vp, _i, _vp, _vp, _vp, _i64, _f, C.c_uint64, _vp, _vp])
_vitb_dropout_bwd = _sig("vitb_dropout_bwd", [_vp, _i, _vp, _vp, _i, _i64, _f, _vp])
_vitb_colsum = _sig("vitb_colsum", [_vp, _i, _i, _i, _i64, _vp, _vp])
_vitb_colsum3 = _sig("vitb_colsum3", [_vp, _i, _i, _i, _i64, _vp, _vp, _vp, _vp])
_vitb_cross_entropy = _sig("vitb_cross_entropy", [_vp, _vp, _i, _i, _vp, _vp, _vp])
_vitb_sgd_momentum = _sig("vitb_sgd_momentum", [_vp, _vp, _vp, _i64, _f, _vp, _f, _f, _f, _i, _i, _vp, _vp, _vp])
_vitb_adamw = _sig("vitb_adamw", [_vp, _vp, _vp, _vp, _i64, _f, _vp, _f, _f, _f, _f, _i, _vp, _vp, _vp, _vp, _vp])
_vitb_adamw_segments = _sig("vitb_adamw_segments", [_vp, _vp, _vp, _vp, _i64, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp])
vitb_p2p_pad_words = _sig("vitb_p2p_pad_words", [])
_vitb_p2p_allreduce = _sig("vitb_p2p_allreduce", [_vp, _vp, _vp, _i, _i, _i64, _f, _vp])
_vitb_sumsq = _sig("vitb_sumsq", [_vp, _i64, _vp, _vp])
_vitb_router_decide_fwd = _sig("vitb_router_decide_fwd", [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp])
_vitb_router_decide_bwd = _sig("vitb_router_decide_bwd", [_vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _i, _f, _vp, _vp])
_vitb_token_mean_fwd = _sig("vitb_token_mean_fwd", [_vp, _i, _i, _i, _i, _i, _vp, _vp])
_vitb_token_mean_bwd = _sig("vitb_token_mean_bwd", [_vp, _i, _i, _i, _i, _i, _vp, _vp])
_vitb_select_rows = _sig("vitb_select_rows", [_vp, _vp, _vp, C.c_uint32, _i, _i, _i, _vp, _vp])
_vitb_select_rows_flag = _sig("vitb_select_rows_flag", [_vp, _vp, _vp, C.c_uint32, _i, _i, _i, _vp, _vp, _vp])
_vitb_compact_rows = _sig("vitb_compact_rows", [_vp, C.c_uint32, _i, _vp, _vp, _vp])
_vitb_gather_rows = _sig("vitb_gather_rows", [_vp, _i64, _i, _vp, _vp, _i, _i, _vp, _i64, _vp])
_vitb_scatter_rows = _sig("vitb_scatter_rows", [_vp, _i64, _i, _vp, _vp, _i, _i, _vp, _i64, _vp])
_vitb_distill_loss = _sig("vitb_distill_loss", [_vp, _i64, _vp, _i64, _i, _i, _i, _vp, _vp, _vp])
_vitb_active_loss = _sig("vitb_active_loss", [_vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp])
_vitb_clip_coef = _sig("vitb_clip_coef", [_vp, _f, _vp, _vp, _vp])
_vitb_resize_tables_host = _sig("vitb_resize_tables_host", [_i, _i, _vp, _vp, _i, C.POINTER(C.c_int)])
_vitb_image_prep = _sig("vitb_image_prep", [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _i,
                                            _vp, _vp, _vp, _vp])

EXPORTED_SYMBOLS = [
    "vitb_version", "vitb_last_error", "vitb_device_check", "vitb_struct_size", "vitb_gemm", "vitb_layernorm_fwd",
    "vitb_layernorm_bwd", "vitb_attn_supported_tc", "vitb_attn_fwd_supported_tc", "vitb_attn_fwd_tc", "vitb_attn_bwd_tc",
    "vitb_attn_fwd_simt", "vitb_attn_bwd_simt", "vitb_cast_split", "vitb_im2col", "vitb_cls_rows",
    "vitb_embed_bwd", "vitb_colsum", "vitb_cross_entropy", "vitb_sgd_momentum", "vitb_adamw",
    "vitb_sumsq", "vitb_clip_coef", "vitb_gelu_bwd", "vitb_router_decide_fwd", "vitb_router_decide_bwd",
    "vitb_token_mean_fwd", "vitb_token_mean_bwd", "vitb_select_rows", "vitb_colsum3",
    "vitb_resize_tables_host", "vitb_image_prep",
    "vitb_distill_loss", "vitb_active_loss", "vitb_compact_rows", "vitb_gather_rows", "vitb_scatter_rows", "vitb_attn_ws_supported", "vitb_attn_fwd_ws", "vitb_attn_bwd_ws",
    "vitb_layernorm_bwd_sparse_res", "vitb_attn_q1_supported", "vitb_attn_q1_bwd", "vitb_dropout_fwd", "vitb_dropout_bwd",
    "vitb_adamw_segments", "vitb_select_rows_flag", "vitb_attn_bwd_long_supported", "vitb_attn_bwd_tc_long",
    "vitb_p2p_pad_words", "vitb_p2p_allreduce",
]
