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
EPI_NONE, EPI_GELU, EPI_GELU_BWD = 0, 1, 2


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
    ]


def _sig(name, argtypes, restype=C.c_int):
    fn = getattr(_lib, name)
    fn.argtypes = argtypes
    fn.restype = restype
    return fn


vitb_version = _sig("vitb_version", [])
_vitb_last_error = _sig("vitb_last_error", [C.c_char_p, C.c_size_t])
vitb_device_check = _sig("vitb_device_check", [])
_vitb_gemm = _sig("vitb_gemm", [C.POINTER(GemmParams), C.c_void_p])


def last_error():
    buf = C.create_string_buffer(1024)
    _vitb_last_error(buf, 1024)
    return buf.value.decode(errors="replace")


def check(status, what):
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
