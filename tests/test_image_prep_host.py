"""CPU checks of the device input transform that need no GPU:

* vitb_resize_tables_host — host code inside libvitb200.so — equals the oracle's tables (Pillow's algorithm);
* the kernel body (csrc/vitb_image_prep_core.h, the code image_prep_kernel runs per CTA) walked block by block and
  thread by thread by the g++ harness under tests/host_harness/, bit-exact against the oracle and the golden vectors
  from the reference's own loaders.  The harness is scaffolding for a GPU-less container; the `-m gpu` tests in
  tests/test_image_prep_gpu.py run the real kernel through the C ABI.
"""
import ctypes as C
import importlib
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import image_prep_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "image_prep.npz")
L = importlib.import_module("vit-of-pytorch_b200._lib")


def lib_tables(n_in, n_out):
    k = C.c_int(0)
    assert L._vitb_resize_tables_host(n_in, n_out, None, None, 0, C.byref(k)) == 0
    bounds = np.zeros((n_out, 2), dtype=np.int32)
    coeffs = np.zeros((n_out, k.value), dtype=np.int32)
    st = L._vitb_resize_tables_host(n_in, n_out, bounds.ctypes.data_as(C.c_void_p), coeffs.ctypes.data_as(C.c_void_p),
                                    coeffs.size, C.byref(k))
    assert st == 0, L.last_error()
    return k.value, bounds, coeffs


@pytest.mark.parametrize("n_in,n_out", [(32, 224), (32, 384), (32, 64), (75, 64), (100, 64), (500, 224), (375, 224),
                                        (7, 224), (224, 32), (33, 32), (1, 5), (5, 1), (1000, 3)])
def test_library_tables_equal_oracle_tables(n_in, n_out):
    k, b, c = lib_tables(n_in, n_out)
    ko, bo, co = O.resample_tables(n_in, n_out)
    assert k == ko
    assert np.array_equal(b, bo)
    assert np.array_equal(c, co)


def test_tables_capacity_is_checked():
    k = C.c_int(0)
    bounds = np.zeros((224, 2), dtype=np.int32)
    coeffs = np.zeros((10,), dtype=np.int32)
    st = L._vitb_resize_tables_host(32, 224, bounds.ctypes.data_as(C.c_void_p), coeffs.ctypes.data_as(C.c_void_p), 10,
                                    C.byref(k))
    assert st < 0 and "capacity" in L.last_error()


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("harness") / "image_prep_host.so"
    src = os.path.join(ROOT, "tests", "host_harness", "image_prep_host.cpp")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(cuda_inc, "cuda_bf16.h")):
        pytest.skip("CUDA headers not found under %s" % cuda_inc)
    res = subprocess.run([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", cuda_inc, src, "-o", str(out)],
                         capture_output=True, text=True)
    assert res.returncode == 0, "host harness failed to compile:\n" + res.stderr[-2000:]
    lib = C.CDLL(str(out))
    lib.image_prep_host.restype = C.c_int
    return lib


def run_harness(lib, x, out_hw, flip=None, P=None, ldk=None, want_lo=False, nthreads=256):
    B, H, W, Cn = x.shape
    oh, ow = out_hw
    vp = C.c_void_p

    def tables(n_in, n_out):
        if n_in == n_out:
            return None, None, 0
        k, b, c = lib_tables(n_in, n_out)
        return b, c, k

    xb, xc, xk = tables(W, ow)
    yb, yc, yk = tables(H, oh)
    lut = np.ascontiguousarray(O.normalize_lut()[:Cn])
    img = np.full((B, Cn, oh, ow), np.nan, dtype=np.float32)
    u8 = np.full((B, oh, ow, Cn), 77, dtype=np.uint8)
    hi = lo = None
    if P:
        gh, gw = oh // P, ow // P
        hi = np.zeros((B * gh * gw, ldk), dtype=np.uint16)
        lo = np.zeros_like(hi) if want_lo else None
    fl = None if flip is None else np.ascontiguousarray(np.asarray(flip).astype(np.uint8))
    p = lambda arr: None if arr is None else arr.ctypes.data_as(vp)
    band, cap = C.c_int(0), C.c_int(0)
    x = np.ascontiguousarray(x)
    st = lib.image_prep_host(p(x), B, H, W, Cn, oh, ow, p(xb), p(xc), xk, p(yb), p(yc), yk, p(fl), p(lut), p(img),
                             P or 0, ldk or 0, p(hi), p(lo), p(u8), nthreads, C.byref(band), C.byref(cap))
    assert st == 0
    return img, u8, hi, lo, band.value, cap.value


def bf16_bits(x):
    return (torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16))


def bf16_lo_bits(x):
    t = torch.from_numpy(np.ascontiguousarray(x))
    return (t - t.to(torch.bfloat16).float()).to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)


@pytest.mark.parametrize("case,src,size", [("cifar_train", "cifar_in", 224), ("cifar_eval", "cifar_in", 64),
                                           ("inet_train", "inet_in", (64, 64))])
def test_kernel_body_matches_reference_loader_output(harness, case, src, size):
    g = np.load(GOLD)
    order = g[case + "_order"]
    flip = g[case + "_flip"] if case + "_flip" in g.files else None
    x = g[src][order]
    oh, ow = O.resize_target(x.shape[1], x.shape[2], size)
    img, u8, _, _, _, _ = run_harness(harness, x, (oh, ow), flip=flip)
    assert np.array_equal(img, g[case + "_out"])


@pytest.mark.parametrize("H,W,out_hw,P,ldk,nthreads", [
    (32, 32, (224, 224), 16, 768, 256),      # CIFAR -> B/16 geometry, vector stores everywhere
    (32, 32, (224, 224), 14, 592, 256),      # H/14: P % 4 != 0 -> scalar patch stores, padded K
    (32, 32, (32, 32), 16, 768, 64),         # no pass at all (both axes keep their size)
    (40, 32, (40, 96), 8, 192, 96),          # horizontal pass only
    (40, 32, (100, 32), 4, 48, 96),          # vertical pass only
    (75, 100, (64, 64), 16, 768, 256),       # downsampling, 5-tap windows
    (333, 500, (61, 45), 7, 152, 128),       # odd sizes: scalar image stores, rows / columns beyond the patch grid
    (9, 700, (300, 10), 5, 80, 32),          # tall output from few rows; wide source
    (1, 1, (17, 19), 4, 48, 32),             # single pixel
])
def test_kernel_body_matches_oracle(harness, H, W, out_hw, P, ldk, nthreads):
    rng = np.random.default_rng(H * 1000 + W)
    B = 3
    x = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    flip = np.array([1, 0, 1], dtype=np.uint8)
    ref_u8, ref = O.image_prep(x, out_hw, flip=flip)
    img, u8, hi, lo, band, cap = run_harness(harness, x, out_hw, flip=flip, P=P, ldk=ldk, want_lo=True, nthreads=nthreads)
    assert np.array_equal(u8, ref_u8)
    assert np.array_equal(img, ref)
    cols = O.patch_columns(ref, P, ldk)
    assert np.array_equal(hi, bf16_bits(cols))
    assert np.array_equal(lo, bf16_lo_bits(cols))
    assert 1 <= band <= 32 and cap <= H


def test_no_flip_pointer_and_partial_outputs(harness):
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, (2, 32, 32, 3), dtype=np.uint8)
    ref_u8, ref = O.image_prep(x, 224)
    img, u8, hi, lo, _, _ = run_harness(harness, x, (224, 224), flip=None, P=16, ldk=768)
    assert np.array_equal(img, ref) and np.array_equal(u8, ref_u8) and lo is None
    assert np.array_equal(hi, bf16_bits(O.patch_columns(ref, 16, 768)))
