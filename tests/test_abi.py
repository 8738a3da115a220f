"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads without a GPU, exports
every symbol include/vitb200.h declares, agrees with the ctypes mirror on struct layout, and the product
path refuses CPU tensors instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vitb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(vitb_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    import vitb200
    lib = ctypes.CDLL(vitb200._lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libvitb200.so does not export %s" % n
    assert sorted(vitb200._lib.EXPORTED_SYMBOLS) == names


def test_struct_layout_matches_ctypes():
    import vitb200
    L = vitb200._lib
    assert L.vitb_struct_size(0) == ctypes.sizeof(L.GemmParams)
    assert L.vitb_struct_size(1) == ctypes.sizeof(L.AttnParams)
    assert L.vitb_version() == 100


def test_library_is_sm100a_only():
    import subprocess
    import vitb200
    out = subprocess.run(["cuobjdump", "--list-elf", vitb200._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_cpu_tensors_are_rejected_not_routed_elsewhere():
    import vitb200
    m = vitb200.VisionTransformer(image_size=(32, 32), patch_size=(16, 16), emb_dim=128, mlp_dim=256, num_heads=2,
                                  num_layers=1, num_classes=10, dropout_rate=0.0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(1, 3, 32, 32))
    with pytest.raises(RuntimeError):
        vitb200.functional.layer_norm(torch.randn(4, 128), torch.ones(128), torch.zeros(128))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "vit-of-pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                # docstrings may NAME the oracle; nothing may import, load or execute it
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), "%s imports the oracle" % f
                assert not re.search(r"""["']oracle["'/]|oracle[./]vit_oracle|oracle/_ref""", src), "%s loads the oracle" % f


def test_missing_library_is_a_loud_error(tmp_path, monkeypatch):
    import importlib.util
    src = os.path.join(ROOT, "vit-of-pytorch_b200", "_lib.py")
    dst = tmp_path / "_lib_copy.py"
    dst.write_text(open(src).read())
    spec = importlib.util.spec_from_file_location("_lib_copy", str(dst))
    mod = importlib.util.module_from_spec(spec)
    with pytest.raises(RuntimeError, match="libvitb200.so is missing"):
        spec.loader.exec_module(mod)


def test_arch_presets_match_reference_table():
    import vitb200
    # src/config.py:57-104
    assert vitb200.get_arch("b16") == dict(patch_size=16, emb_dim=768, mlp_dim=3072, num_heads=12, num_layers=12,
                                           attn_dropout_rate=0.0, dropout_rate=0.0)
    assert vitb200.get_arch("h14")["emb_dim"] == 1280 and vitb200.get_arch("h14")["num_layers"] == 32
    assert vitb200.get_arch("l32")["patch_size"] == 32 and vitb200.get_arch("l16")["num_heads"] == 16
    assert sorted(vitb200.ARCHS) == ["b16", "b32", "h14", "l16", "l32"]
