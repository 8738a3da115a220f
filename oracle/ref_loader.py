"""Loads the UNMODIFIED reference modules from /root/reference by file path (build container only —
the GPU box has no /root/reference; tests that need it skip there).  Both reference directories
define modules named `model`, so each is loaded under its own name."""
import importlib.util
import os
import sys

REF_ROOT = os.environ.get("VITB_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.exists(os.path.join(REF_ROOT, "src", "model.py"))


def load_src_model():
    spec = importlib.util.spec_from_file_location("ref_src_model", os.path.join(REF_ROOT, "src", "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_resvit_model():
    d = os.path.join(REF_ROOT, "res-vit")
    sys.path.insert(0, d)          # res-vit/model.py does `from model_utils import ...`
    try:
        for name in ("model_utils",):
            sys.modules.pop(name, None)
        spec = importlib.util.spec_from_file_location("ref_resvit_model", os.path.join(d, "model.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mu = sys.modules.get("model_utils")
    finally:
        sys.path.remove(d)
    return mod, mu
