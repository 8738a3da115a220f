"""Loads the UNMODIFIED reference modules by file path: from /root/reference in the build container, from the
byte-identical staged copy baseline/_ref/ (git-ignored, shipped with the snapshot) on the GPU box.  Both reference
directories define modules named `model`, so each is loaded under its own name."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")      # unmodified copies staged by __graft_entry__.build()


def _root():
    env = os.environ.get("VITB_REFERENCE_ROOT")
    if env:
        return env
    if os.path.exists("/root/reference/src/model.py"):
        return "/root/reference"
    return _STAGED


REF_ROOT = _root()


def stage(dst=_STAGED, src="/root/reference"):
    """Copy the reference's model files, byte for byte, into the git-ignored baseline/_ref/ so that they travel to the GPU
    box with the snapshot (bench.py --impl reference times THEM, not a restatement).  No-op without /root/reference."""
    import shutil
    files = [("src", "model.py"), ("res-vit", "model.py"), ("res-vit", "model_utils.py")]
    if not all(os.path.exists(os.path.join(src, d, f)) for d, f in files):
        return False
    for d, f in files:
        os.makedirs(os.path.join(dst, d), exist_ok=True)
        shutil.copyfile(os.path.join(src, d, f), os.path.join(dst, d, f))
    return True


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
