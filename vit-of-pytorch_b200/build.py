"""In-tree build of libvitb200.so (sm_100a only) with nvcc.

The shared library lands next to this file so that it travels with the repo snapshot to the GPU
box; it is git-ignored.  `python vit-of-pytorch_b200/build.py` or `__graft_entry__.build()`.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libvitb200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    # host code only: the resize tables are double arithmetic that must round exactly as Pillow's C does
    "-Xcompiler", "-ffp-contract=off",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC)")


# Diagnostics library (NOT part of the product): objects compiled from the same sources with other macros and linked into
# libvitb200_tools.so by build_tools() — (source, object stem, extra flags)
TOOL_VARIANTS = [
    # ablation build of the GEMM (entry points vitb_gemm_diag / vitb_gemm_diag_mask, tools/epi_ablate.py)
    ("vitb_gemm.cu", "vitb_gemm_diag", ["-DVITB_GEMM_DIAG=1"]),
]
TOOLS_LIB = os.path.join(HERE, "libvitb200_tools.so")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path, flags):
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    for dep in [path] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + [
        os.path.join(INCLUDE, "vitb200.h")
    ]:
        with open(dep, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _compile_one(nvcc, src, verbose, stem=None, extra=()):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, (stem or src[:-3]) + ".o")
    stamp = obj + ".sha"
    dig = _digest(path, NVCC_FLAGS + list(extra))
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, ""
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + ["-I", INCLUDE, "-c", path, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
    with open(stamp, "w") as fh:
        fh.write(dig)
    return obj, res.stderr if verbose else ""


def build(verbose=False, force=False):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    jobs = [(s, None, ()) for s in sources()]
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        results = list(ex.map(lambda j: _compile_one(nvcc, j[0], verbose, j[1], j[2]), jobs))
    objs = [r[0] for r in results]
    for _, log in results:
        if log:
            sys.stderr.write(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return LIB


def build_tools(verbose=False):
    """libvitb200_tools.so: the ablation build of the GEMM plus the library-level objects it needs (error text, TMA
    descriptor encoding).  Loaded only by tools/epi_ablate.py."""
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    objs = [_compile_one(nvcc, src, verbose, stem, tuple(extra))[0] for src, stem, extra in TOOL_VARIANTS]
    objs.append(_compile_one(nvcc, "vitb_api.cu", verbose)[0])
    res = subprocess.run([nvcc, "-shared", "-o", TOOLS_LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return TOOLS_LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
    if "--tools" in sys.argv:
        print(build_tools(verbose="-v" in sys.argv))
