"""Builds the CUDA extension ``pibronic_b200/_pbx.so`` in-tree with nvcc for sm_100a.

    python -m pibronic_b200.build [--force]

One translation unit per kernel shape (csrc/shapes.def) compiled in parallel; objects go to
``build/pbx/``, the shared library next to the package so that it travels with the tree.
"""
import hashlib
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from os.path import abspath, dirname, join

PKG = dirname(abspath(__file__))
ROOT = dirname(PKG)
CSRC = join(PKG, "csrc")
# developer knobs for kernel experiments (tools/): another library name, extra -D flags, fewer shapes
_VARIANT = os.environ.get("PBX_VARIANT", "")
OUT = join(PKG, f"_pbx{('_' + _VARIANT) if _VARIANT else ''}.so")
OBJ_DIR = join(ROOT, "build", "pbx" + (("_" + _VARIANT) if _VARIANT else ""))

NVCC = os.environ.get("NVCC", "nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# PBX_WITH_MTAU=1: also build the PBX_FLAG_M_TAU_PM kernels (consistent estimator, not part of the reference's path)
# -Xfatbin=-compress-all: the device code of ~300 kernels (with line info) shrinks 3x; cuobjdump / ncu read it as before
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xfatbin=-compress-all"] + \
    (["-DPBX_WITH_MTAU=1"] if os.environ.get("PBX_WITH_MTAU") == "1" else []) + \
    os.environ.get("PBX_EXTRA_NVCC_FLAGS", "").split()


BIG_SURFACES = range(1, 17)     # keep in step with PBX_BIG_LIST in csrc/pbx_api.cu


def shapes():
    with open(join(CSRC, "shapes.def")) as fh:
        return [tuple(int(v) for v in m.groups())
                for m in re.finditer(r"^PBX_SHAPE\((\d+),\s*(\d+),\s*(\d+)\)", fh.read(), re.M)]


def _source_digest():
    h = hashlib.sha256()
    for base in (CSRC, join(ROOT, "include")):
        for name in sorted(os.listdir(base)):
            with open(join(base, name), "rb") as fh:
                h.update(name.encode())
                h.update(fh.read())
    h.update(" ".join(ARCH + CFLAGS).encode())
    return h.hexdigest()


def _run(cmd):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("command failed: {}\n{}\n{}".format(" ".join(cmd), proc.stdout, proc.stderr))
    return proc.stdout + proc.stderr


def build(force=False, verbose=False):
    """compile (if sources changed) and return the path of the shared library"""
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = join(OBJ_DIR, "digest.txt")
    digest = _source_digest()
    if not force and os.path.isfile(OUT) and os.path.isfile(stamp) and open(stamp).read() == digest:
        return OUT
    jobs = []
    # PBX_VARIANT_ONLY_FAST=1: a variant build recompiles the register-resident shapes only and links the main build's
    # other objects (kernel experiments on the small shapes in a minute instead of ten)
    main_dir = join(ROOT, "build", "pbx")
    reuse = bool(_VARIANT) and os.environ.get("PBX_VARIANT_ONLY_FAST") == "1"
    for name in ("pbx_api.cu", "pbx_fast_registry.cu"):
        obj = join(main_dir if reuse else OBJ_DIR, name.replace(".cu", ".o"))
        jobs.append((obj, None if reuse else [NVCC, *ARCH, *CFLAGS, "-c", join(CSRC, name), "-o", obj]))
    for (A, N, AR) in shapes():
        obj = join(OBJ_DIR, f"pbx_fast_{A}_{N}_{AR}.o")
        jobs.append((obj, [NVCC, *ARCH, *CFLAGS, f"-DPBX_A={A}", f"-DPBX_N={N}", f"-DPBX_AR={AR}",
                           "-c", join(CSRC, "pbx_fast_inst.cu"), "-o", obj]))
    for A in BIG_SURFACES:      # fused large-A kernel, one translation unit per number of surfaces
        obj = join(main_dir if reuse else OBJ_DIR, f"pbx_big_{A}.o")
        jobs.append((obj, None if reuse else [NVCC, *ARCH, *CFLAGS, f"-DPBX_BIG_AT={A}", "-c", join(CSRC, "pbx_big_inst.cu"), "-o", obj]))
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        logs = list(pool.map(lambda job: _run(job[1]) if job[1] else "", jobs))
    if verbose:
        print("\n".join(logs))
    _run([NVCC, *ARCH, "-shared", "-o", OUT, *[obj for obj, _ in jobs], "-cudart", "static", "-ldl"])
    with open(stamp, "w") as fh:
        fh.write(digest)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
