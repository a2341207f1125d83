"""Register-resident kernels for model shapes that were not compiled into ``_pbx.so``.

The one-sample-per-thread kernels (csrc/pbx_fast.cuh, pbx_fast_ws.cuh) need the number of surfaces, modes and sampling
surfaces at compile time; ``csrc/shapes.def`` lists the shapes the library ships with.  Any other shape runs, by
default, on the fused tensor-core kernel (csrc/pbx_big.cuh: one launch, any 1 <= A <= 16).  For SMALL shapes the
register-resident form is faster; ``ensure_shape(A, N, A_rho)`` compiles it on demand:

    nvcc -shared ... -DPBX_A=.. -DPBX_N=.. -DPBX_AR=.. -DPBX_JIT_LIBRARY csrc/pbx_fast_inst.cu -o _jit/pbx_fast_A_N_AR_<digest>.so

(the same translation unit the build uses per shape, without the Jacobi cross-check variants) and registers it with
``pbx_register_shape_library``; plans created afterwards pick it up.  The library is cached next to the package under
``_jit/`` and keyed by the digest of the sources, so a stale one is never loaded.  ``_cabi.Plan(..., jit=True)`` or
``PBX_JIT=1`` call this automatically.
"""
import os
import subprocess
from os.path import isfile, join

from . import _cabi, build as pbx_build

JIT_DIR = join(pbx_build.PKG, "_jit")
MAX_SURFACES, MAX_MODES, MAX_RHO_SURFACES = 5, 12, 8       # beyond this the per-thread state no longer fits the register file

_registered = set()


def eligible(A, N, Ar):
    return 1 <= A <= MAX_SURFACES and 1 <= N <= MAX_MODES and 1 <= Ar <= MAX_RHO_SURFACES


def ensure_shape(A, N, Ar, verbose=False):
    """True if a register-resident kernel for (A, N, A_rho) is available after the call (built in, cached, or just compiled)"""
    L = _cabi.lib()
    A, N, Ar = int(A), int(N), int(Ar)
    if L.pbx_has_register_kernel(A, N, Ar):
        return True
    if not eligible(A, N, Ar):
        return False
    os.makedirs(JIT_DIR, exist_ok=True)
    digest = pbx_build._source_digest()[:16]
    path = join(JIT_DIR, f"pbx_fast_{A}_{N}_{Ar}_{digest}.so")
    if not isfile(path):
        cmd = [pbx_build.NVCC, *pbx_build.ARCH, *pbx_build.CFLAGS, "-shared", "-cudart", "static", "-DPBX_JIT_LIBRARY", "-DPBX_WITH_JACOBI=0",
               f"-DPBX_A={A}", f"-DPBX_N={N}", f"-DPBX_AR={Ar}", join(pbx_build.CSRC, "pbx_fast_inst.cu"), "-o", path + ".tmp"]
        if verbose:
            print("pibronic_b200.jit:", " ".join(cmd))
        try:
            proc = subprocess.run(cmd, capture_output=True, text=True)
        except FileNotFoundError:
            return False                         # no nvcc here: the shape stays on the fused tensor-core kernel
        if proc.returncode != 0:
            raise _cabi.PbxError(f"run-time compilation of the ({A}, {N}, {Ar}) kernel failed:\n{proc.stderr[-2000:]}")
        os.replace(path + ".tmp", path)
    if path not in _registered:
        _cabi._check(L.pbx_register_shape_library(path.encode()))
        _registered.add(path)
    return bool(L.pbx_has_register_kernel(A, N, Ar))
