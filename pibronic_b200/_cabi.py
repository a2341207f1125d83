"""ctypes binding of the pbx C ABI (include/pbx.h) -- the only way Python reaches the CUDA kernels.

There is no CPU fallback: if ``_pbx.so`` is missing, or no CUDA device is visible when a plan is
created, an exception is raised.  torch is used by callers only to own device buffers and to name
the current CUDA stream; no torch type crosses this boundary.
"""
import ctypes as C
import os
from os.path import abspath, dirname, join

import numpy as np

# PBX_LIB: developer override used by the kernel experiments in tools/ (a variant build of the same sources)
_LIB_PATH = os.environ.get("PBX_LIB") or join(dirname(abspath(__file__)), "_pbx.so")

OK = 0
FLAG_PM = 1 << 0
QUIRK_RHO_TRUNC = 1 << 1
FLAG_M_TAU_PM = 1 << 2
FLAG_EIG_JACOBI = 1 << 3
FLAG_FORCE_GENERIC = 1 << 4
FLAG_NO_SCALING = 1 << 5
FLAG_NO_WARPSPEC = 1 << 6
FLAG_NO_FUSED_DMMA = 1 << 7
FLAG_PREFER_DMMA = 1 << 8
QUIRK_RHO_DOUBLE_SHIFT = 1 << 9
PATH_GENERIC, PATH_REGISTER, PATH_BLOCKED, PATH_FUSED_DMMA = 0, 1, 2, 3
NSUMS = 8
SUM_NAMES = ("r", "r_plus", "r_minus", "r_sq", "d1", "d2", "d1_sq", "d2_sq")
NSTATS = 10
STAT_NAMES = ("Z", "Z error", "E", "E error", "Cv", "Cv error", "jk_E", "jk_E error", "jk_Cv", "jk_Cv error")

_dp = C.POINTER(C.c_double)


class PbxModel(C.Structure):
    _fields_ = [("A", C.c_int32), ("N", C.c_int32), ("energy", _dp), ("omega", _dp), ("linear", _dp),
                ("quadratic", _dp)]


class PbxRho(C.Structure):
    _fields_ = [("A", C.c_int32), ("N", C.c_int32), ("energy", _dp), ("omega", _dp), ("linear", _dp)]


class PbxError(RuntimeError):
    pass


_lib = None


def library_path():
    return _LIB_PATH


def lib():
    """the loaded shared library (built by ``python -m pibronic_b200.build`` / ``__graft_entry__.build()``)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(_LIB_PATH):
        raise PbxError(f"CUDA extension {_LIB_PATH} is missing: run `python -m pibronic_b200.build` "
                       "(pibronic_b200 has no CPU fallback)")
    L = C.CDLL(_LIB_PATH)
    vp, i32, i64, u32, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
    sigs = {
        "pbx_abi_version": (C.c_int, []),
        "pbx_library_features": (C.c_int, []),
        "pbx_last_error": (C.c_char_p, []),
        "pbx_device_count": (C.c_int, []),
        "pbx_plan_create": (C.c_int, [C.POINTER(PbxModel), C.POINTER(PbxRho), i32, dbl, dbl, u32, i32, C.POINTER(vp)]),
        "pbx_plan_destroy": (C.c_int, [vp]),
        "pbx_plan_table": (i64, [vp, C.c_char_p, _dp, i64]),
        "pbx_has_register_kernel": (C.c_int, [i32, i32, i32]),
        "pbx_register_shape_library": (C.c_int, [C.c_char_p]),
        "pbx_plan_is_fast": (C.c_int, [vp]),
        "pbx_plan_kernel_path": (C.c_int, [vp]),
        "pbx_plan_launch_count": (i64, [vp]),
        "pbx_plan_launch_param_bytes": (i64, [vp]),
        "pbx_sample_eval_dev": (C.c_int, [vp, u64, i64, i64, vp, vp]),
        "pbx_sample_eval_host": (C.c_int, [vp, u64, i64, i64, vp, i64, i64, vp]),
        "pbx_eval_coords_dev": (C.c_int, [vp, vp, i64, vp, vp]),
        "pbx_eval_coords_host": (C.c_int, [vp, vp, i64, vp, i64]),
        "pbx_sample_coords_dev": (C.c_int, [vp, u64, i64, i64, vp, vp, vp]),
        "pbx_eval_stages_dev": (C.c_int, [vp, vp, i64, vp, vp, vp, vp, vp, vp]),
        "pbx_chain_trace_dev": (C.c_int, [vp, vp, vp, i64, vp, vp]),
        "pbx_block_sums_dev": (C.c_int, [vp, vp, i64, i64, vp, vp]),
        "pbx_stats_dev": (C.c_int, [vp, vp, i64, _dp, vp]),
        "pbx_stats_host": (C.c_int, [vp, vp, i64, i64, _dp]),
        "pbx_stats_last": (C.c_int, [vp, _dp]),
        "pbx_stats_arrays_dev": (C.c_int, [vp, i64, dbl, dbl, i32, _dp, vp]),
        "pbx_stats_arrays_host": (C.c_int, [vp, i64, i64, dbl, dbl, i32, _dp]),
        "pbx_math_probe_dev": (C.c_int, [i32, vp, vp, i64, vp]),
        "pbx_fp64_peak_tflops": (C.c_int, [i32, _dp]),
        "pbx_fp64_peak_tflops_kind": (C.c_int, [i32, i32, _dp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


EXPORTED_SYMBOLS = ("pbx_abi_version", "pbx_library_features", "pbx_last_error", "pbx_device_count", "pbx_plan_create", "pbx_plan_destroy",
                    "pbx_has_register_kernel", "pbx_register_shape_library", "pbx_plan_table", "pbx_plan_is_fast", "pbx_plan_kernel_path", "pbx_plan_launch_count", "pbx_plan_launch_param_bytes", "pbx_sample_eval_dev",
                    "pbx_sample_eval_host", "pbx_eval_coords_dev", "pbx_eval_coords_host", "pbx_sample_coords_dev",
                    "pbx_eval_stages_dev", "pbx_chain_trace_dev", "pbx_block_sums_dev", "pbx_stats_dev", "pbx_stats_host",
                    "pbx_stats_last", "pbx_stats_arrays_dev", "pbx_stats_arrays_host", "pbx_math_probe_dev", "pbx_fp64_peak_tflops", "pbx_fp64_peak_tflops_kind")


def _check(rc):
    if rc != OK:
        raise PbxError(f"pbx error {rc}: {lib().pbx_last_error().decode()}")


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _devptr(t):
    """raw device address of a torch CUDA tensor (or an int / None passed through)"""
    if t is None:
        return None
    if isinstance(t, int):
        return t
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return t.data_ptr()


def _stream_handle(stream):
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    return getattr(stream, "cuda_stream", stream)


class Plan:
    """One (model, rho, beads, beta) evaluation plan on one GPU: owns the device constant tables."""

    def __init__(self, energy, omega, linear, quadratic, rho_energy, rho_omega, rho_linear, beads, beta,
                 delta_beta, flags=FLAG_PM, device=0, jit=None):
        """jit=True (or PBX_JIT=1 in the environment): a shape without a compiled register-resident kernel gets one
        built with nvcc on first use (pibronic_b200/jit.py; cached next to the package), if it is small enough"""
        L = lib()
        self._handle = None
        energy, omega = _f64(energy), _f64(omega)
        A, N = energy.shape[0], omega.shape[0]
        assert energy.shape == (A, A)
        linear = None if linear is None else _f64(linear)
        quadratic = None if quadratic is None else _f64(quadratic)
        if linear is not None:
            assert linear.shape == (N, A, A)
        if quadratic is not None:
            assert quadratic.shape == (N, N, A, A)
        rho_energy, rho_omega = _f64(rho_energy), _f64(rho_omega)
        Ar = rho_energy.shape[0]
        rho_linear = None if rho_linear is None else _f64(rho_linear)
        if rho_linear is not None:
            assert rho_linear.shape == (rho_omega.shape[0], Ar)
        if (os.environ.get("PBX_JIT", "0") == "1") if jit is None else jit:
            from . import jit as _jit
            _jit.ensure_shape(A, N, Ar)
        vib = PbxModel(A, N, _ptr(energy), _ptr(omega), _ptr(linear), _ptr(quadratic))
        rho = PbxRho(Ar, rho_omega.shape[0], _ptr(rho_energy), _ptr(rho_omega), _ptr(rho_linear))
        handle = C.c_void_p()
        _check(L.pbx_plan_create(C.byref(vib), C.byref(rho), int(beads), float(beta), float(delta_beta),
                                 int(flags), int(device), C.byref(handle)))
        self._handle = handle
        self.A, self.N, self.Ar, self.P = A, N, Ar, int(beads)
        self.flags, self.device = int(flags), int(device)
        self.pm = bool(flags & FLAG_PM)
        self.delta_beta = float(delta_beta)

    def close(self):
        if self._handle is not None and _lib is not None:
            _lib.pbx_plan_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection
    @property
    def is_fast(self):
        return bool(lib().pbx_plan_is_fast(self._handle))

    @property
    def kernel_path(self):
        """PATH_GENERIC / PATH_REGISTER / PATH_BLOCKED / PATH_FUSED_DMMA"""
        return int(lib().pbx_plan_kernel_path(self._handle))

    @property
    def launch_count(self):
        return int(lib().pbx_plan_launch_count(self._handle))

    @property
    def launch_param_bytes(self):
        return int(lib().pbx_plan_launch_param_bytes(self._handle))

    def table(self, name):
        n = lib().pbx_plan_table(self._handle, name.encode(), None, 0)
        if n < 0:
            _check(int(n))
        out = np.empty(n, dtype=np.float64)
        lib().pbx_plan_table(self._handle, name.encode(), _ptr(out), n)
        return out

    # ---- device entry points (torch CUDA tensors, float64)
    def sample_eval(self, seed, first_sample, n_samples, out4, stream=None):
        _check(lib().pbx_sample_eval_dev(self._handle, int(seed), int(first_sample), int(n_samples), _devptr(out4),
                                         _stream_handle(stream)))

    def eval_coords(self, R, out4, stream=None):
        n = R.shape[0]
        assert tuple(R.shape[1:]) == (self.N, self.P), f"coordinates must be (X, {self.N}, {self.P})"
        _check(lib().pbx_eval_coords_dev(self._handle, _devptr(R), int(n), _devptr(out4), _stream_handle(stream)))

    def sample_coords(self, seed, first_sample, n_samples, R, src=None, stream=None):
        _check(lib().pbx_sample_coords_dev(self._handle, int(seed), int(first_sample), int(n_samples), _devptr(R),
                                           _devptr(src), _stream_handle(stream)))

    def eval_stages(self, R, o_rho=None, o_vib=None, scale=None, v_mat=None, m_mat=None, stream=None):
        _check(lib().pbx_eval_stages_dev(self._handle, _devptr(R), int(R.shape[0]), _devptr(o_rho), _devptr(o_vib),
                                         _devptr(scale), _devptr(v_mat), _devptr(m_mat), _stream_handle(stream)))

    def chain_trace(self, m_mat, o_diag, g_out, stream=None):
        _check(lib().pbx_chain_trace_dev(self._handle, _devptr(m_mat), _devptr(o_diag), int(m_mat.shape[0]),
                                         _devptr(g_out), _stream_handle(stream)))

    def block_sums(self, out4, n_samples, block_size, sums, stream=None):
        _check(lib().pbx_block_sums_dev(self._handle, _devptr(out4), int(n_samples), int(block_size), _devptr(sums),
                                        _stream_handle(stream)))

    # ---- statistics (Z, E, Cv + jackknife, without the sampling model's harmonic contribution)
    def _stats_dict(self, values):
        return dict(zip(STAT_NAMES, (float(v) for v in values)))

    def stats(self, out4, n_samples=None, stream=None):
        """from a device [4][n] tensor"""
        values = (C.c_double * NSTATS)()
        n = int(out4.shape[1] if n_samples is None else n_samples)
        _check(lib().pbx_stats_dev(self._handle, _devptr(out4), n, values, _stream_handle(stream)))
        return self._stats_dict(values)

    def stats_host(self, out4):
        """from a host (4, n) array with contiguous rows"""
        out4, ld = self._host_out(out4, out4.shape[1])
        assert out4.shape[0] == 4
        values = (C.c_double * NSTATS)()
        _check(lib().pbx_stats_host(self._handle, out4.ctypes.data, int(ld), int(out4.shape[1]), values))
        return self._stats_dict(values)

    def stats_last(self):
        """of the most recent *_host call (its results are still on the device)"""
        values = (C.c_double * NSTATS)()
        _check(lib().pbx_stats_last(self._handle, values))
        return self._stats_dict(values)

    # ---- host entry points (numpy arrays; copies happen inside the library)
    @staticmethod
    def _host_out(out4, n):
        if out4 is None:
            out4 = np.full((4, n), np.nan)
        assert out4.dtype == np.float64 and out4.ndim == 2 and out4.shape[0] in (2, 4) and out4.shape[1] >= n
        if out4.shape[1] == 0:
            return out4, max(n, 1)
        assert out4.strides[1] == 8 and out4.strides[0] % 8 == 0, "rows must be contiguous float64"
        return out4, out4.strides[0] // 8

    def sample_eval_host(self, seed, first_sample, n_samples, out4=None, block_size=None):
        """rows rho, g[, g+, g-] of `out4` (row stride free) are filled for n_samples samples;
        with block_size also returns the (blocks, NSUMS) per-block sums"""
        out4, ld = self._host_out(out4, n_samples)
        assert out4.shape[0] == 4 or not self.pm
        sums = None
        if block_size is not None:
            sums = np.zeros((-(-n_samples // block_size), NSUMS))
        _check(lib().pbx_sample_eval_host(self._handle, int(seed), int(first_sample), int(n_samples),
                                          out4.ctypes.data, int(ld), int(block_size or 0),
                                          sums.ctypes.data if sums is not None else None))
        return out4 if sums is None else (out4, sums)

    def eval_coords_host(self, R, out4=None):
        R = _f64(R)
        n = R.shape[0]
        assert R.shape[1:] == (self.N, self.P), f"coordinates must be (X, {self.N}, {self.P})"
        out4, ld = self._host_out(out4, n)
        assert out4.shape[0] == 4 or not self.pm
        _check(lib().pbx_eval_coords_host(self._handle, R.ctypes.data, int(n), out4.ctypes.data, int(ld)))
        return out4


def stats_arrays_host(out4, beta, delta_beta, device=0):
    """Z, E, Cv + jackknife (STAT_NAMES) of a host (4, n) array; no plan needed: only beta and delta_beta enter"""
    assert out4.dtype == np.float64 and out4.ndim == 2 and out4.shape[0] == 4 and out4.strides[1] == 8
    values = (C.c_double * NSTATS)()
    _check(lib().pbx_stats_arrays_host(out4.ctypes.data, out4.strides[0] // 8, int(out4.shape[1]), float(beta),
                                       float(delta_beta), int(device), values))
    return dict(zip(STAT_NAMES, (float(v) for v in values)))


def stats_arrays_dev(out4, beta, delta_beta, stream=None):
    """the same from a contiguous CUDA tensor (4, n)"""
    values = (C.c_double * NSTATS)()
    _check(lib().pbx_stats_arrays_dev(_devptr(out4), int(out4.shape[1]), float(beta), float(delta_beta),
                                      int(out4.device.index or 0), values, _stream_handle(stream)))
    return dict(zip(STAT_NAMES, (float(v) for v in values)))


FEATURE_MTAU = 1


def has_feature(bit):
    return bool(lib().pbx_library_features() & bit)


def device_count():
    return int(lib().pbx_device_count())


def math_probe(kind, x, out, stream=None):
    """device self-test of log_pos / sqrt_pos / exp_fast / sincos_2pi (kinds 0..4) on CUDA tensors"""
    _check(lib().pbx_math_probe_dev(int(kind), _devptr(x), _devptr(out), int(x.numel()), _stream_handle(stream)))


def fp64_peak_tflops(device=0, kind=0):
    """measured FP64 FMA rate: kind 0 vector DFMA chains, 1 tensor DMMA (mma.sync m8n8k4), 2 the larger of the two"""
    out = C.c_double(0.0)
    _check(lib().pbx_fp64_peak_tflops_kind(int(device), int(kind), C.byref(out)))
    return out.value
