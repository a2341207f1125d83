"""Synthetic vibronic models for the benchmark configurations of BASELINE.json / SURVEY.md section 8(d).

Parameter ranges follow the reference's artificial systems
(examples/artificial_systems/generate_initial_model_parameters.py:18-32, input_json/model_4x6.json):
a displaced-oscillator model that is diagonal in the surfaces is rotated by a fixed orthogonal
matrix, then a small symmetric quadratic coupling is added.  The draw order is part of the
benchmark definition -- do not reorder the rng calls.
"""
import os

import numpy as np
import scipy.linalg

from . import model_io as vIO
from .model_io import VMK

BENCH_SEED = 20260417


def coupled_model(A, N, w_range, e_range, seed=BENCH_SEED, linear=0.2, mixing=0.25, quadratic=0.05):
    """returns a model dictionary (VMK keys) with linear + quadratic coupling"""
    rng = np.random.default_rng(seed)
    w = np.linspace(w_range[0], w_range[1], N)
    E = np.diag(np.linspace(e_range[0], e_range[1], A))
    L = np.zeros((N, A, A))
    for n in range(N):
        L[n] = np.diag(rng.uniform(-linear, linear, A))
    K = rng.uniform(-1, 1, (A, A))
    U = scipy.linalg.expm(mixing * (K - K.T))
    E = U @ E @ U.T
    L = np.einsum('bj,ajk,ck->abc', U, L, U)
    Q = np.zeros((N, N, A, A))
    for n in range(N):
        for m in range(n, N):
            S = rng.uniform(-1, 1, (A, A))
            S = np.tril(S) + np.tril(S, -1).T
            Q[n, m] = Q[m, n] = quadratic * np.sqrt(w[n] * w[m]) / N * S
    return {VMK.N: N, VMK.A: A, VMK.E: E, VMK.w: w, VMK.G1: L, VMK.G2: Q}


def model_c2(seed=BENCH_SEED):
    """A=4 surfaces, N=6 modes (BASELINE.json configs[1]; run with P=64, X=1e6, T=300 K)"""
    return coupled_model(4, 6, (0.14, 0.45), (10.3, 10.9), seed)


def model_c4(seed=BENCH_SEED):
    """A=12 surfaces, N=24 modes (BASELINE.json configs[3]; run with P=256, X=1e7 over 8 GPUs)"""
    return coupled_model(12, 24, (0.1, 0.39), (14.0, 14.8), seed)


def diagonal_of(model):
    """the sampling model made of the surface-diagonal part of a coupled model
    (what vIO.create_basic_diagonal_model writes)"""
    out = {VMK.N: model[VMK.N], VMK.A: model[VMK.A], VMK.w: model[VMK.w].copy()}
    for key in (VMK.E, VMK.G1, VMK.G2):
        if key in model:
            v = model[key]
            out[key] = np.diagonal(v, axis1=v.ndim-2, axis2=v.ndim-1).copy()
    return out


def write_data_set(FS, model, rho_model=None):
    """writes coupled_model.json and sampling_model.json into a FileStructure; returns the two paths"""
    vIO.save_model_to_JSON(FS.path_vib_model, model)
    if rho_model is None:
        vIO.create_basic_diagonal_model(FS)
    else:
        vIO.save_diagonal_model_to_JSON(FS.path_rho_model, rho_model)
    assert os.path.isfile(FS.path_rho_model)
    return FS.path_vib_model, FS.path_rho_model
