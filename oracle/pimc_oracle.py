"""CPU oracle for the Pibronic PIMC estimator hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``pibronic_b200/`` may import this module.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the
checker / CPU baseline, never as the thing shipped.

It restates, in plain numpy, the algorithm of the reference's per-block
estimator (all citations are ``/root/reference/pibronic/pimc/pimc.py`` unless
another file is named):

=====================  =====================================================
this module            reference lines it follows
=====================  =====================================================
``beta_of``            ``pibronic/constants.py:24-41``
``fold_vibronic``      ``172-177`` (Delta), ``208-218`` (shift d, zeroing)
``fold_sampling``      ``336-351`` (Delta, mixture weights), ``372-381``
``harmonic_tables``    ``59-89``  (coth, csch, O-matrix prefactor)
``ring_modes``         ``682-689`` (circulant matrix + eigh)
``mode_sigmas``        ``400-404`` (inverse covariance -> standard deviation)
``draw_block``         ``326-334`` (normal draw), ``386-389`` (sources),
                       ``613-631`` (collective -> bead transform + shift)
``o_factors``          ``1087-1129`` (build_o_matrix)
``scale_factors``      ``1076-1084``, ``1062-1066``
``denominator``        ``1132-1136``
``coupling_matrices``  ``1139-1160``
``m_matrices``         ``1171-1187`` (eigh + U exp(-tau lambda) U^T)
``chain_trace``        ``1194-1209`` (bead chain, trace)
``estimate_block``     ``1413-1449`` (order of operations of block_compute_pm)
``run_blocks``         ``1341-1385`` / ``1388-1462``
``sample_scalar``      the same maths written per sample with Python loops
                        -- small cases only
=====================  =====================================================

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this module
against (i) the reference's own known-answer vectors
``tests/pimc/explicit_data/*`` (repacked as ``tests/golden/explicit_kat.npz``)
and (ii) outputs of the *running* reference on seeded inputs for the cases its
tests do not pin (quadratic coupling, the +/- path, S-scaling, A_rho != A),
generated in the build container by ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field

import numpy as np

# pibronic/constants.py:12,24,28
J_PER_EV = np.float64(1.6021766208e-19)
BOLTZMANN_EV = np.float64(1.38064852e-23) / J_PER_EV
DELTA_BETA = 2.0e-4


def beta_of(temperature):
    """1/(kB T) with kB in eV/K (constants.py:38-41)."""
    return 1.0 / (temperature * BOLTZMANN_EV)


# --------------------------------------------------------------------------
# model files (format: SURVEY.md App. D; vibronic_model_io.py:556-650)
# --------------------------------------------------------------------------
_KEY_N, _KEY_A = "number of modes", "number of surfaces"
_KEY_E, _KEY_W = "energies", "frequencies"
_KEY_L, _KEY_Q = "linear couplings", "quadratic couplings"


def _read_json(path):
    with open(path, "r", encoding="UTF8") as fh:
        return json.loads(fh.read())


def load_vibronic_json(path):
    """coupled_model.json -> dict(A, N, E(A,A), w(N), L(N,A,A), Q(N,N,A,A)); absent keys are zeros."""
    raw = _read_json(path)
    A, N = int(raw[_KEY_A]), int(raw[_KEY_N])
    shapes = {_KEY_E: (A, A), _KEY_W: (N,), _KEY_L: (N, A, A), _KEY_Q: (N, N, A, A)}
    out = {"A": A, "N": N}
    for key, name in ((_KEY_E, "E"), (_KEY_W, "w"), (_KEY_L, "L"), (_KEY_Q, "Q")):
        arr = np.zeros(shapes[key])
        if key in raw:
            arr[...] = np.array(raw[key], dtype=np.float64)
        out[name] = arr
    return out


def load_sampling_json(path):
    """sampling_model.json -> dict(A, N, E(A,), w(N), L(N,A)); quadratic terms are read by the reference but never used."""
    raw = _read_json(path)
    A, N = int(raw[_KEY_A]), int(raw[_KEY_N])
    shapes = {_KEY_E: (A,), _KEY_W: (N,), _KEY_L: (N, A)}
    out = {"A": A, "N": N}
    for key, name in ((_KEY_E, "E"), (_KEY_W, "w"), (_KEY_L, "L")):
        arr = np.zeros(shapes[key])
        if key in raw:
            arr[...] = np.array(raw[key], dtype=np.float64)
        out[name] = arr
    return out


# --------------------------------------------------------------------------
# temperature / model precompute
# --------------------------------------------------------------------------
@dataclass
class HarmonicTables:
    """One TemperatureDependentClass (pimc.py:59-89) for a given tau."""
    tau: float
    coth: np.ndarray     # (N,)
    csch: np.ndarray     # (N,)
    prefactor: np.ndarray  # (A,)  exp(-tau*Etilde_a) * sqrt(prod_n csch_n)


def harmonic_tables(omega, tilde_energy, tau):
    coth = np.tanh(tau * omega) ** (-1.0)
    csch = np.sinh(tau * omega) ** (-1.0)
    pref = np.exp(-tau * tilde_energy) * np.prod(csch) ** 0.5
    return HarmonicTables(tau=float(tau), coth=coth, csch=csch, prefactor=pref)


def fold_vibronic(vib):
    """Fold the diagonal linear terms into displaced oscillators (pimc.py:172-177, 208-218).

    Returns (delta(A,), shift d(A,N), tilde_energy(A,), E_off(A,A), L_off(N,A,A))."""
    A = vib["A"]
    E, w, L = vib["E"], vib["w"], vib["L"]
    idx = np.arange(A)
    Ldiag = L[:, idx, idx]                      # (N, A)
    delta = -0.5 * (Ldiag ** 2.0 / w[:, None]).sum(axis=0)
    shift = (-Ldiag / w[:, None]).T.copy()      # (A, N)
    tilde = np.diag(E).copy() + delta
    E_off = E.copy()
    E_off[idx, idx] = 0.0
    L_off = L.copy()
    L_off[:, idx, idx] = 0.0
    return delta, shift, tilde, E_off, L_off


def fold_sampling(rho, beta):
    """pimc.py:336-351, 372-381.  Returns (delta(Ar,), shift(Ar,N), tilde(Ar,), weights(Ar,))."""
    E, w, L = rho["E"], rho["w"], rho["L"]
    delta = -0.5 * (L ** 2.0 / w[:, None]).sum(axis=0)
    tilde = E + delta
    weight = np.exp(-beta * tilde)
    weight /= np.prod(np.sinh((beta * w) / 2.0))
    weight /= weight.sum()
    shift = (-L / w[:, None]).T.copy()
    return delta, shift, tilde, weight


def ring_modes(P):
    """Eigen-system of the ring adjacency matrix circ(0,1,0,...,0,1) (pimc.py:682-689)."""
    assert P >= 3, "circulant matrix requires 3 or more beads"
    C = np.zeros((P, P))
    i = np.arange(P)
    C[i, (i + 1) % P] = 1.0
    C[i, (i - 1) % P] = 1.0
    lam, V = np.linalg.eigh(C, UPLO="L")
    return C, lam, V


def mode_sigmas(rho_tab, lam):
    """sigma[n,k] = (2 coth_n - csch_n lam_k)^(-1/2) (pimc.py:402-404); surface independent."""
    inv_cov = 2.0 * rho_tab.coth[:, None] - rho_tab.csch[:, None] * lam[None, :]
    return np.sqrt(1.0 / inv_cov)


@dataclass
class Tables:
    A: int
    Ar: int
    N: int
    P: int
    beta: float
    delta_beta: float
    tau: float
    # vibronic model (after folding)
    d_vib: np.ndarray
    delta_vib: np.ndarray
    E_off: np.ndarray
    L_off: np.ndarray
    Q: np.ndarray
    vib: HarmonicTables
    vib_plus: HarmonicTables
    vib_minus: HarmonicTables
    # sampling model
    d_rho: np.ndarray
    delta_rho: np.ndarray
    weights: np.ndarray
    rho: HarmonicTables
    ring_matrix: np.ndarray
    ring_eigvals: np.ndarray
    ring_eigvecs: np.ndarray
    sigma: np.ndarray          # (N, P)
    rho_trunc: bool = False    # reference quirk Q1: rho evaluated on the first A surfaces only
    extra: dict = field(default_factory=dict)


def precompute(vib, rho, P, temperature, delta_beta=DELTA_BETA, rho_trunc=False):
    """Everything BoxDataPM.preprocess() prepares (pimc.py:647-692, 714-738, 246-269, 409-424)."""
    assert vib["N"] == rho["N"]
    beta = beta_of(temperature)
    tau = beta / P
    tau_p = (beta + delta_beta) / P
    tau_m = (beta - delta_beta) / P
    delta_v, d_v, tilde_v, E_off, L_off = fold_vibronic(vib)
    delta_r, d_r, tilde_r, weights = fold_sampling(rho, beta)
    C, lam, V = ring_modes(P)
    rho_tab = harmonic_tables(rho["w"], tilde_r, tau)
    return Tables(
        A=vib["A"], Ar=rho["A"], N=vib["N"], P=P, beta=float(beta), delta_beta=float(delta_beta), tau=float(tau),
        d_vib=d_v, delta_vib=delta_v, E_off=E_off, L_off=L_off, Q=vib["Q"].copy(),
        vib=harmonic_tables(vib["w"], tilde_v, tau),
        vib_plus=harmonic_tables(vib["w"], tilde_v, tau_p),
        vib_minus=harmonic_tables(vib["w"], tilde_v, tau_m),
        d_rho=d_r, delta_rho=delta_r, weights=weights, rho=rho_tab,
        ring_matrix=C, ring_eigvals=lam, ring_eigvecs=V, sigma=mode_sigmas(rho_tab, lam),
        rho_trunc=rho_trunc,
    )


# --------------------------------------------------------------------------
# sampler
# --------------------------------------------------------------------------
def draw_sources(tab, X, rng):
    """Mixture component per sample (pimc.py:386-389)."""
    return rng.choice(tab.Ar, size=X, p=tab.weights)


def draw_block(tab, sources, rng):
    """Bead coordinates R (B,N,P) for the given mixture components (pimc.py:326-334, 613-631)."""
    B = len(sources)
    cc = rng.normal(loc=0.0, scale=np.broadcast_to(tab.sigma, (B, tab.N, tab.P)))
    R = np.einsum("ab,ijb->ija", tab.ring_eigvecs, cc)
    R += tab.d_rho[sources][:, :, None]
    return R


# --------------------------------------------------------------------------
# estimator pieces, block-vectorised exactly like the reference
# --------------------------------------------------------------------------
def o_factors(R, shift, tab_t, n_surf=None):
    """Diagonal of the harmonic bead-pair propagator, (B,P,S) (pimc.py:1087-1129).

    R is (B,N,P); shift is (S,N).  n_surf < S reproduces quirk Q1 (trailing surfaces stay zero)."""
    q1 = R[:, None, :, :] - shift[None, :, :, None]           # (B,S,N,P)
    q2 = np.roll(q1, shift=-1, axis=3)
    coth = tab_t.coth[None, None, :, None]
    csch = tab_t.csch[None, None, :, None]
    expo = -0.5 * np.sum(coth * (q1 ** 2.0 + q2 ** 2.0) - 2.0 * csch * q1 * q2, axis=2).swapaxes(1, 2)
    out = np.exp(expo) * tab_t.prefactor[None, None, :]
    if n_surf is not None and n_surf < out.shape[2]:
        out[:, :, n_surf:] = 0.0
    return out


def scale_factors(o_rho, o_vib):
    """S[b,p] = max over surfaces of both models (pimc.py:1076-1084)."""
    return np.maximum(o_rho.max(axis=2), o_vib.max(axis=2))


def denominator(o_rho_scaled):
    """rho(R) = sum_a prod_p O_rho[b,p,a] (pimc.py:1132-1136)."""
    return o_rho_scaled.prod(axis=1).sum(axis=1)


def coupling_matrices(tab, R):
    """V[b,p,i,j] (pimc.py:1147-1160)."""
    V = np.einsum("bnp,nmij,bmp->bpij", R, 0.5 * tab.Q, R)
    V += np.einsum("nij,bnp->bpij", tab.L_off, R)
    V += tab.E_off[None, None, :, :]
    return V


def m_matrices(tab, V, tau=None):
    """M = U exp(-tau lambda) U^T -- always data.tau, also for the +/- variants (pimc.py:1171-1187, quirk Q2).
    `tau` overrides it for the consistent estimator (PBX_FLAG_M_TAU_PM), which the reference does not have."""
    lam, U = np.linalg.eigh(V, UPLO="L")
    return np.einsum("abcd,abd,abed->abce", U, np.exp(-(tab.tau if tau is None else tau) * lam), U, optimize="optimal")


def chain_trace(M, o_diag, faithful=True):
    """g[b] = tr prod_p (M[b,p] . diag(O[b,p])) (pimc.py:1194-1209).

    ``faithful=True`` keeps the reference's per-(b,p) ``ndarray.dot`` double loop with a dense
    diagonal O matrix -- this is what the CPU baseline times.  ``faithful=False`` batches
    over b (same arithmetic order per sample) for the bigger parity tests."""
    B, P, A, _ = M.shape
    if faithful:
        o_dense = np.zeros((B, P, A, A))
        idx = np.arange(A)
        o_dense[:, :, idx, idx] = o_diag
        acc = np.empty((B, A, A))
        for b in range(B):
            acc[b] = np.identity(A)
        for b in range(B):
            for p in range(P):
                acc[b].dot(M[b, p], out=acc[b])
                acc[b].dot(o_dense[b, p], out=acc[b])
        return np.trace(acc, axis1=1, axis2=2)
    acc = np.broadcast_to(np.identity(A), (B, A, A)).copy()
    for p in range(P):
        acc = np.matmul(acc, M[:, p])
        acc = acc * o_diag[:, p, None, :]
    return np.trace(acc, axis1=1, axis2=2)


def estimate_block(tab, R, pm=True, faithful=True, scale=True, details=None, m_tau_pm=False):
    """(rho, g[, g_plus, g_minus]) for bead coordinates R (B,N,P); order of operations of
    block_compute / block_compute_pm (pimc.py:1364-1381, 1420-1449).
    m_tau_pm=True is NOT the reference: g+- then use exp(-tau+- V) (PBX_FLAG_M_TAU_PM)."""
    n_rho = min(tab.A, tab.Ar) if tab.rho_trunc else None
    o_rho = o_factors(R, tab.d_rho, tab.rho, n_rho)
    o_vib = o_factors(R, tab.d_vib, tab.vib)
    if scale:
        S = scale_factors(o_rho, o_vib)
    else:
        S = np.ones(o_rho.shape[:2])
    o_rho = o_rho / S[..., None]
    o_vib = o_vib / S[..., None]
    rho = denominator(o_rho)
    V = coupling_matrices(tab, R)
    M = m_matrices(tab, V)
    g = chain_trace(M, o_vib, faithful)
    if details is not None:
        details.update(o_rho=o_rho, o_vib=o_vib, S=S, V=V, M=M)
    if not pm:
        return rho, g
    o_p = o_factors(R, tab.d_vib, tab.vib_plus) / S[..., None]
    gp = chain_trace(m_matrices(tab, V, tab.vib_plus.tau) if m_tau_pm else M, o_p, faithful)
    o_m = o_factors(R, tab.d_vib, tab.vib_minus) / S[..., None]
    gm = chain_trace(m_matrices(tab, V, tab.vib_minus.tau) if m_tau_pm else M, o_m, faithful)
    return rho, g, gp, gm


def run_blocks(tab, X, block_size, rng, pm=True, faithful=True, keep_R=False):
    """block_compute[_pm] (pimc.py:1341-1462): draw + estimate, block by block."""
    blocks = X // block_size
    sources = draw_sources(tab, X, rng)
    n_out = 4 if pm else 2
    out = np.full((n_out, X), np.nan)
    kept = []
    for blk in range(blocks):
        view = slice(blk * block_size, (blk + 1) * block_size)
        R = draw_block(tab, sources[view], rng)
        if keep_R:
            kept.append(R)
        res = estimate_block(tab, R, pm=pm, faithful=faithful)
        for k in range(n_out):
            out[k, view] = res[k]
    if keep_R:
        return out, np.concatenate(kept, axis=0)
    return out


# --------------------------------------------------------------------------
# per-sample scalar restatement -- small cases only
# --------------------------------------------------------------------------
def sample_scalar(tab, Rx, pm=True):
    """(rho, g, g+, g-) for ONE sample Rx (N,P), written with explicit loops."""
    A, Ar, N, P = tab.A, tab.Ar, tab.N, tab.P
    n_rho = min(A, Ar) if tab.rho_trunc else Ar

    def o_of(shift, t, S):
        o = np.zeros((P, S))
        for p in range(P):
            pn = (p + 1) % P
            for a in range(S):
                acc = 0.0
                for n in range(N):
                    q = Rx[n, p] - shift[a, n]
                    qn = Rx[n, pn] - shift[a, n]
                    acc += t.coth[n] * (q * q + qn * qn) - 2.0 * t.csch[n] * q * qn
                o[p, a] = t.prefactor[a] * np.exp(-0.5 * acc)
        return o

    o_rho = o_of(tab.d_rho, tab.rho, Ar)
    o_rho[:, n_rho:] = 0.0
    variants = [tab.vib] + ([tab.vib_plus, tab.vib_minus] if pm else [])
    o_v = [o_of(tab.d_vib, t, A) for t in variants]
    S = np.maximum(o_rho.max(axis=1), o_v[0].max(axis=1))
    rho = 0.0
    for a in range(Ar):
        prod = 1.0
        for p in range(P):
            prod *= o_rho[p, a] / S[p]
        rho += prod
    Ms = []
    for p in range(P):
        V = tab.E_off.copy()
        for n in range(N):
            V += tab.L_off[n] * Rx[n, p]
            for m in range(N):
                V += 0.5 * tab.Q[n, m] * Rx[n, p] * Rx[m, p]
        lam, U = np.linalg.eigh(V)
        Ms.append((U * np.exp(-tab.tau * lam)[None, :]) @ U.T)
    out = [rho]
    for o in o_v:
        T = np.identity(A)
        for p in range(P):
            T = T @ Ms[p]
            T = T * (o[p] / S[p])[None, :]
        out.append(np.trace(T))
    return tuple(out)


# --------------------------------------------------------------------------
# downstream estimators (pibronic/stats/stats.py:38-55, 84-123) -- used by the
# statistical parity tests, not part of the hot path itself
# --------------------------------------------------------------------------
def property_terms(delta_beta, rho, g, gp, gm):
    ratio = g / rho
    d1 = (gp - gm) / rho / (2.0 * delta_beta)
    d2 = (gp - 2.0 * g + gm) / rho / delta_beta ** 2
    return ratio, d1, d2


def basic_properties(X, T, ratio, d1, d2):
    Z = np.mean(ratio)
    Z_err = np.std(ratio, ddof=0) / np.sqrt(X - 1)
    E = -np.mean(d1) / Z
    Cv = (np.mean(d2) / Z - E ** 2) / (BOLTZMANN_EV * T ** 2)
    return {"Z": Z, "Z error": Z_err, "E": E, "Cv": Cv}
