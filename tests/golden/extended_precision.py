"""Extended-precision (x87 80-bit long double, eps = 1.1e-19) evaluation of the estimator of SURVEY.md
App. A for one run: the yardstick for cases where the reference's own float64 result is ill-conditioned
(the chained product of 128 bead matrices amplifies rounding by 1e5..1e6 for some paper models, so the
reference itself is only good to ~1e-10 there).  Test infrastructure: used by make_c3_paper.py to store
`exact` next to the reference's outputs.  numpy only, no LAPACK (expm by scaling and squaring)."""
import numpy as np

LD = np.longdouble


def _expm_ld(X):
    """exp of a stack of small matrices (..., A, A) in long double: 2^-12 scaling, degree-18 Taylor, 12 squarings"""
    A = X.shape[-1]
    Y = X / LD(4096)
    eye = np.broadcast_to(np.eye(A, dtype=LD), X.shape)
    term, out = eye.copy(), eye.copy()
    for k in range(1, 19):
        term = np.matmul(term, Y) / LD(k)
        out = out + term
    for _ in range(12):
        out = np.matmul(out, out)
    return out


def evaluate(vib, rho, P, T, R, delta_beta=2e-4, rho_trunc=True):
    """returns [4][X] long double: scaled rho, g, g+, g- exactly as the reference defines them"""
    A, N, Ar = int(vib["A"]), int(vib["N"]), int(rho["A"])
    kB = LD(1.38064852e-23) / LD(1.6021766208e-19)        # the reference forms k_B in float64 (constants.py:24)
    kB = LD(np.float64(1.38064852e-23) / np.float64(1.6021766208e-19))
    beta = LD(np.float64(1.0) / (np.float64(kB) * np.float64(T)))   # and beta likewise (pimc.py:664)
    taus = [beta / P, (beta + LD(delta_beta)) / P, (beta - LD(delta_beta)) / P]
    taus = [LD(np.float64(t)) for t in taus]                # tau, tau+- are float64 numbers in the reference
    E, w, L, Q = (np.asarray(vib[k], dtype=LD) for k in ("E", "w", "L", "Q"))
    Er, wr, Lr = (np.asarray(rho[k], dtype=LD) for k in ("E", "w", "L"))
    Rl = np.asarray(R, dtype=LD)                            # (X, N, P)
    idx = np.arange(A)
    d = -L[:, idx, idx].T / w[None, :]                      # (A, N)
    Et = E[idx, idx] - (L[:, idx, idx].T ** 2 / w[None, :]).sum(1) / 2
    dr = -Lr.T / wr[None, :]                                # (Ar, N)
    Etr = Er - (Lr.T ** 2 / wr[None, :]).sum(1) / 2
    n_rho = min(A, Ar) if rho_trunc else Ar

    def log_o(t, shift, Etil, ww, na):
        q = Rl[:, None, :, :] - shift[None, :na, :, None]  # (X, a, N, P)
        qp = np.roll(q, -1, axis=-1)
        x = t * ww
        coth, csch = np.cosh(x) / np.sinh(x), 1 / np.sinh(x)
        s = (coth[None, None, :, None] * (q * q + qp * qp) - 2 * csch[None, None, :, None] * q * qp).sum(2)   # (X, a, P)
        lpref = -t * Etil[:na] + np.log(csch).sum() / 2
        return (lpref[None, :, None] - s / 2).transpose(0, 2, 1)    # (X, P, a)

    lv = [log_o(t, d, Et, w, A) for t in taus]
    lr = log_o(taus[0], dr, Etr, wr, n_rho)
    logS = np.maximum(lv[0].max(-1), lr.max(-1))            # (X, P)
    rho_val = np.exp((lr - logS[..., None]).sum(1)).sum(-1)
    Eoff = E.copy(); Eoff[idx, idx] = 0
    Loff = L.copy(); Loff[:, idx, idx] = 0
    V = Eoff[None, None] + np.einsum("nij,xnp->xpij", Loff, Rl) + np.einsum("nmij,xnp,xmp->xpij", Q, Rl, Rl) / 2
    M = _expm_ld(-taus[0] * V)                              # (X, P, A, A)
    out = [rho_val]
    for v in range(3):
        O = np.exp(lv[v] - logS[..., None])                # (X, P, A)
        Tm = np.broadcast_to(np.eye(A, dtype=LD), (Rl.shape[0], A, A)).copy()
        for p in range(P):
            Tm = np.matmul(Tm, M[:, p]) * O[:, p][:, None, :]
        out.append(np.trace(Tm, axis1=1, axis2=2))
    return np.stack(out)
