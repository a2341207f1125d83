"""CPU restatement of the DEVICE sampler (TEST INFRASTRUCTURE ONLY -- see oracle/pimc_oracle.py header).

The reference draws with numpy's global MT19937 in ring normal-mode co-ordinates
(/root/reference/pibronic/pimc/pimc.py:326-334, 386-406, 613-631); that sampler is restated in
``pimc_oracle.draw_block``.  The CUDA path draws the *same distribution* differently
(DESIGN.md "sampler"): Philox4x32-10 counter streams, FP64 Box-Muller, and a sequential
cyclic-tridiagonal-Cholesky recurrence along the ring.  This module restates THAT algorithm in
numpy so the tests can check the device sampler value-for-value (up to libm rounding) and, on the
CPU, check the distribution exactly (covariance algebra) and statistically at scale.

Follows pibronic_b200/csrc/pbx_device.cuh (philox4x32_10, u01_*, normal_pair, pick_source) and
pbx_tables.hpp (ring_recurrence).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)
STREAM_NORMALS, STREAM_SOURCE = 0, 1


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """vectorised Philox4x32-10; all inputs uint32 arrays (broadcastable); returns 4 uint32 arrays"""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _bits53(hi, lo):
    return ((hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)) >> np.uint64(11)


def u01_open_low(hi, lo):
    """(0, 1]"""
    return (_bits53(hi, lo).astype(np.float64) + 1.0) * 2.0 ** -53


def u01_half_open(hi, lo):
    """[0, 1)"""
    return _bits53(hi, lo).astype(np.float64) * 2.0 ** -53


def normal_pairs(r0, r1, r2, r3):
    u1 = u01_open_low(r0, r1)
    u2 = u01_half_open(r2, r3)
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * np.pi * u2
    return rad * np.cos(ang), rad * np.sin(ang)


def sample_sources(wcum, seed, first, n):
    idx = np.arange(first, first + n, dtype=np.uint64)
    lo, hi = (idx & MASK32).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32)
    r = philox4x32_10(lo, hi, np.uint32(0), np.uint32(STREAM_SOURCE), seed & 0xFFFFFFFF, seed >> 32)
    u = u01_half_open(r[0], r[1])
    src = np.zeros(n, dtype=np.int32)
    for a in range(len(wcum) - 1):
        src += (u >= wcum[a]).astype(np.int32)
    return src


def standard_normals(seed, first, n, N, P):
    """z[x, j, n]: the N(0,1) variate used for bead j (generation order), mode n of global sample first+x"""
    idx = np.arange(first, first + n, dtype=np.uint64)
    lo, hi = (idx & MASK32).astype(np.uint32)[:, None], (idx >> np.uint64(32)).astype(np.uint32)[:, None]
    half = (N + 1) // 2
    draw = np.arange(P * half, dtype=np.uint32)[None, :]
    r = philox4x32_10(lo, hi, draw, np.uint32(STREAM_NORMALS), seed & 0xFFFFFFFF, seed >> 32)
    z0, z1 = normal_pairs(*r)                       # (n, P*half)
    z = np.stack([z0, z1], axis=-1).reshape(n, P, 2 * half)
    return z[:, :, :N]


def ring_recurrence_dense(alpha, s, P):
    """(a, b, e) of y_j = a_j z_j + b_j y_{j-1} + e_j y_0 from a DENSE Cholesky of alpha*I - s*C
    (independent of the O(P) algorithm in pbx_tables.hpp)"""
    i = np.arange(P)
    lam = alpha * np.eye(P)
    lam[i, (i + 1) % P] -= s
    lam[i, (i - 1) % P] -= s
    L = np.linalg.cholesky(lam)
    a, b, e = np.zeros(P), np.zeros(P), np.zeros(P)
    for j in range(P):
        row = P - 1 - j
        a[j] = 1.0 / L[row, row]
        if j >= 1:
            b[j] = -L[row + 1, row] / L[row, row]
        if j >= 2:
            e[j] = -L[P - 1, row] / L[row, row]
    return a, b, e


def sample_coords(samp, wcum, d_rho, seed, first, n):
    """R[x, n, p] and the mixture component of each sample, exactly as pbx_sample_coords_kernel does.
    samp: (P, N, 3) recurrence table; wcum: (Ar,) cumulative weights; d_rho: (Ar, N)."""
    P, N, _ = samp.shape
    src = sample_sources(wcum, seed, first, n)
    z = standard_normals(seed, first, n, N, P)
    y = np.zeros((n, P, N))
    y[:, 0] = samp[0, :, 0] * z[:, 0]
    for j in range(1, P):
        y[:, j] = samp[j, :, 0] * z[:, j] + samp[j, :, 1] * y[:, j - 1] + samp[j, :, 2] * y[:, 0]
    R = y.transpose(0, 2, 1) + d_rho[src][:, :, None]
    return np.ascontiguousarray(R), src
