"""CPU check of the table-driven device math (pibronic_b200/csrc/pbx_device.cuh: neg2_log_pos, sqrt_pos, sincos_2pi_bits):
the same tables (parsed from the generated header) and the same formulas restated in numpy, against 80-bit references.
The GPU test of the real functions is tests/test_gpu_parity.py::test_device_math; this one pins the tables and the
truncation orders without a device (an FMA is emulated in long double, so the last bit may differ from the device)."""
import re
from os.path import dirname, join

import numpy as np
import pytest

LD = np.longdouble
pytestmark = pytest.mark.skipif(np.finfo(LD).eps > 2e-19, reason="needs an 80-bit long double")


def _tables():
    with open(join(dirname(dirname(__file__)), "pibronic_b200", "csrc", "pbx_math_tables.h")) as fh:
        text = fh.read()

    def table(name):
        body = text[text.index(name):]
        body = re.sub(r"//.*", "", body[body.index("{") + 1:body.index("};")])
        return np.array([float.fromhex(v) for v in re.findall(r"-?0x[0-9a-f.]+p[+-]\d+", body)])
    return table("kLogTab[").reshape(-1, 2), table("kLogExpTab["), table("kSinCosTab[").reshape(-1, 2)


def _fma(a, b, c):
    return (a.astype(LD) * b + c).astype(np.float64)


def neg2_log(u, log_tab, exp_tab):
    bits = u.view(np.int64)
    hi, lo = bits >> 32, bits & 0xffffffff
    frac = hi & 0xfffff
    big = frac >= 0x80000
    j = 1023 - (hi >> 20) - big
    idx = np.where(big, ((frac + 0x800) >> 12) - 128, ((frac + 0x400) >> 11) + 128)
    m = (((frac | np.where(big, 0x3fe00000, 0x3ff00000)) << 32) | lo).view(np.float64)
    base = exp_tab[np.clip(j, 0, 63)] + log_tab[idx, 1]
    t = _fma(m, log_tab[idx, 0], -1.0)
    q = _fma(t, np.float64(-2.0 / 5), 0.5)
    for c in (-2.0 / 3, 1.0, -2.0):
        q = _fma(q, t, c)
    return _fma(t, q, base), idx, t


def test_tables_have_the_documented_shape_and_exact_centre():
    log_tab, exp_tab, sincos_tab = _tables()
    assert log_tab.shape == (385, 2) and exp_tab.shape == (64,) and sincos_tab.shape == (256, 2)
    assert log_tab[128, 0] == 1.0 and log_tab[128, 1] == 0.0 and exp_tab[0] == 0.0      # u -> 1 keeps relative accuracy
    assert np.allclose(sincos_tab[:, 0] ** 2 + sincos_tab[:, 1] ** 2, 1.0, rtol=0, atol=3e-16)
    assert sincos_tab[0, 0] == 0.0 and sincos_tab[0, 1] == 1.0 and sincos_tab[64, 0] == 1.0 and sincos_tab[128, 1] == -1.0


def test_neg2_log_restated():
    log_tab, exp_tab, _ = _tables()
    rng = np.random.default_rng(0)
    u = (rng.integers(0, 1 << 53, 200_000, dtype=np.int64) + 1).astype(np.float64) * 2.0 ** -53
    u = np.concatenate([u, 2.0 ** -rng.integers(1, 54, 4096).astype(np.float64),
                        [1.0, 2.0 ** -53, 0.5, 0.75, np.nextafter(0.75, 0), 0.375, np.nextafter(0.5, 0), 1 - 2.0 ** -53,
                         1 - 2.0 ** -30, 1 - 1.0 / 1024, 1 - 1.0 / 1023, 1 + 0.0]])
    got, idx, t = neg2_log(u, log_tab, exp_tab)
    want = (-2 * np.log(u.astype(LD)))
    assert idx.min() >= 0 and idx.max() <= 384 and np.abs(t).max() <= 1.0 / 768 * (1 + 1e-12)
    rel = np.abs(got - want) / np.maximum(np.abs(want), LD(1e-300))
    assert float(rel.max()) < 7e-16        # 6e-16 at the edge of the cell around 1 (series cut at t^5), 3e-16 elsewhere


def test_sqrt_series_restated():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.random(200_000) * 80, rng.random(4096) * 1e-6, [75.0, 1.0, 4.0, 0.0, 2.2e-16]])
    seed = (1.0 / np.sqrt(np.maximum(x.astype(np.float32), np.float32(1e-30)).astype(np.float64)))
    r = (seed * (1 + rng.uniform(-1, 1, x.shape) * 2.0 ** -22.5)).astype(np.float32).astype(np.float64)   # MUFU: 22 bits
    e = _fma(-x, r * r, 1.0)
    w = x * r
    got = _fma(w, (e * 0.375 + 0.5) * e, w)
    want = np.sqrt(x.astype(LD))
    assert got[-2] == 0.0
    assert float((np.abs(got - want) / np.maximum(want, LD(1e-300))).max()) < 4e-16


def test_sincos_restated():
    _, _, sincos_tab = _tables()
    rng = np.random.default_rng(2)
    b = np.concatenate([rng.integers(0, 1 << 53, 200_000, dtype=np.int64),
                        np.array([0, 1 << 51, 1 << 52, 3 << 51, 1 << 50, (1 << 53) - 1, 1 << 44, (1 << 44) - 1, (255 << 45) + (1 << 44)])])
    t = b + (1 << 44)
    idx = (t >> 45) & 255
    f = (t & ((1 << 45) - 1)) - (1 << 44)
    S, C = sincos_tab[idx, 0], sincos_tab[idx, 1]
    h = f.astype(np.float64) * (2 * np.pi * 2.0 ** -53)
    assert np.abs(h).max() <= np.pi / 256 * (1 + 1e-12)
    h2 = h * h
    sh = _fma(h * h2, _fma(h2, np.float64(1.0 / 120), -1.0 / 6), h)
    cm1 = h2 * _fma(h2, _fma(h2, np.float64(-1.0 / 720), 1.0 / 24), -0.5)
    sn = _fma(S, cm1, _fma(C, sh, S))
    cs = _fma(C, cm1, _fma(-S, sh, C))
    ang = b.astype(LD) * (2 * LD(np.pi) + LD("1.2246467991473531772e-16") * 2) / LD(2.0) ** 53    # 2 pi to 35 digits
    assert float(np.abs(sn - np.sin(ang)).max()) < 3e-16 and float(np.abs(cs - np.cos(ang)).max()) < 3e-16
