"""GPU probe of the fused large-A kernel (pbx_big.cuh): parity on the golden cases, fused sampler == sampler + estimator,
throughput on the c4 shape against the blocked kernels.

    python tools/probe_big.py [samples]
"""
import sys
from os.path import abspath, dirname, join

ROOT = dirname(dirname(abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, join(ROOT, "tests"))

import numpy as np
import torch

from pibronic_b200 import _cabi, constants, synthetic
from pibronic_b200.model_io import VMK


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


def parity():
    from conftest import CASE_NAMES, GoldenCase
    for name in CASE_NAMES:
        case = GoldenCase(name)
        for pm in (True, False):
            flags = (_cabi.FLAG_PM if pm else 0) | _cabi.QUIRK_RHO_TRUNC | _cabi.FLAG_PREFER_DMMA
            plan = case.plan(flags)
            rows = 4 if pm else 2
            out = plan.eval_coords_host(case.R, out4=np.full((rows, case.R.shape[0]), np.nan))
            err = rel(out, case.expected[:rows])
            n = 300
            fused = plan.sample_eval_host(7, 1234, n, out4=np.full((rows, n), np.nan))
            R = torch.empty((n, plan.N, plan.P), dtype=torch.float64, device="cuda")
            plan.sample_coords(7, 1234, n, R)
            two = plan.eval_coords_host(R.cpu().numpy(), out4=np.full((rows, n), np.nan))
            print(f"{name:16s} pm={pm!s:5s} path={plan.kernel_path} coords-vs-reference {err:.2e}  fused==two-step {np.array_equal(fused, two)} "
                  f"(max rel {rel(fused, two):.1e})", flush=True)
            plan.close()


def c4(X):
    model = synthetic.model_c4()
    rho = synthetic.diagonal_of(model)
    res = {}
    for name, extra in (("fused_dmma", 0), ("blocked", _cabi.FLAG_NO_FUSED_DMMA)):
        plan = _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model[VMK.G2], rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                          256, constants.beta(300.0), constants.delta_beta, flags=_cabi.FLAG_PM | extra, device=0)
        out = torch.empty((4, X), dtype=torch.float64, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for k in range(3):
            e0.record()
            plan.sample_eval(100, 0, X, out)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[name] = out.cpu().numpy()
        print(f"c4 {name:10s} path={plan.kernel_path} X={X}: {best:.2f} ms, {X * 256 / best * 1e3:.3e} samples*beads/s", flush=True)
        plan.close()
    print("c4 fused vs blocked max rel", rel(res["fused_dmma"], res["blocked"]))


if __name__ == "__main__":
    X = int(float(sys.argv[1])) if len(sys.argv) > 1 else 148 * 8 * 4
    parity()
    c4(X)
