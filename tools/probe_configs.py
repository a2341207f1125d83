"""Throughput of the fused call on the BASELINE.json configurations other than the bench one (GPU box).

    python tools/probe_configs.py
"""
import sys
from os.path import abspath, dirname

sys.path.insert(0, dirname(dirname(abspath(__file__))))

import numpy as np
import torch

from pibronic_b200 import _cabi, constants, synthetic
from pibronic_b200.model_io import VMK


def run(name, model, rho, P, X, T=300.0, flags=_cabi.FLAG_PM):
    plan = _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model.get(VMK.G2), rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                      P, constants.beta(T), constants.delta_beta, flags=flags, device=0)
    out = torch.empty((4, X), dtype=torch.float64, device="cuda")
    plan.sample_eval(1, 0, X, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for k in range(3):
        e0.record()
        plan.sample_eval(2 + k, 0, X, out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    o = out.cpu().numpy()
    r = o[1] / o[0]
    print(f"{name:34s} fast={plan.is_fast!s:5s} X={X:8d} P={P:4d} {best:9.3f} ms  {X * P / best * 1e3:.3e} samples*beads/s   "
          f"<g/rho> = {r.mean():.5g} +- {r.std() / np.sqrt(X):.2g}")
    plan.close()


def main():
    c1 = synthetic.coupled_model(2, 2, (0.01, 0.02), (0.0, 0.1), seed=1, linear=0.05, quadratic=0.0, mixing=0.0)
    run("c1-like  A=2 N=2 P=12 X=1e4", c1, synthetic.diagonal_of(c1), 12, 10_000)
    run("c1-like  A=2 N=2 P=12 X=1e7", c1, synthetic.diagonal_of(c1), 12, 10_000_000)
    c2 = synthetic.model_c2()
    run("c2       A=4 N=6 P=64 X=1e6", c2, synthetic.diagonal_of(c2), 64, 1_000_000)
    run("c2 non-PM (block_compute)", c2, synthetic.diagonal_of(c2), 64, 1_000_000, flags=0)
    c3 = synthetic.coupled_model(2, 2, (0.02, 0.04), (0.1, 0.2), seed=3, linear=0.05, quadratic=0.0)
    run("c3-like  A=2 N=2 P=128 X=1e6", c3, synthetic.diagonal_of(c3), 128, 1_000_000)
    free = synthetic.coupled_model(4, 6, (0.14, 0.45), (10.3, 10.9), mixing=0.0, quadratic=0.0)
    run("c5       c2 with another rho", c2, synthetic.diagonal_of(free), 64, 1_000_000)
    run("c2 consistent estimator", c2, synthetic.diagonal_of(c2), 64, 1_000_000, flags=_cabi.FLAG_PM | _cabi.FLAG_M_TAU_PM)
    c4 = synthetic.model_c4()
    run("c4       A=12 N=24 P=256 X=16384", c4, synthetic.diagonal_of(c4), 256, 16384)


if __name__ == "__main__":
    main()
