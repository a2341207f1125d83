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


def run(name, model, rho, P, X, T=300.0, flags=_cabi.FLAG_PM, jit=False):
    plan = _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model.get(VMK.G2), rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                      P, constants.beta(T), constants.delta_beta, flags=flags, device=0, jit=jit)
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
    sys.path.insert(0, dirname(dirname(abspath(__file__))))
    from bench import algorithmic_flops_per_sample
    tf = algorithmic_flops_per_sample(plan.A, plan.N, P, plan.Ar) * X / (best * 1e-3) / 1e12
    print(f"{name:38s} path={plan.kernel_path} X={X:8d} P={P:4d} {best:9.3f} ms  {X * P / best * 1e3:.3e} samples*beads/s  "
          f"{tf:6.2f} TFLOP/s algorithmic   <g/rho> = {r.mean():.5g} +- {r.std() / np.sqrt(X):.2g}")
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
    if _cabi.has_feature(_cabi.FEATURE_MTAU):
        run("c2 consistent estimator", c2, synthetic.diagonal_of(c2), 64, 1_000_000, flags=_cabi.FLAG_PM | _cabi.FLAG_M_TAU_PM)
    c4 = synthetic.model_c4()
    run("c4       A=12 N=24 P=256 X=18944", c4, synthetic.diagonal_of(c4), 256, 18944)
    run("c4       blocked kernels (round 1)", c4, synthetic.diagonal_of(c4), 256, 18944, flags=_cabi.FLAG_PM | _cabi.FLAG_NO_FUSED_DMMA)
    # shapes that are not in csrc/shapes.def: the shape of the reference's largest example model (model_7x12.json) and a (5,3,5) one
    m712 = synthetic.coupled_model(7, 12, (0.02, 0.3), (0.5, 1.2), seed=23, quadratic=0.06)
    run("7x12     fused tensor-core kernel", m712, synthetic.diagonal_of(m712), 64, 200_000)
    run("7x12     blocked kernels (round 1)", m712, synthetic.diagonal_of(m712), 64, 200_000, flags=_cabi.FLAG_PM | _cabi.FLAG_NO_FUSED_DMMA)
    m535 = synthetic.coupled_model(5, 3, (0.05, 0.2), (1.0, 1.4), seed=535, quadratic=0.08)
    run("5x3      fused tensor-core kernel", m535, synthetic.diagonal_of(m535), 64, 1_000_000)
    run("5x3      blocked kernels (round 1)", m535, synthetic.diagonal_of(m535), 64, 1_000_000, flags=_cabi.FLAG_PM | _cabi.FLAG_NO_FUSED_DMMA)
    run("5x3      run-time compiled register kernel", m535, synthetic.diagonal_of(m535), 64, 1_000_000, jit=True)


if __name__ == "__main__":
    main()
