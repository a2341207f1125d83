"""Times the fused launch on the 2-surface configurations (c1, c3) on the GPU box: best of 10 launches, CUDA events.

    [PBX_LIB=pibronic_b200/_pbx_<variant>.so] python tools/time_small.py
"""
import sys
from os.path import abspath, dirname

sys.path.insert(0, dirname(dirname(abspath(__file__))))

import numpy as np
import torch

from bench import algorithmic_flops_per_sample
from pibronic_b200 import _cabi, constants, synthetic
from pibronic_b200.model_io import VMK


def run(name, model, rho, P, X, flags=_cabi.FLAG_PM):
    plan = _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model.get(VMK.G2), rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                      P, constants.beta(300.0), constants.delta_beta, flags=flags, device=0)
    out = torch.empty((4, X), dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for k in range(13):
        e0.record()
        plan.sample_eval(100 + k, 0, X, out)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    best = min(ms[3:])
    tf = algorithmic_flops_per_sample(plan.A, plan.N, P, plan.Ar) * X / (best * 1e-3) / 1e12
    r = (out[1] / out[0]).mean().item()
    print(f"{name:34s} X={X:8d} {best:8.4f} ms  {X * P / best * 1e3:.4e} samples*beads/s  {tf:6.2f} TFLOP/s  <g/rho> = {r:.6f}")
    plan.close()


c1 = synthetic.coupled_model(2, 2, (0.01, 0.02), (0.0, 0.1), seed=1, linear=0.05, quadratic=0.0, mixing=0.0)
c3 = synthetic.coupled_model(2, 2, (0.02, 0.04), (0.1, 0.2), seed=3, linear=0.05, quadratic=0.0)
run("c1 A=2 N=2 P=12", c1, synthetic.diagonal_of(c1), 12, 10_000)
run("c1 A=2 N=2 P=12", c1, synthetic.diagonal_of(c1), 12, 10_000_000)
run("c3 A=2 N=2 P=128", c3, synthetic.diagonal_of(c3), 128, 100_000)
run("c3 A=2 N=2 P=128", c3, synthetic.diagonal_of(c3), 128, 1_000_000)
run("c3 A=2 N=2 P=128 non-PM", c3, synthetic.diagonal_of(c3), 128, 1_000_000, flags=0)
run("c3 one-role kernel", c3, synthetic.diagonal_of(c3), 128, 1_000_000, flags=_cabi.FLAG_PM | _cabi.FLAG_NO_WARPSPEC)
run("c3 one-role kernel", c3, synthetic.diagonal_of(c3), 128, 100_000, flags=_cabi.FLAG_PM | _cabi.FLAG_NO_WARPSPEC)
run("c1 one-role kernel", c1, synthetic.diagonal_of(c1), 12, 10_000, flags=_cabi.FLAG_PM | _cabi.FLAG_NO_WARPSPEC)
m323 = synthetic.coupled_model(3, 3, (0.02, 0.04), (0.1, 0.2), seed=5, linear=0.05, quadratic=0.02)
run("A=3 N=3 P=64", m323, synthetic.diagonal_of(m323), 64, 1_000_000)
c2 = synthetic.model_c2()
run("c2 A=4 N=6 P=64", c2, synthetic.diagonal_of(c2), 64, 1_000_000)
