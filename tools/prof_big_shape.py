"""Times the fused large-A kernel on a synthetic (A, N, P) model (GPU box): python tools/prof_big_shape.py A N P samples"""
import sys
from os.path import abspath, dirname

sys.path.insert(0, dirname(dirname(abspath(__file__))))

import torch

from bench import algorithmic_flops_per_sample
from pibronic_b200 import _cabi, constants, synthetic
from pibronic_b200.model_io import VMK

A, N, P, X = (int(float(v)) for v in sys.argv[1:5])
model = synthetic.coupled_model(A, N, (0.1, 0.39), (14.0, 14.8))
rho = synthetic.diagonal_of(model)
plan = _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model[VMK.G2], rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                  P, constants.beta(300.0), constants.delta_beta, flags=_cabi.FLAG_PM | _cabi.FLAG_PREFER_DMMA, device=0)
out = torch.empty((4, X), dtype=torch.float64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e30
for k in range(4):
    e0.record()
    plan.sample_eval(100 + k, 0, X, out)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
tf = algorithmic_flops_per_sample(A, N, P, A) * X / (best * 1e-3) / 1e12
print(f"A={A} N={N} P={P} path={plan.kernel_path} X={X}: {best:.2f} ms, {X * P / best * 1e3:.3e} samples*beads/s, {tf:.2f} TFLOP/s algorithmic")
