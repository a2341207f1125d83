"""Profiling driver (GPU box): estimator on device-resident coordinates (no sampling) for the c2 workload."""
import sys
from os.path import abspath, dirname

sys.path.insert(0, dirname(dirname(abspath(__file__))))

import torch

from pibronic_b200 import _cabi, constants, synthetic
from pibronic_b200.model_io import VMK

X = int(float(sys.argv[1])) if len(sys.argv) > 1 else 262144
model = synthetic.model_c2()
rho = synthetic.diagonal_of(model)
plan = _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model[VMK.G2], rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                  64, constants.beta(300.0), constants.delta_beta, flags=_cabi.FLAG_PM, device=0)
R = torch.empty((X, 6, 64), dtype=torch.float64, device="cuda")
out = torch.empty((4, X), dtype=torch.float64, device="cuda")
out2 = torch.empty((4, X), dtype=torch.float64, device="cuda")
plan.sample_coords(5, 0, X, R)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(3):
    e0.record()
    plan.eval_coords(R, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"eval_coords X={X}: {ms:.3f} ms  {X * 64 / ms * 1e3:.3e} samples*beads/s, coordinate bytes {R.numel() * 8 / ms * 1e-6:.1f} GB/s")
plan.sample_eval(5, 0, X, out2)
torch.cuda.synchronize()
print("fused == coords path:", float((out2 / out - 1).abs().max()))
