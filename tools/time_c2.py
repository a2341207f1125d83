"""Times the fused c2 launch (1e6 samples, PM) on the GPU box: best and median of 10 launches, CUDA events."""
import sys
from os.path import abspath, dirname

sys.path.insert(0, dirname(dirname(abspath(__file__))))

import numpy as np
import torch

from pibronic_b200 import _cabi, constants, synthetic
from pibronic_b200.model_io import VMK

X = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
model = synthetic.model_c2()
rho = synthetic.diagonal_of(model)
plan = _cabi.Plan(model[VMK.E], model[VMK.w], model[VMK.G1], model[VMK.G2], rho[VMK.E], rho[VMK.w], rho[VMK.G1],
                  64, constants.beta(300.0), constants.delta_beta, flags=_cabi.FLAG_PM, device=0)
out = torch.empty((4, X), dtype=torch.float64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms = []
for k in range(13):
    e0.record()
    plan.sample_eval(100 + k, 0, X, out)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms = np.array(ms[3:])
r = (out[1] / out[0]).mean().item()
print(f"c2 X={X}: best {ms.min():.3f} ms, median {np.median(ms):.3f} ms, {X * 64 / ms.min() * 1e3:.4e} samples*beads/s, <g/rho> = {r:.5f}")
