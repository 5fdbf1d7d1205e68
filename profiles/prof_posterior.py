"""Short driver for ncu: one fit + a few posterior launches at N=4096 on a grid of 2 waves of CTAs (fast to replay)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import synth
from mfgp_coverage_b200 import simulator as sim

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
G = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 32 * 2
xy = synth.grid(1024)[:G * 8:8].copy()
f = synth.truth_function(synth.grid(1024))[:G * 8:8].copy()
X_L, y_L, X_H, y_H = synth.training_set(synth.grid(256), synth.truth_function(synth.grid(256)), N)
m = sim.init_MFGP(synth.MF_HYP, np.column_stack((X_L, y_L)))
m.updt_info(X_L, y_L, X_H, y_H)
xs = torch.from_numpy(xy).cuda()
for _ in range(3):
    mu, var = m.predict_device(xs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); mu, var = m.predict_device(xs); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"N={N} G={G} posterior {ms:.3f} ms  {G * N * N / ms * 1e-9:.2f} TFLOP/s  var[0]={float(var[0]):.6e}")
