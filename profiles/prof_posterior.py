"""Short driver for ncu: one fit + a few posterior launches at N (default 4096) on a 148x64 tensor grid = 2 CTAs per SM.
usage: prof_posterior.py [N] [mode: grid|general]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import synth
from mfgp_coverage_b200 import simulator as sim
from mfgp_coverage_b200._coverage import CoverageGrid

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = sys.argv[2] if len(sys.argv) > 2 else "grid"
ux, uy = np.linspace(0, 1, 148), np.linspace(0, 1, 64)
xy = np.stack(np.meshgrid(ux, uy, indexing="ij"), axis=-1).reshape(-1, 2)
G = xy.shape[0]
X_L, y_L, X_H, y_H = synth.training_set(synth.grid(256), synth.truth_function(synth.grid(256)), N)
m = sim.init_MFGP(synth.MF_HYP, np.column_stack((X_L, y_L)))
m.updt_info(X_L, y_L, X_H, y_H)
m.use_separable = mode == "grid"
g = CoverageGrid(xy)
assert g.axes is not None
for _ in range(3):
    mu, var = m.predict_device(g.xy, grid=g)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); mu, var = m.predict_device(g.xy, grid=g); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"mode={mode} N={N} G={G} posterior {ms:.3f} ms  {G * N * N / ms * 1e-9:.2f} TFLOP/s  var[0]={float(var[0]):.6e} mu[5]={float(mu[5]):.6e}")
