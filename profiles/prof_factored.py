"""Short driver for ncu: fit + the Chebyshev-factored posterior on an n x n tensor grid with N training samples.
usage: prof_factored.py [n=1024] [N=4096] [reps=3]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import synth
from mfgp_coverage_b200 import simulator as sim
from mfgp_coverage_b200._coverage import CoverageGrid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
xy = synth.grid(n)
base = synth.grid(min(n, 256))
X_L, y_L, X_H, y_H = synth.training_set(base, synth.truth_function(base), N)
m = sim.init_MFGP(synth.MF_HYP, np.column_stack((X_L, y_L)))
m.updt_info(X_L, y_L, X_H, y_H)
g = CoverageGrid(xy)
mu = torch.empty(g.G, dtype=torch.float64, device=g.device)
var = torch.empty(g.G, dtype=torch.float64, device=g.device)
for _ in range(2):
    m.predict_device(g.xy, mu, var, grid=g)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.predict_device(g.xy, mu, var, grid=g); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
plan = m.engine._fplan[1] if m.engine._fplan else None
print(f"n={n} N={N} factored posterior {min(ts):.3f} ms (plan {plan and {k: plan[k] for k in ('rxL','ryL','rxH','ryH','chunk')}}) "
      f"var[0]={float(var[0]):.6e} mu[5]={float(mu[5]):.6e}")
