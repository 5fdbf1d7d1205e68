import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from tests import synth
import bench
from mfgp_coverage_b200 import simulator as sim, _coverage as cv
from mfgp_coverage_b200._engine import TensorAxes
w = bench.make_workload("c4")
dev = torch.device("cuda", 0)
model = sim.init_MFGP(synth.MF_HYP, np.column_stack((w["X_L"], w["y_L"])))
model.updt_info(w["X_L"], w["y_L"], w["X_H"], w["y_H"])
eng = model.engine
eng.defer_fit = True
axes = TensorAxes(w["ux"], w["uy"], dev)
grid = cv.CoverageGrid(w["xy"], w["f"], base_index=0, axes=axes)
mu = torch.empty(grid.G, dtype=torch.float64, device=dev); var = torch.empty_like(mu)
eng.incremental = True
eng.refactor(check=False)
eng.posterior(grid.xy, mu, var, axes=axes, g_lo=0)
base = synth.grid(1024)
rng = np.random.default_rng(5)
def ev(): return torch.cuda.Event(enable_timing=True)
for s in range(5):
    idx = rng.choice(base.shape[0], 64, replace=False)
    x_new = base[idx]; y_new = rng.normal(0, 0.1, (64, 1))
    torch.cuda.synchronize()
    e = [ev() for _ in range(3)]
    t0 = time.perf_counter()
    e[0].record(); eng.append_hifi(x_new, y_new, check=False); e[1].record()
    t1 = time.perf_counter()
    eng.posterior(grid.xy, mu, var, axes=axes, g_lo=0); e[2].record()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"step {s}: append gpu {e[0].elapsed_time(e[1]):.2f} ms host {1e3*(t1-t0):.2f} | posterior gpu {e[1].elapsed_time(e[2]):.2f} ms host {1e3*(t2-t1):.2f} | N={eng.N} npad={eng.npad} cap={eng.cap} plan={eng._fplan[1] and eng._fplan[1]['chunk']}")
