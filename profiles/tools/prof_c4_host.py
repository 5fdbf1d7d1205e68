"""Host-side cProfile of the c4 bench step: where the host spends the time during which the GPU has nothing queued."""
import cProfile, pstats, io, sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from tests import synth
from mfgp_coverage_b200 import simulator as sim
from mfgp_coverage_b200 import _coverage as cv
from mfgp_coverage_b200._engine import TensorAxes
w = bench.make_workload("c4")
dev = torch.device("cuda", 0)
lo, hi = w["lo"], w["hi"]
npts = hi - lo
bbox = np.array([0.0, 1.0, 0.0, 1.0])
model = sim.init_MFGP(synth.MF_HYP, np.column_stack((w["X_L"], w["y_L"])))
model.updt_info(w["X_L"], w["y_L"], w["X_H"], w["y_H"])
eng = model.engine
eng.defer_fit = True
axes = TensorAxes(w["ux"], w["uy"], dev)
grid = cv.CoverageGrid(w["xy"], w["f"], base_index=lo, axes=axes)
mu = torch.empty(npts, dtype=torch.float64, device=dev)
var = torch.empty(npts, dtype=torch.float64, device=dev)
marks = {}
def step():
    t0 = time.perf_counter()
    eng.refactor(check=False)
    eng.posterior(grid.xy, mu, var, axes=axes, g_lo=lo)
    t1 = time.perf_counter()
    loss_vor = sim.voronoi_bounded(w["pos"], bbox)
    lloyd_vor = sim.voronoi_bounded(w["cen"], bbox)
    t2 = time.perf_counter()
    res = grid.assign_reduce(lloyd_vor, loss_vor, w=mu, var=var)
    t3 = time.perf_counter()
    loss = cv.loss_from_partials(res["lossp"].cpu().numpy(), loss_vor.areas())
    t4 = time.perf_counter()
    cent = cv.centroids_from_partials(res["cent"].cpu().numpy(), lloyd_vor.areas(), 0.0, 1.0, 0.0, 1.0)
    idx = res["amax_idx"].cpu().numpy()
    t5 = time.perf_counter()
    for k, v in (("launch posterior", t1 - t0), ("voronoi_bounded x2", t2 - t1), ("assign_reduce launch", t3 - t2),
                 ("first .cpu() (waits for the GPU) + loss", t4 - t3), ("2 more .cpu() + centroids", t5 - t4)):
        marks.setdefault(k, []).append(v)
    return loss, cent, idx
for _ in range(3):
    step()
marks.clear()
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    step()
pr.disable()
for k, v in marks.items():
    print(f"{k:45s} {1e3 * np.median(v):8.3f} ms")
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue())
