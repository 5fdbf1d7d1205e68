"""Host-side timeline of bench.py's e2e step (reference-style calls with host arrays) at c4: wall time per call, with a
device synchronisation after each so the numbers add up.  usage: e2e_breakdown.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from mfgp_coverage_b200 import simulator as sim
w = bench.make_workload("c4")
from tests import synth
model = sim.init_MFGP(synth.MF_HYP, np.column_stack((w["X_L"], w["y_L"])))
model.updt_info(w["X_L"], w["y_L"], w["X_H"], w["y_H"])
xs = np.ascontiguousarray(w["xy"]); truth = np.ascontiguousarray(np.column_stack((w["xy"], w["f"])))
bbox = np.array([0.0, 1.0, 0.0, 1.0])
ex, ey = np.empty((0, 2)), np.empty((0, 1))
def step(tl):
    def tick(name, t0):
        torch.cuda.synchronize(); t1 = time.perf_counter(); tl.setdefault(name, []).append((t1 - t0) * 1e3); return t1
    t = time.perf_counter()
    model.updt_hifi(ex, ey); t = tick("updt_hifi", t)
    mu, var = model.predict(xs); t = tick("predict (fit + posterior + D2H mu, var)", t)
    lv = sim.voronoi_bounded(w["pos"], bbox); t = tick("voronoi_bounded(pos)", t)
    loss = sim.compute_loss(lv, truth); t = tick("compute_loss", t)
    cv_ = sim.voronoi_bounded(w["cen"], bbox); t = tick("voronoi_bounded(cen)", t)
    c = sim.compute_centroids(cv_, xs, mu); t = tick("compute_centroids (H2D mu)", t)
    a = sim.compute_max_var(cv_, truth, var); t = tick("compute_max_var (H2D var)", t)
for i in range(3): step({})
tl = {}
for i in range(10): step(tl)
tot = 0.0
for k, v in tl.items():
    print(f"{k:45s} {np.median(v):7.3f} ms"); tot += np.median(v)
print(f"{'sum':45s} {tot:7.3f} ms")
# raw PCIe copies of 8 MB, pinned
h = torch.empty(1 << 20, dtype=torch.float64, pin_memory=True); d = torch.empty(1 << 20, dtype=torch.float64, device="cuda")
for name, f in (("H2D 8 MB pinned", lambda: d.copy_(h, non_blocking=True)), ("D2H 8 MB pinned", lambda: h.copy_(d, non_blocking=True))):
    ts = []
    for i in range(10):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{name}: {np.median(ts):.3f} ms = {8.39 / np.median(ts):.1f} GB/s")
