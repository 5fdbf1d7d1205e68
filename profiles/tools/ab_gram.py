"""A/B of the two routes of the factored posterior's per-column Gram stage at c4 (MFGP_GRAM=direct|m), whole posterior call.
usage: ab_gram.py [c4] [world=1]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import synth
from mfgp_coverage_b200 import simulator as sim
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # > 1: the column slice of rank world // 2 of a strong-scaling run
w = bench.make_workload(name, world, world // 2, "strong")
model = sim.init_MFGP(synth.MF_HYP, np.column_stack((w["X_L"], w["y_L"])))
model.updt_info(w["X_L"], w["y_L"], w["X_H"], w["y_H"])
eng = model.engine
state = sim._Sim(np.column_stack((w["xy"], w["f"])))
res = {}
for route in ("direct", "m", "direct", "m"):
    os.environ["MFGP_GRAM"] = route
    ts = []
    for it in range(6):
        eng.refactor(check=False)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); model.predict_device(state.grid.xy, state.mu, state.var, grid=state.grid); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    res[route] = (state.mu.clone(), state.var.clone())
    print(route, "posterior call ms:", ["%.3f" % t for t in ts])
print("max |dvar| / k0 =", float((res["m"][1] - res["direct"][1]).abs().max()) / 0.0672, " max |dmu| =", float((res["m"][0] - res["direct"][0]).abs().max()))
