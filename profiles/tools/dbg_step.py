import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from tests import synth
from mfgp_coverage_b200 import simulator as sim, _coverage as cv
import bench
w = bench.make_workload("c4", 1, 0, "strong")
model = sim.init_MFGP(synth.MF_HYP, np.column_stack((w["X_L"], w["y_L"])))
model.updt_info(w["X_L"], w["y_L"], w["X_H"], w["y_H"])
eng = model.engine
state = sim._Sim(np.column_stack((w["xy"], w["f"])))
bbox = np.array([0., 1., 0., 1.])
for it in range(4):
    eng.refactor(check=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lv = cv.HybridVoronoi(w["pos"], bbox, reuse=state._clip[0]); pv = cv.HybridVoronoi(w["cen"], bbox, reuse=state._clip[1]); state._clip=[lv,pv]
    torch.cuda.synchronize(); t1 = time.perf_counter()
    eng.lazy_check = eng.defer_fit = True
    model.predict_device(state.grid.xy, state.mu, state.var, grid=state.grid)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    res = state.grid.assign_reduce(pv, lv, w=state.mu, var=state.var, amax_k0=0.067, amax_rel=1e-10)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    out = state.grid.finish(res, pv, lv, bbox, info=eng.info, with_ties=True)
    t4 = time.perf_counter()
    print("clip %.3f post %.3f cov %.3f finish %.3f ms ties=%d" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, out[4]))
for it in range(3):
    eng.refactor(check=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = state.step(model, w["pos"], w["cen"])
    torch.cuda.synchronize(); print("step %.3f ms" % ((time.perf_counter()-t0)*1e3), r[0], type(r[4]).__name__, r[4]._qhull is not None)
