"""Time the Choi greedy planner (compute_sample_points -> choi_greedy: one pass over the cached V per pick) on an n x n
grid with N training samples; reports per-pick time and the achieved HBM bandwidth of the append kernel
(algorithmic bytes per pick = 8 * rows * G).  usage: prof_choi.py [n=256] [N=1024] [frac=0.5]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import synth
from mfgp_coverage_b200 import simulator as sim

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
xy = synth.grid(n)
G = xy.shape[0]
f = synth.truth_function(xy)
X_L, y_L, X_H, y_H = synth.training_set(xy, f, N)
m = sim.init_MFGP(synth.MF_HYP, np.column_stack((X_L, y_L)))
m.updt_info(X_L, y_L, X_H, y_H)
mu, var = m.predict(xy)
thr = frac * var.max()
sim.compute_sample_points(m, xy, thr)          # warm-up (uploads the grid, sizes the caches)
sim.compute_sample_points(m, xy, 0.8 * thr)    # ... including the allocator blocks of the grown V cache
torch.cuda.synchronize()


def timed(th):
    t0 = time.perf_counter()
    _, idx = sim.compute_sample_points(m, xy, th, return_indices=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, idx


# two thresholds: the difference isolates the per-pick cost from the initial posterior + V-cache fill
t1, idx1 = timed(thr)
t2, idx2 = timed(0.8 * thr)
k1, k2 = len(idx1), len(idx2)
peak = 6542.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
per = (t2 - t1) / max(k2 - k1, 1)
rows = N + (k1 + k2 - 1) / 2.0
gbs = 8.0 * rows * G / per * 1e-9
print(f"choi planner n={n} G={G} N={N}: {k1} picks in {t1*1e3:.1f} ms, {k2} picks in {t2*1e3:.1f} ms (each incl. the initial "
      f"posterior + V cache) -> {per*1e6:.1f} us per additional pick = {gbs:.0f} GB/s algorithmic (8 B x {rows:.0f} rows x G) "
      f"= {gbs/peak:.3f} of measured HBM copy ({peak} GB/s); first picks {idx1[:5].tolist()}")
