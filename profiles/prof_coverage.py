"""Time the fused coverage kernels (cov_assign_reduce: both partitions, per-cell sums, per-cell arg-max + finalize) on an
n x n grid with A agents.  The launches of NB independent copies of the grid (NB x 40 B x G > L2, so every launch
streams from HBM) are captured in one CUDA graph and replayed: no host gaps inside the timed region.  Reports the
achieved algorithmic bandwidth (40 B per grid point) against MEASURED_PEAKS.json's HBM copy figure.
usage: prof_coverage.py [n=1024] [A=64] [ongrid=0|1] [graph=1|0]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import synth
from mfgp_coverage_b200 import simulator as sim
from mfgp_coverage_b200._coverage import CoverageGrid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
A = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ongrid = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
use_graph = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
xy = synth.grid(n)
G = xy.shape[0]
f = synth.truth_function(xy)
rng = np.random.default_rng(3)
NB = max(2, int(np.ceil(400e6 / (40.0 * G))))          # copies: 400 MB total > 126 MB L2
NB = min(NB, 64)
grids = [CoverageGrid(xy, f) for _ in range(NB)]
mus = [torch.from_numpy(f + 0.1 * rng.standard_normal(G)).cuda() for _ in range(NB)]
vrs = [torch.from_numpy(rng.random(G)).cuda() for _ in range(NB)]
pos, cen = synth.agents(A, 7), synth.agents(A, 8)
if ongrid:      # agents on grid points: grid points sit exactly on bisectors (tie path)
    pos = xy[rng.choice(G, A, replace=False)]
    cen = xy[rng.choice(G, A, replace=False)]
bbox = np.array([0.0, 1.0, 0.0, 1.0])
lv, pv = sim.voronoi_bounded(cen, bbox), sim.voronoi_bounded(pos, bbox)
dl, dp = grids[0].upload(lv), grids[0].upload(pv)
outs = [g.assign_reduce(dl, dp, w=m, var=v) for g, m, v in zip(grids, mus, vrs)]     # warm-up, allocates outputs
torch.cuda.synchronize()


def sweep():
    for g, m, v, o in zip(grids, mus, vrs, outs):
        g.assign_reduce(dl, dp, w=m, var=v, out=o)


if use_graph:
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        sweep()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            sweep()
    run = graph.replay
else:
    run = sweep
for _ in range(3):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / NB)
ms = float(np.median(ts))
peak = 6542.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
gbs = 40.0 * G / (ms * 1e-3) * 1e-9
res = outs[0]
print(f"coverage n={n} G={G} A={A} ongrid={int(ongrid)} graph={int(use_graph)} copies={NB}: assign+finalize per grid median "
      f"{ms*1e3:.2f} us (min {min(ts)*1e3:.2f}) -> {gbs:.0f} GB/s algorithmic (40 B/pt) = {gbs/peak:.3f} of measured HBM copy "
      f"({peak} GB/s); cent[0]={res['cent'][0].tolist()} amax_idx[:4]={res['amax_idx'][:4].tolist()}")
