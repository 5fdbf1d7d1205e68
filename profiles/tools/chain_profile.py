"""Per-block-column timeline of the tiled Cholesky's chain tasks (MFGP_DF_TRACE=1): time from W_{d-1} published to W_d
published, for every d, in groups of 8 -- shows which phase of the factorisation is chain-bound and which is throughput-bound.
usage: chain_profile.py [N=4096] [R=1344]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["MFGP_DF_TRACE"] = "1"
import numpy as np, torch
from tests import synth
from mfgp_coverage_b200 import simulator as sim, _native as nat
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1344
base = synth.grid(256)
X_L, y_L, X_H, y_H = synth.training_set(base, synth.truth_function(base), N)
m = sim.init_MFGP(synth.MF_HYP, np.column_stack((X_L, y_L)))
m.updt_info(X_L, y_L, X_H, y_H)
e = m.engine
e.ensure_factor()
lib = nat.lib(); st = nat.stream_ptr(); pp = ctypes.byref(e.pstruct); npad, ld = e.npad, e.cap
B0 = torch.randn(npad, max(R, 64), dtype=torch.float64, device="cuda")
sw = torch.empty(int(lib.mfgp_cholesky_solve_workspace_bytes(npad, max(R, 64))) // 8 + 8, dtype=torch.float64, device="cuda")
for rep in range(3):
    lib.mfgp_build_train_cov(nat.ptr(e.Xt), e.NL, e.NH, pp, nat.ptr(e.K), npad, ld, nat.ptr(e.Tt), st)
    B = B0.clone(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
    if R: lib.mfgp_cholesky_solve(nat.ptr(e.K), npad, ld, nat.ptr(e.W), ld, nat.ptr(e.info), nat.ptr(B), R, R, nat.ptr(sw), sw.numel() * 8, st)
    else: lib.mfgp_cholesky(nat.ptr(e.K), npad, ld, nat.ptr(e.W), ld, nat.ptr(e.info), nat.ptr(e.work), st)
    b.record(); torch.cuda.synchronize()
nbk = npad // 64
buf = np.zeros((nbk, 16), dtype=np.int64)
n = lib.mfgp_debug_chol_trace(buf.ctypes.data_as(ctypes.c_void_p), nbk)
gt = buf[:n, :7].astype(float)
t_end = gt[:, 6]
print(f"N={N} R={R}: kernel {a.elapsed_time(b):.3f} ms; chain finished {1e-3 * (t_end[-1] - gt[0, 4]):.0f} us after its start")
hop = np.diff(t_end) * 1e-3
wait = (gt[1:, 1] - gt[1:, 0]) * 1e-3          # k loop done -> W of the previous column seen (chain task idle, waiting)
late = (gt[1:, 0] - t_end[:-1]) * 1e-3         # > 0: the chain task's own k loop ended AFTER the previous W was published
for g0 in range(0, n - 1, 8):
    s = slice(g0, min(g0 + 8, n - 1))
    print(f"  columns {g0 + 1:2d}..{min(g0 + 8, n - 1):2d}: hop mean {hop[s].mean():6.1f} us  (min {hop[s].min():5.1f} max {hop[s].max():6.1f});"
          f" chain task waited {np.maximum(wait[s], 0).mean():6.1f} us, was late by {np.maximum(late[s], 0).mean():6.1f} us")
