"""Host launch time vs device time of mfgp_cholesky (classic) and mfgp_cholesky_solve (fused, with R right-hand sides)."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tests import synth
from mfgp_coverage_b200 import simulator as sim, _native as nat
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
base = synth.grid(256)
X_L, y_L, X_H, y_H = synth.training_set(base, synth.truth_function(base), N)
m = sim.init_MFGP(synth.MF_HYP, np.column_stack((X_L, y_L)))
m.updt_info(X_L, y_L, X_H, y_H)
e = m.engine
lib = nat.lib(); st = nat.stream_ptr(); pp = ctypes.byref(e.pstruct); npad, ld = e.npad, e.cap
def ev(): return torch.cuda.Event(enable_timing=True)
for R in (0, 64, 1024, 2560):
    B = torch.randn(npad * max(R, 64), dtype=torch.float64, device="cuda")
    sw = torch.empty(int(lib.mfgp_cholesky_solve_workspace_bytes(npad, max(R, 64))) // 8 + 8, dtype=torch.float64, device="cuda")
    for rep in range(3):
        lib.mfgp_build_train_cov(nat.ptr(e.Xt), e.NL, e.NH, pp, nat.ptr(e.K), npad, ld, nat.ptr(e.Tt), st)
        torch.cuda.synchronize()
        a, b = ev(), ev()
        t0 = time.perf_counter(); a.record()
        if R == 0:
            lib.mfgp_cholesky(nat.ptr(e.K), npad, ld, nat.ptr(e.W), ld, nat.ptr(e.info), nat.ptr(e.work), st)
        else:
            lib.mfgp_cholesky_solve(nat.ptr(e.K), npad, ld, nat.ptr(e.W), ld, nat.ptr(e.info), nat.ptr(B), R, R, nat.ptr(sw), sw.numel() * 8, st)
        b.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
    print(f"N={N} R={R}: host launch {1e3*(t1-t0):.2f} ms, device {a.elapsed_time(b):.2f} ms")
