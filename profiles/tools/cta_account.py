"""Per-CTA time accounting of the tiled Cholesky (MFGP_DF_TRACE=1): share of CTA time spent waiting for tile flags, in the k
loops, elsewhere; spread of the CTAs' exit times.  usage: MFGP_DF_TRACE=1 python cta_account.py [N=4096] [R=1344]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
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
lib = nat.lib(); st = nat.stream_ptr(); pp = ctypes.byref(e.pstruct); npad, ld = e.npad, e.cap
B0 = torch.randn(npad, max(R, 64), dtype=torch.float64, device="cuda")
sw = torch.empty(int(lib.mfgp_cholesky_solve_workspace_bytes(npad, max(R, 64))) // 8 + 8, dtype=torch.float64, device="cuda")
for rep in range(3):
    lib.mfgp_build_train_cov(nat.ptr(e.Xt), e.NL, e.NH, pp, nat.ptr(e.K), npad, ld, nat.ptr(e.Tt), st)
    B = B0.clone(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
    if R == 0:
        lib.mfgp_cholesky(nat.ptr(e.K), npad, ld, nat.ptr(e.W), ld, nat.ptr(e.info), nat.ptr(e.work), st)
    else:
        lib.mfgp_cholesky_solve(nat.ptr(e.K), npad, ld, nat.ptr(e.W), ld, nat.ptr(e.info), nat.ptr(B), R, R, nat.ptr(sw), sw.numel() * 8, st)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
buf = np.zeros((1024, 8), dtype=np.int64)
n = lib.mfgp_debug_chol_trace(buf.ctypes.data_as(ctypes.c_void_p), -1)
c = buf[:n].astype(float)
tot = c[:, 2].sum()
print(f"N={N} R={R}: {ms:.3f} ms, {n} CTAs, info={int(e.info.item())}")
print(f"  CTA time: waiting for flags {c[:,0].sum()/tot:.1%}, k loops (incl. their waits) {c[:,1].sum()/tot:.1%}, "
      f"k loops without waits ~{(c[:,1].sum()-c[:,0].sum())/tot:.1%} (upper bound: the epilogue's diagonal wait is outside the loop)")
print(f"  per CTA total clocks: min {c[:,2].min():.0f} median {np.median(c[:,2]):.0f} max {c[:,2].max():.0f}  (kernel {ms*1e-3*1.965e9:.0f} clk)")
ex = c[:, 4] - c[:, 4].min()
print(f"  exit times after the first CTA's exit [us]: median {np.median(ex)/1e3:.1f}, p90 {np.percentile(ex,90)/1e3:.1f}, max {ex.max()/1e3:.1f}")
print(f"  tasks per CTA: min {c[:,3].min():.0f} median {np.median(c[:,3]):.0f} max {c[:,3].max():.0f}")
# wait share split: per-SM pairs
w = c[:, 0] / c[:, 2]
print(f"  wait share per CTA: p10 {np.percentile(w,10):.1%} median {np.median(w):.1%} p90 {np.percentile(w,90):.1%}")
