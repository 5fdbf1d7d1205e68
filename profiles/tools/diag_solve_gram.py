"""mfgp_cholesky_solve_gram: correctness of L, Y and M = Y^T Y (lower 64x64 tiles) and device time against the plain solve
followed by nothing.  usage: diag_solve_gram.py [N=4096] [R=1344]   (MFGP_DF_MG, MFGP_DF_MLEAD tune the Gram tasks)"""
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
def build():
    lib.mfgp_build_train_cov(nat.ptr(e.Xt), e.NL, e.NH, pp, nat.ptr(e.K), npad, ld, nat.ptr(e.Tt), st)
B0 = torch.randn(npad, R, dtype=torch.float64, device="cuda")
sw = torch.empty(int(lib.mfgp_cholesky_solve_gram_workspace_bytes(npad, R)) // 8 + 8, dtype=torch.float64, device="cuda")
M = torch.full((R, R), float("nan"), dtype=torch.float64, device="cuda")
for gram in (0, 1):
    ts = []
    for rep in range(4):
        build(); B = B0.clone(); M.fill_(float("nan")); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
        rc = lib.mfgp_cholesky_solve_gram(nat.ptr(e.K), npad, ld, nat.ptr(e.W), ld, nat.ptr(e.info), nat.ptr(B), R, R,
                                          nat.ptr(M) if gram else None, R, nat.ptr(sw), sw.numel() * 8, st)
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    msg = f"N={N} R={R} gram={gram}: rc={rc} info={int(e.info.item())} device {min(ts):.3f} ms (runs {['%.2f' % t for t in ts]})"
    if gram:
        Y = B
        Mref = Y.T @ Y
        blk = torch.arange(R, device="cuda") // 64
        low = blk[:, None] >= blk[None, :]
        err = ((M - Mref).abs()[low]).max().item() / Mref.abs().max().item()
        msg += f" errM={err:.2e} (lower tiles; NaN left above: {bool(torch.isnan(M[~low]).all().item())})"
        M2 = M.clone()
        build(); B = B0.clone(); M.fill_(float("nan"))
        lib.mfgp_cholesky_solve_gram(nat.ptr(e.K), npad, ld, nat.ptr(e.W), ld, nat.ptr(e.info), nat.ptr(B), R, R, nat.ptr(M), R, nat.ptr(sw), sw.numel() * 8, st)
        torch.cuda.synchronize()
        msg += f" bitwise repeatable: {bool((M[low] == M2[low]).all().item())}"
    print(msg, flush=True)
