"""Time the GP fit (K assembly, Cholesky, inverse, whiten) at N (default 4096); check L against numpy."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import synth
from mfgp_coverage_b200 import simulator as sim, _native as nat
from oracle import gp as ogp

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
X_L, y_L, X_H, y_H = synth.training_set(synth.grid(256), synth.truth_function(synth.grid(256)), N)
m = sim.init_MFGP(synth.MF_HYP, np.column_stack((X_L, y_L)))
m.updt_info(X_L, y_L, X_H, y_H)
e = m.engine
lib = nat.lib(); st = nat.stream_ptr(); pp = ctypes.byref(e.pstruct); npad, ld = e.npad, e.cap
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(3):
    t = [ev() for _ in range(5)]
    t[0].record()
    lib.mfgp_build_train_cov(nat.ptr(e.Xt), e.NL, e.NH, pp, nat.ptr(e.K), npad, ld, nat.ptr(e.Tt), st); t[1].record()
    lib.mfgp_cholesky(nat.ptr(e.K), npad, ld, nat.ptr(e.W), ld, nat.ptr(e.info), nat.ptr(e.work), st); t[2].record()
    lib.mfgp_tri_inverse(nat.ptr(e.K), npad, ld, nat.ptr(e.W), ld, nat.ptr(e.work), st); t[3].record()
    lib.mfgp_whiten(nat.ptr(e.W), npad, ld, nat.ptr(e.y), e.NL, e.NH, pp, nat.ptr(e.z), st); t[4].record()
    torch.cuda.synchronize()
print(f"N={N} build_cov {t[0].elapsed_time(t[1]):.3f} ms  cholesky {t[1].elapsed_time(t[2]):.3f} ms  "
      f"tri_inverse {t[2].elapsed_time(t[3]):.3f} ms  whiten {t[3].elapsed_time(t[4]):.3f} ms  total {t[0].elapsed_time(t[4]):.3f} ms")
p = ogp.GPParams.from_hyp(synth.MF_HYP)
L = np.linalg.cholesky(ogp.train_cov(p, X_L, X_H))
Ld = torch.tril(e.K[:N, :N]).cpu().numpy()
W = torch.tril(e.W[:N, :N]).cpu().numpy()
print("max |L - L_numpy| / max|L| =", np.abs(Ld - L).max() / np.abs(L).max(), " max |W L - I| =", np.abs(W @ L - np.eye(N)).max())
