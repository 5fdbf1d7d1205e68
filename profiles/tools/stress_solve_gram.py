"""Stress of mfgp_cholesky_solve_gram over random sizes (block columns 1..48, right-hand-side tiles 1..16, every Gram group
size / lead combination in turn): every launch must terminate, report info = 0 and reproduce L, Y and the lower tiles of
M = Y^T Y (against LAPACK for the small cases, against the plain entry point for the rest).  usage: stress_solve_gram.py [cases=120]"""
import sys, os, ctypes, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from mfgp_coverage_b200 import _native as nat
lib = nat.lib(); st = nat.stream_ptr()
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.default_rng(2026)
worst = 0.0
GUARD = 512          # doubles of canary on either side of every buffer the kernel writes (compute-sanitizer is closed on this pool)


def guarded(n_doubles, fill):
    """A buffer of n doubles inside a larger allocation whose margins hold a canary pattern."""
    big = torch.full((n_doubles + 2 * GUARD,), -7.25e300, dtype=torch.float64, device="cuda")
    big[GUARD:GUARD + n_doubles] = fill
    return big, big[GUARD:GUARD + n_doubles]


def intact(big, n_doubles):
    return bool((big[:GUARD] == -7.25e300).all().item()) and bool((big[GUARD + n_doubles:] == -7.25e300).all().item())


t0 = time.time()
for case in range(cases):
    nb = int(rng.integers(1, 49)); npad = 64 * nb; R = 64 * int(rng.integers(1, 17))
    A = torch.randn(npad, npad, dtype=torch.float64, device="cuda", generator=None)
    K0 = (A @ A.T) / npad + torch.eye(npad, dtype=torch.float64, device="cuda") * (0.05 + float(rng.random()))
    B0 = torch.randn(npad, R, dtype=torch.float64, device="cuda")
    out = {}
    for gram in (False, True):
        gK, K = guarded(npad * npad, 0.0); K = K.view(npad, npad); K.copy_(K0)
        gB, B = guarded(npad * R, 0.0); B = B.view(npad, R); B.copy_(B0)
        gW, W = guarded(npad * npad, 0.0); W = W.view(npad, npad)
        gM, M = guarded(R * R, float("nan")); M = M.view(R, R)
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        nsw = int(lib.mfgp_cholesky_solve_gram_workspace_bytes(npad, R)) // 8          # exactly what the library asks for
        gS, sw = guarded(nsw, 0.0)
        rc = lib.mfgp_cholesky_solve_gram(nat.ptr(K), npad, npad, nat.ptr(W), npad, nat.ptr(info), nat.ptr(B), R, R,
                                          nat.ptr(M) if gram else None, R, nat.ptr(sw), nsw * 8, st)
        torch.cuda.synchronize()
        assert rc == 0 and int(info.item()) == 0, (case, nb, R, rc, int(info.item()))
        assert intact(gK, npad * npad) and intact(gB, npad * R) and intact(gW, npad * npad) and intact(gM, R * R) and intact(gS, nsw), \
            (case, nb, R, "a write landed outside its buffer")
        K, B, M = K.clone(), B.clone(), M.clone()
        out[gram] = (torch.tril(K), B, M)
    assert torch.equal(out[True][0], out[False][0]) and torch.equal(out[True][1], out[False][1]), (case, "gram tasks changed L or Y")
    L, Y, M = out[True]
    Lref = torch.linalg.cholesky(K0)
    eL = float((L - Lref).abs().max() / Lref.abs().max())
    Yref = torch.linalg.solve_triangular(Lref, B0, upper=False)
    eY = float((Y - Yref).abs().max() / Yref.abs().max())
    blk = torch.arange(R, device="cuda") // 64
    low = blk[:, None] >= blk[None, :]
    Mref = Y.T @ Y
    eM = float((M - Mref).abs()[low].max() / Mref.abs().max())
    assert eL < 1e-11 and eY < 1e-9 and eM < 1e-12 and bool(torch.isnan(M[~low]).all()), (case, nb, R, eL, eY, eM)
    worst = max(worst, eL, eM)
print(f"{cases} cases ok (canaries around K, W, B, M and the workspace intact) in {time.time() - t0:.1f} s; worst relative error of L / M: {worst:.2e}")
