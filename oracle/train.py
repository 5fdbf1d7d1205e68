"""Oracle: negative log marginal likelihood of the reference's GPs and its analytic gradient (TEST INFRASTRUCTURE).

Restates /root/reference/gaussian_process.py `SFGP.likelihood` (:81-105) and `MFGP.likelihood` (:344-384):
    NLML(hyp) = 1/2 y^T K^-1 y + sum(log diag L) + 1/2 N log(2 pi),    L = chol(K + jitter I),
with the current exp() mean convention (mean = exp(hyp[0]); MF: mean_L = exp(hyp[0]), mean_H = rho mean_L + exp(hyp[3]))
and the K assembly of :99-100 / :373-380.  The reference differentiates this with autograd (`value_and_grad`, :118,
:397; autograd is not installed here), so the gradient below is the closed form of the same function:
    dNLML/dh = 1/2 sum_ij (K^-1 - alpha alpha^T)_ij dK_ij/dh - alpha^T dm/dh,        alpha = K^-1 (y - m),
checked against central finite differences of the live reference's own `likelihood` (tests/test_train.py).
`train` mirrors `SFGP.train` / `MFGP.train` (:107-119, :386-399): scipy L-BFGS-B with the analytic gradient.
"""
import numpy as np
from scipy.linalg import solve_triangular

from . import gp as ogp


def _sqdist_scaled(X, Xp, length):
    a, b = X / length, Xp / length
    d = a[:, None, :] - b[None, :, :]
    return np.sum(d ** 2, axis=2)


def nlml_and_grad(hyp, X_L, y_L, X_H, y_H, jitter=ogp.JITTER):
    """(NLML, gradient[len(hyp)]) for hyp of length 4 (SF: all data in X_H / y_H, X_L empty) or 9 (MF)."""
    hyp = np.asarray(hyp, dtype=np.float64).reshape(-1)
    X_L = np.asarray(X_L, dtype=np.float64).reshape(-1, 2)
    X_H = np.asarray(X_H, dtype=np.float64).reshape(-1, 2)
    y_L = np.asarray(y_L, dtype=np.float64).reshape(-1, 1)
    y_H = np.asarray(y_H, dtype=np.float64).reshape(-1, 1)
    NL, NH = X_L.shape[0], X_H.shape[0]
    N = NL + NH
    multi = hyp.size == 9
    if not multi and hyp.size != 4:
        raise TypeError("Hyperparameters must be of length 4 (single-fidelity) or 9 (multi-fidelity)")
    X = np.vstack((X_L, X_H))
    if multi:
        m_L, s_L, l_L, e_H, s_H, l_H, rho, n_L, n_H = np.exp(hyp)
        m_H = rho * m_L + e_H
    else:
        m_H, s_H, l_H, n_H = np.exp(hyp)
        m_L, s_L, l_L, rho, n_L, e_H = 0.0, 0.0, 1.0, 1.0, 0.0, m_H
        assert NL == 0
    isL = np.arange(N) < NL
    c = np.where(isL[:, None] & isL[None, :], 1.0, np.where(isL[:, None] | isL[None, :], rho, rho ** 2))
    HH = (~isL[:, None]) & (~isL[None, :])
    qL = _sqdist_scaled(X, X, l_L)
    qH = _sqdist_scaled(X, X, l_H)
    kL = s_L * np.exp(-0.5 * qL) if multi else np.zeros((N, N))
    kH = np.where(HH, s_H * np.exp(-0.5 * qH), 0.0)
    noise = np.where(isL, n_L, n_H)
    K = c * kL + kH + np.diag(noise) + jitter * np.eye(N)
    y = np.vstack((y_L - m_L, y_H - m_H))
    L = np.linalg.cholesky(K)
    z = solve_triangular(L, y, lower=True)
    alpha = solve_triangular(L.T, z, lower=False)
    nlml = 0.5 * float(z.T @ z) + float(np.sum(np.log(np.diag(L)))) + 0.5 * np.log(2.0 * np.pi) * N
    W = solve_triangular(L, np.eye(N), lower=True)
    Q = W.T @ W - alpha @ alpha.T
    aL, aH = float(alpha[:NL].sum()), float(alpha[NL:].sum())
    if multi:
        cross = isL[:, None] ^ isL[None, :]
        g = np.array([
            -(m_L * aL + rho * m_L * aH),                                        # mu_lo
            0.5 * np.sum(Q * c * kL),                                            # s^2_lo
            0.5 * np.sum(Q * c * kL * qL),                                       # L_lo
            -e_H * aH,                                                           # mu_hi
            0.5 * np.sum(Q * kH),                                                # s^2_hi
            0.5 * np.sum(Q * kH * qH),                                           # L_hi
            0.5 * (np.sum(Q[cross] * (rho * kL)[cross]) + np.sum(Q[HH] * (2 * rho ** 2 * kL)[HH])) - rho * m_L * aH,   # rho
            0.5 * n_L * np.trace(Q[:NL, :NL]),                                   # noise_lo
            0.5 * n_H * np.trace(Q[NL:, NL:]),                                   # noise_hi
        ])
    else:
        g = np.array([-m_H * aH, 0.5 * np.sum(Q * kH), 0.5 * np.sum(Q * kH * qH), 0.5 * n_H * np.trace(Q)])
    return nlml, g


def nlml(hyp, X_L, y_L, X_H, y_H):
    return nlml_and_grad(hyp, X_L, y_L, X_H, y_H)[0]


def train(hyp0, X_L, y_L, X_H, y_H, callback=None, **options):
    """gaussian_process.py:107-119 / :386-399: L-BFGS-B on (NLML, gradient) from hyp0; returns scipy's result."""
    from scipy.optimize import minimize
    return minimize(lambda h: nlml_and_grad(h, X_L, y_L, X_H, y_H), np.asarray(hyp0, dtype=np.float64), jac=True,
                    method="L-BFGS-B", callback=callback, options=options or None)
