"""Oracle: single- and two-level multi-fidelity RBF GP posterior (numpy restatement; TEST INFRASTRUCTURE).

Follows /root/reference/gaussian_process.py:
  rbf                 <- SFGP.kernel :66-79, MFGP.kernel :329-342
  GPParams.from_hyp   <- hyper-parameter conventions :132, :248-251, :411-416, :510-514 (all log-scaled; noise is
                         added to the diagonal UN-squared; means are exp()-ed in today's code -- `raw_means=True`
                         gives the 2020 convention mean=hyp[0] that Data/ex_gp.csv and two_corners_* were logged with)
  train_cov           <- SFGP.updt_info :229-255, MFGP.updt_info :493-529
  posterior           <- SFGP.predict :121-148, MFGP.predict :401-438, restated DIAGONAL-ONLY and chunked
                         (precedent: gaussian_process_numba.py:478-503).  `exact_solve=True` keeps the reference's
                         general `np.linalg.solve` calls; the default uses triangular solves (agreement <= 3e-14).
A single-fidelity model is the N_L = 0, "no lofi term" special case of the multi-fidelity one.
"""
from dataclasses import dataclass

import numpy as np
from scipy.linalg import solve_triangular

JITTER = 1e-8   # gaussian_process.py:42, :298


def rbf(x, xp, scale, length):
    """scale * exp(-0.5 * sum_d (x_d/l - xp_d/l)^2) -- same operation order as gaussian_process.py:75-79."""
    diffs = np.expand_dims(x / length, 1) - np.expand_dims(xp / length, 0)
    return scale * np.exp(-0.5 * np.sum(diffs ** 2, axis=2))


@dataclass
class GPParams:
    """Evaluated (non-log) parameters.  SF: multi=False and only the *_H fields are used."""
    multi: bool
    s_L: float
    l_L: float
    s_H: float
    l_H: float
    rho: float
    noise_L: float
    noise_H: float
    mean_L: float
    mean_H: float

    @property
    def k0(self):
        """Prior variance k(x,x): gaussian_process.py:146 (SF), :435-436 (MF) on the diagonal."""
        if self.multi:
            return self.rho ** 2 * self.s_L + self.s_H
        return self.s_H

    @staticmethod
    def from_hyp(hyp, raw_means=False):
        hyp = np.asarray(hyp, dtype=np.float64).reshape(-1)
        if hyp.size == 4:       # [mu, s^2, L, noise]  simulator.py:83-84
            mean = hyp[0] if raw_means else np.exp(hyp[0])
            return GPParams(False, 0.0, 1.0, float(np.exp(hyp[1])), float(np.exp(hyp[2])), 1.0,
                            0.0, float(np.exp(hyp[3])), 0.0, float(mean))
        if hyp.size == 9:       # [mu_lo, s^2_lo, L_lo, mu_hi, s^2_hi, L_hi, rho, noise_lo, noise_hi]  simulator.py:53-54
            rho = np.exp(hyp[6])
            if raw_means:
                mean_L = hyp[0]
                mean_H = rho * mean_L + hyp[3]
            else:
                mean_L = np.exp(hyp[0])
                mean_H = rho * mean_L + np.exp(hyp[3])
            return GPParams(True, float(np.exp(hyp[1])), float(np.exp(hyp[2])), float(np.exp(hyp[4])),
                            float(np.exp(hyp[5])), float(rho), float(np.exp(hyp[7])), float(np.exp(hyp[8])),
                            float(mean_L), float(mean_H))
        raise TypeError("Hyperparameters must be of length 4 (single-fidelity) or 9 (multi-fidelity)")


def train_cov(p, X_L, X_H):
    """K + jitter*I exactly as the reference assembles it (before the Cholesky)."""
    X_L = np.asarray(X_L, dtype=np.float64).reshape(-1, 2)
    X_H = np.asarray(X_H, dtype=np.float64).reshape(-1, 2)
    NL, NH = X_L.shape[0], X_H.shape[0]
    N = NL + NH
    if not p.multi:
        assert NL == 0
        K = rbf(X_H, X_H, p.s_H, p.l_H) + np.eye(NH) * p.noise_H
    else:
        K_LL = rbf(X_L, X_L, p.s_L, p.l_L) + np.eye(NL) * p.noise_L
        K_LH = p.rho * rbf(X_L, X_H, p.s_L, p.l_L)
        K_HH = p.rho ** 2 * rbf(X_H, X_H, p.s_L, p.l_L) + rbf(X_H, X_H, p.s_H, p.l_H) + np.eye(NH) * p.noise_H
        K = np.vstack((np.hstack((K_LL, K_LH)), np.hstack((K_LH.T, K_HH))))
    return K + np.eye(N) * JITTER


def cholesky(K):
    return np.linalg.cholesky(K)


def cross_cov(p, X_star, X_L, X_H):
    """psi [G, N]: gaussian_process.py:139 (SF), :426-429 (MF)."""
    X_L = np.asarray(X_L, dtype=np.float64).reshape(-1, 2)
    X_H = np.asarray(X_H, dtype=np.float64).reshape(-1, 2)
    if not p.multi:
        return rbf(X_star, X_H, p.s_H, p.l_H)
    psi1 = p.rho * rbf(X_star, X_L, p.s_L, p.l_L)
    psi2 = p.rho ** 2 * rbf(X_star, X_H, p.s_L, p.l_L) + rbf(X_star, X_H, p.s_H, p.l_H)
    return np.hstack((psi1, psi2))


def centered_y(p, y_L, y_H):
    y_L = np.asarray(y_L, dtype=np.float64).reshape(-1, 1)
    y_H = np.asarray(y_H, dtype=np.float64).reshape(-1, 1)
    if not p.multi:
        return y_H - p.mean_H
    return np.vstack((y_L - p.mean_L, y_H - p.mean_H))


def posterior(p, X_star, X_L, y_L, X_H, y_H, L=None, chunk=4096, exact_solve=False):
    """Posterior mean [G] and variance [G] (the diagonal the reference's callers take: simulator.py:301,341,685,855).

    With an empty model the reference returns the constant mean and k(0) (gaussian_process.py:139-146 with N=0)."""
    X_star = np.asarray(X_star, dtype=np.float64).reshape(-1, 2)
    X_L = np.asarray(X_L, dtype=np.float64).reshape(-1, 2)
    X_H = np.asarray(X_H, dtype=np.float64).reshape(-1, 2)
    G = X_star.shape[0]
    N = X_L.shape[0] + X_H.shape[0]
    mu = np.full(G, p.mean_H, dtype=np.float64)
    var = np.empty(G, dtype=np.float64)
    if N == 0:
        # reference: k(X*,X*) diagonal = scale*exp(-0.5*0); MF: rho^2*s_L*1 + s_H*1
        var[:] = (p.rho ** 2 * (p.s_L * np.exp(-0.0)) + p.s_H * np.exp(-0.0)) if p.multi else p.s_H * np.exp(-0.0)
        return mu, var
    if L is None:
        L = cholesky(train_cov(p, X_L, X_H))
    y = centered_y(p, y_L, y_H)
    if exact_solve:
        alpha = np.linalg.solve(L.T, np.linalg.solve(L, y))
    else:
        alpha = solve_triangular(L.T, solve_triangular(L, y, lower=True), lower=False)
    for s in range(0, G, chunk):
        xs = X_star[s:s + chunk]
        psi = cross_cov(p, xs, X_L, X_H)
        mu[s:s + chunk] = p.mean_H + (psi @ alpha)[:, 0]
        if exact_solve:
            beta = np.linalg.solve(L.T, np.linalg.solve(L, psi.T))
            var[s:s + chunk] = p.k0 - np.einsum("ij,ji->i", psi, beta)
        else:
            v = solve_triangular(L, psi.T, lower=True)
            var[s:s + chunk] = p.k0 - np.einsum("ij,ij->j", v, v)
    return mu, var


class Model:
    """Mutable GP state with the reference's update verbs (updt_info / updt / updt_hifi: gaussian_process.py:229-268,
    :493-542).  New hifi points are appended at the END of [X_L; X_H]; the factor is recomputed from scratch."""

    def __init__(self, params, X_L=None, y_L=None, X_H=None, y_H=None):
        self.p = params
        self.X_L = np.empty((0, 2)) if X_L is None else np.asarray(X_L, dtype=np.float64).reshape(-1, 2)
        self.y_L = np.empty((0, 1)) if y_L is None else np.asarray(y_L, dtype=np.float64).reshape(-1, 1)
        self.X_H = np.empty((0, 2)) if X_H is None else np.asarray(X_H, dtype=np.float64).reshape(-1, 2)
        self.y_H = np.empty((0, 1)) if y_H is None else np.asarray(y_H, dtype=np.float64).reshape(-1, 1)
        self.L = np.empty((0, 0))

    @staticmethod
    def from_prior(params, prior_xyz):
        """simulator.py:47-102: a prior conditions the LOFI level of an MF model, or is the data of an SF model."""
        m = Model(params)
        if prior_xyz is not None and len(prior_xyz) > 0:
            pr = np.asarray(prior_xyz, dtype=np.float64).reshape(-1, 3)
            if params.multi:
                m.X_L, m.y_L = pr[:, :2].copy(), pr[:, 2:3].copy()
            else:
                m.X_H, m.y_H = pr[:, :2].copy(), pr[:, 2:3].copy()
        return m

    def copy(self):
        m = Model(self.p, self.X_L.copy(), self.y_L.copy(), self.X_H.copy(), self.y_H.copy())
        m.L = self.L.copy()
        return m

    @property
    def N(self):
        return self.X_L.shape[0] + self.X_H.shape[0]

    def updt_info(self):
        self.L = cholesky(train_cov(self.p, self.X_L, self.X_H)) if self.N else np.empty((0, 0))

    def append(self, X_new, y_new):
        """updt (SF) / updt_hifi (MF)."""
        X_new = np.asarray(X_new, dtype=np.float64).reshape(-1, 2)
        y_new = np.asarray(y_new, dtype=np.float64).reshape(-1, 1)
        self.X_H = np.vstack((self.X_H, X_new))
        self.y_H = np.vstack((self.y_H, y_new))
        self.updt_info()

    def predict(self, X_star, **kw):
        return posterior(self.p, X_star, self.X_L, self.y_L, self.X_H, self.y_H, L=self.L if self.N else None, **kw)
