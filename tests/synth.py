"""Synthetic inputs of the shapes BASELINE.json names, shared by tests and bench.py."""
import numpy as np

# Data/australia9_mf_hyp.csv-like multi-fidelity hyper-parameters (log-scaled): noise_lo = log 0.01, noise_hi = log 0.1
MF_HYP = np.array([-0.8460513962053721, -2.5372906071161143, -0.5420349584953533, -18.003211372750144,
                   -3.0873278390400363, -1.6010169416738749, -0.6486247923907078, np.log(0.01), np.log(0.1)])
SF_HYP = np.array([-1.21006642748467, -2.5343035850381024, -1.4288231301940826, -1.429481927317756])


def grid(n):
    g = np.linspace(0, 1, n)
    return np.stack(np.meshgrid(g, g, indexing="ij"), axis=-1).reshape(-1, 2)    # x-major like distribution.py:337-339


def truth_function(xy):
    """Sum of exponential bumps (distribution.py:40-71 style), normalised to [1e-4, 1]."""
    centres = np.array([[0.2, 0.25], [0.75, 0.7], [0.55, 0.2], [0.3, 0.8]])
    widths = np.array([0.12, 0.18, 0.08, 0.1])
    f = np.zeros(xy.shape[0])
    for c, w in zip(centres, widths):
        f += np.exp(-np.sum((xy - c) ** 2, axis=1) / (2 * w * w))
    f = (f - f.min()) / (f.max() - f.min())
    return 1e-4 + (1 - 1e-4) * f


def training_set(xy, f, N, seed=1234, multi=True):
    """N_L = N/4 lofi points uniform in the unit square, N_H = 3N/4 distinct hifi grid points (MF); all hifi (SF)."""
    rng = np.random.default_rng(seed)
    NL = N // 4 if multi else 0
    NH = N - NL
    X_L = rng.random((NL, 2))
    y_L = 0.8 * truth_function(X_L) + 0.05 + rng.normal(0, 0.01, NL) if NL else np.empty(0)
    idx = rng.choice(xy.shape[0], NH, replace=False)
    X_H = xy[idx]
    y_H = f[idx] + rng.normal(0, 0.1, NH)
    return X_L, y_L.reshape(-1, 1), X_H, y_H.reshape(-1, 1)


def agents(A, seed=7):
    return np.random.default_rng(seed).random((A, 2))
