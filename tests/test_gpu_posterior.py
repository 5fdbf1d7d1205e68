"""GPU parity: GP fit + posterior through the C-ABI vs the oracle (fp64, tolerance 1e-9 relative to k(0))."""
import numpy as np
import pytest

from oracle import gp as ogp
from tests import synth

pytestmark = pytest.mark.gpu
TOL = 1e-9      # BASELINE.json north_star: posterior mean / variance within a relative 1e-9 in fp64


def _models(hyp, X_L, y_L, X_H, y_H):
    from mfgp_coverage_b200.gaussian_process import MFGP, SFGP
    p = ogp.GPParams.from_hyp(hyp)
    om = ogp.Model(p, X_L, y_L, X_H, y_H)
    om.updt_info()
    if p.multi:
        m = MFGP(X_L, y_L, X_H, y_H, 1, 1)
        m.hyp = hyp
        m.updt_info(X_L, y_L, X_H, y_H)
    else:
        m = SFGP(X_H, y_H, 1)
        m.hyp = hyp
        m.updt_info(X_H, y_H)
    return p, om, m


@pytest.mark.parametrize("n,N,multi", [(51, 9, True), (51, 40, True), (51, 130, True), (64, 300, True),
                                       (51, 64, False), (51, 200, False), (96, 700, True), (64, 1100, True)])
def test_posterior_matches_oracle(n, N, multi):
    xy = synth.grid(n)
    f = synth.truth_function(xy)
    X_L, y_L, X_H, y_H = synth.training_set(xy, f, N, multi=multi)
    hyp = synth.MF_HYP if multi else synth.SF_HYP
    p, om, m = _models(hyp, X_L, y_L, X_H, y_H)
    mu_o, var_o = om.predict(xy)
    mu, var = m.predict(xy)
    assert mu.shape == (xy.shape[0], 1) and var.shape == (xy.shape[0],)
    scale = p.k0
    assert np.max(np.abs(var - var_o)) <= TOL * scale
    assert np.max(np.abs(mu[:, 0] - mu_o)) <= TOL * max(1.0, np.max(np.abs(mu_o)))
    # Cholesky factor itself
    L = m.factor()
    assert np.max(np.abs(L - om.L)) <= 1e-10 * np.max(np.abs(om.L))


def test_empty_model_is_the_prior():
    from mfgp_coverage_b200.gaussian_process import MFGP, SFGP
    xy = synth.grid(20)
    m = MFGP(np.empty([0, 2]), np.empty([0, 1]), np.empty([0, 2]), np.empty([0, 1]), 1, 1)
    m.hyp = synth.MF_HYP
    mu, var = m.predict(xy)
    p = ogp.GPParams.from_hyp(synth.MF_HYP)
    assert np.all(var == p.k0) and np.all(mu == p.mean_H)
    s = SFGP(np.empty([0, 2]), np.empty([0, 1]), 1)
    s.hyp = synth.SF_HYP
    mu, var = s.predict(xy)
    ps = ogp.GPParams.from_hyp(synth.SF_HYP)
    assert np.all(var == ps.k0) and np.all(mu == ps.mean_H)


def test_append_equals_refit_and_deepcopy():
    import copy
    from mfgp_coverage_b200.gaussian_process import MFGP
    xy = synth.grid(40)
    f = synth.truth_function(xy)
    X_L, y_L, X_H, y_H = synth.training_set(xy, f, 120)
    hyp = synth.MF_HYP
    m = MFGP(X_L, y_L, X_H[:50], y_H[:50], 1, 1)
    m.hyp = hyp
    m.updt_info(X_L, y_L, X_H[:50], y_H[:50])
    snap = copy.deepcopy(m)
    m.updt_hifi(X_H[50:], y_H[50:])
    m.updt_hifi(np.empty([0, 2]), np.empty([0, 1]))        # the reference refits even with nothing new
    p = ogp.GPParams.from_hyp(hyp)
    om = ogp.Model(p, X_L, y_L, X_H, y_H)
    om.updt_info()
    mu_o, var_o = om.predict(xy)
    mu, var = m.predict(xy)
    assert np.max(np.abs(var - var_o)) <= TOL * p.k0
    assert np.max(np.abs(mu[:, 0] - mu_o)) <= TOL
    om2 = ogp.Model(p, X_L, y_L, X_H[:50], y_H[:50])
    om2.updt_info()
    mu2_o, var2_o = om2.predict(xy)
    mu2, var2 = snap.predict(xy)                             # the deep copy kept the old state
    assert np.max(np.abs(var2 - var2_o)) <= TOL * p.k0
    assert np.max(np.abs(mu2[:, 0] - mu2_o)) <= TOL


def test_not_spd_raises_linalgerror():
    """np.linalg.cholesky raises LinAlgError on a non-positive pivot (gaussian_process.py:254); so must the device path."""
    from mfgp_coverage_b200._engine import DeviceGP
    from mfgp_coverage_b200.gaussian_process import evaluate_hyp
    params = evaluate_hyp(synth.SF_HYP)
    params["noise_H"] = -1.0          # diagonal = scale - 1 < 0: indefinite on purpose
    X = np.array([[0.1, 0.1], [0.4, 0.2], [0.5, 0.5]])
    eng = DeviceGP()
    with pytest.raises(np.linalg.LinAlgError):
        eng.fit(X, np.zeros((3, 1)), 0, 3, params)


@pytest.mark.parametrize("N,multi", [(130, True), (600, True), (200, False)])
def test_general_point_list_path(N, multi):
    """Scattered (non tensor-grid) points and the forced general path: one exp per (point, sample) pair."""
    from mfgp_coverage_b200._engine import detect_tensor_grid
    xy = synth.grid(40)
    f = synth.truth_function(xy)
    X_L, y_L, X_H, y_H = synth.training_set(xy, f, N, multi=multi)
    hyp = synth.MF_HYP if multi else synth.SF_HYP
    p, om, m = _models(hyp, X_L, y_L, X_H, y_H)
    pts = np.random.default_rng(3).random((3001, 2))
    assert detect_tensor_grid(pts) is None
    for q, sep in ((pts, True), (xy, False)):
        m.use_separable = sep
        mu_o, var_o = om.predict(q)
        mu, var = m.predict(q)
        assert np.max(np.abs(var - var_o)) <= TOL * p.k0
        assert np.max(np.abs(mu[:, 0] - mu_o)) <= TOL * max(1.0, np.max(np.abs(mu_o)))


def test_separable_path_on_a_grid_slice():
    """Grid sharding: a contiguous slice [lo, hi) of a tensor grid goes through the separable path with the full axes."""
    import torch
    from mfgp_coverage_b200._coverage import CoverageGrid
    from mfgp_coverage_b200._engine import TensorAxes, detect_tensor_grid
    xy = synth.grid(37)               # 37 is not a multiple of the 32-point CTA tile: tiles straddle grid rows
    f = synth.truth_function(xy)
    X_L, y_L, X_H, y_H = synth.training_set(xy, f, 150)
    p, om, m = _models(synth.MF_HYP, X_L, y_L, X_H, y_H)
    ux, uy = detect_tensor_grid(xy)
    assert len(ux) == 37 and len(uy) == 37
    mu_o, var_o = om.predict(xy)
    lo, hi = 401, 1203
    axes = TensorAxes(ux, uy, torch.device("cuda"))
    g = CoverageGrid(xy[lo:hi], f[lo:hi], base_index=lo, axes=axes)
    mu, var = m.predict_device(g.xy, grid=g)
    assert np.max(np.abs(var.cpu().numpy() - var_o[lo:hi])) <= TOL * p.k0
    assert np.max(np.abs(mu.cpu().numpy() - mu_o[lo:hi])) <= TOL
