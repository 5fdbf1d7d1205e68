"""GPU parity at BASELINE.json's full sizes (c3: 256x256 / N=1024 / 16 agents; c4: 1024x1024 / N=4096 / 64 agents).
The oracle cannot sweep a 1M-point grid in seconds, so: (i) oracle spot-check on a random subset of grid points with
the FULL training set, (ii) size-independent properties: grid-sharded == unsharded (indices exact), the separable and
general kernels agree, variance bounds, data-point consistency, and coverage conservation laws."""
import numpy as np
import pytest
import torch

from oracle import coverage as ocov
from oracle import gp as ogp
from tests import synth

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module", params=["c3", "c4"])
def workload(request):
    from mfgp_coverage_b200 import simulator as sim
    from mfgp_coverage_b200._coverage import CoverageGrid
    n, N, A = {"c3": (256, 1024, 16), "c4": (1024, 4096, 64)}[request.param]
    xy = synth.grid(n)
    f = synth.truth_function(xy)
    X_L, y_L, X_H, y_H = synth.training_set(xy, f, N)
    m = sim.init_MFGP(synth.MF_HYP, np.column_stack((X_L, y_L)))
    m.updt_info(X_L, y_L, X_H, y_H)
    g = CoverageGrid(xy, f)
    mu, var = m.predict_device(g.xy, grid=g)
    p = ogp.GPParams.from_hyp(synth.MF_HYP)
    return dict(n=n, N=N, A=A, xy=xy, f=f, X_L=X_L, y_L=y_L, X_H=X_H, y_H=y_H, m=m, g=g, mu=mu, var=var, p=p)


def test_oracle_spot_check_full_training_set(workload):
    w = workload
    idx = np.sort(np.random.default_rng(0).choice(w["xy"].shape[0], 3000, replace=False))
    om = ogp.Model(w["p"], w["X_L"], w["y_L"], w["X_H"], w["y_H"])
    om.updt_info()
    mu_o, var_o = om.predict(w["xy"][idx])
    mu, var = w["mu"].cpu().numpy()[idx], w["var"].cpu().numpy()[idx]
    assert np.max(np.abs(var - var_o)) <= TOL * w["p"].k0
    assert np.max(np.abs(mu - mu_o)) <= TOL * max(1.0, np.max(np.abs(mu_o)))


def test_fused_fit_path_oracle_spot_check(workload):
    """The exact path bench.py times (`value`): deferred fit -> mfgp_build_train_cov + mfgp_factored_prepare +
    mfgp_cholesky_solve (Cholesky fused with the forward substitution) + mfgp_posterior_grid_factored_solved, at the
    BASELINE sizes, against the oracle on a random subset of grid points with the FULL training set; and against the
    non-fused path (explicit inverse + W B product) on every grid point."""
    w = workload
    e = w["m"].engine
    g = w["g"]
    mu = torch.full((g.G,), float("nan"), dtype=torch.float64, device=g.device)
    var = torch.full((g.G,), float("nan"), dtype=torch.float64, device=g.device)
    e.defer_fit = True
    try:
        for rep in range(2):
            e.refactor(check=False)
            assert e._dirty
            w["m"].predict_device(g.xy, mu, var, grid=g)
            assert not e._dirty and e._w_partial and e._fplan[1] is not None        # fused + factored, not the dense kernel
        e.check_factor(force=True)
        L = torch.tril(e.K[:w["N"], :w["N"]]).cpu().numpy()
    finally:
        e.defer_fit = False
        e.refactor(check=True)      # back to the module fixture's state (explicit inverse, z = W (y - m)) for the other tests
    idx = np.sort(np.random.default_rng(0).choice(w["xy"].shape[0], 3000, replace=False))
    om = ogp.Model(w["p"], w["X_L"], w["y_L"], w["X_H"], w["y_H"])
    om.updt_info()
    mu_o, var_o = om.predict(w["xy"][idx])
    assert np.max(np.abs(var.cpu().numpy()[idx] - var_o)) <= TOL * w["p"].k0
    assert np.max(np.abs(mu.cpu().numpy()[idx] - mu_o)) <= TOL * max(1.0, np.max(np.abs(mu_o)))
    assert float((var - w["var"]).abs().max()) <= 1e-11 * w["p"].k0
    assert float((mu - w["mu"]).abs().max()) <= 1e-11
    assert np.max(np.abs(L - om.L)) <= 1e-10 * np.max(np.abs(om.L))


def test_general_kernel_agrees_with_separable_kernel(workload):
    w = workload
    G = w["xy"].shape[0]
    lo, hi = G // 3, G // 3 + 50000                     # a slice is enough: the general kernel pays one exp per pair
    xs = w["g"].xy[lo:hi].contiguous()
    w["m"].use_separable = False
    mu, var = w["m"].predict_device(xs)
    w["m"].use_separable = True
    assert float((var - w["var"][lo:hi]).abs().max()) <= 1e-12 * w["p"].k0
    assert float((mu - w["mu"][lo:hi]).abs().max()) <= 1e-12


def test_variance_bounds_and_data_consistency(workload):
    w = workload
    var = w["var"]
    assert float(var.max()) <= w["p"].k0 * (1 + 1e-12) and float(var.min()) > 0
    # at a hifi sample location the posterior variance is below the noise-limited bound k0 * noise / (k0 + noise)
    nx = w["n"]
    ij = np.rint(w["X_H"][:200] * (nx - 1)).astype(np.int64)
    flat = torch.from_numpy(ij[:, 0] * nx + ij[:, 1]).cuda()
    bound = w["p"].k0 * (w["p"].noise_H + 1e-8) / (w["p"].k0 + w["p"].noise_H + 1e-8)
    assert float(var[flat].max()) <= bound * (1 + 1e-9)


def test_grid_sharded_equals_unsharded(workload):
    """Emulates N = 4 ranks in one process: rank-local posterior + coverage partials, merged with the same functions the
    NCCL path uses (sharding.merge_argmax; sums added in rank order)."""
    from mfgp_coverage_b200 import sharding
    from mfgp_coverage_b200 import simulator as sim
    from mfgp_coverage_b200._coverage import CoverageGrid
    w = workload
    G, A = w["xy"].shape[0], w["A"]
    bbox = np.array([0.0, 1.0, 0.0, 1.0])
    lv = sim.voronoi_bounded(synth.agents(A, 8), bbox)
    pv = sim.voronoi_bounded(synth.agents(A, 7), bbox)
    k0 = w["p"].k0
    full = w["g"].assign_reduce(lv, pv, w=w["mu"], var=w["var"], amax_k0=k0, amax_rel=1e-10)
    parts = []
    for r in range(4):
        lo, hi = sharding.shard_bounds(G, 4, r)
        g = CoverageGrid(w["xy"][lo:hi], w["f"][lo:hi], base_index=lo, axes=w["g"].axes)
        mu, var = w["m"].predict_device(g.xy, grid=g)
        # the posterior is per-point; a shard re-expands the x factor on ITS interval (lower Chebyshev order, fewer
        # right-hand sides), so the two evaluations differ by rounding only
        assert float((var - w["var"][lo:hi]).abs().max()) <= 1e-12 * k0 and float((mu - w["mu"][lo:hi]).abs().max()) <= 1e-12
        parts.append(g.assign_reduce(lv, pv, w=mu, var=var, amax_k0=k0, amax_rel=1e-10))
    cent = sum(p["cent"] for p in parts)
    lossp = sum(p["lossp"] for p in parts)
    v, i = sharding.merge_argmax(torch.stack([p["amax_val"] for p in parts]), torch.stack([p["amax_idx"] for p in parts]),
                                 k0, 1e-10)
    assert torch.equal(i, full["amax_idx"]) and float((v - full["amax_val"]).abs().max()) <= 1e-12 * k0
    assert torch.equal(cent[:, 3], full["cent"][:, 3]) and torch.equal(lossp[:, 1], full["lossp"][:, 1])   # counts
    assert float(((cent - full["cent"]).abs() / (full["cent"].abs() + 1e-300)).max()) <= 1e-11
    assert float(((lossp - full["lossp"]).abs() / (full["lossp"].abs() + 1e-300)).max()) <= 1e-11


def test_coverage_conservation_and_subset_oracle(workload):
    from mfgp_coverage_b200 import simulator as sim
    w = workload
    G, A = w["xy"].shape[0], w["A"]
    bbox = np.array([0.0, 1.0, 0.0, 1.0])
    seeds = synth.agents(A, 8)
    lv = sim.voronoi_bounded(seeds, bbox)
    res = w["g"].assign_reduce(lv, lv, w=w["mu"], var=w["var"], want_members=True)
    cent = res["cent"].cpu().numpy()
    # off-grid random seeds: every grid point lies in exactly one cell
    assert cent[:, 3].sum() == G and res["lossp"].cpu().numpy()[:, 1].sum() == G
    assert abs(cent[:, 0].sum() - float(w["mu"].sum())) <= 1e-9 * float(w["mu"].abs().sum())
    assert abs(res["amax_val"].max().item() - w["var"].max().item()) == 0.0
    # membership vs the oracle's crossings test on a random subset
    idx = np.sort(np.random.default_rng(1).choice(G, 20000, replace=False))
    ovor = ocov.voronoi_bounded(seeds, bbox)
    want = ocov.membership(ovor, w["xy"][idx])
    m = res["members"].cpu().numpy().view(np.uint64)[idx]
    got = np.stack([((m[:, i // 64] >> np.uint64(i % 64)) & np.uint64(1)).astype(bool) for i in range(A)])
    assert np.array_equal(got, want)
    # per-cell arg-max: first index of the per-cell maximum, checked with numpy on the full arrays
    cell = np.argmax(res["members"].cpu().numpy().view(np.uint64)[:, 0:1] >> np.arange(64, dtype=np.uint64) & np.uint64(1), axis=1) \
        if A <= 64 else None
    var = w["var"].cpu().numpy()
    for c in range(0, A, max(1, A // 8)):
        sel = np.nonzero(cell == c)[0]
        assert res["amax_idx"][c].item() == sel[np.argmax(var[sel])]


def test_choi_period_at_c3_size():
    """BASELINE config 3 (synthetic 256x256 grid, 1024 MF training samples, 16 agents, choi_hmf): one Choi period -- greedy
    sample planner (compute_sample_points), clustering, tour planner, then the period's 8 coverage iterations -- through the
    drop-in simulator, against the oracle: identical pick sequence (V-cached CPU twin of the reference's refit-per-pick
    loop, pinned to the literal loop in tests/test_oracle_golden.py), identical clusters and tours, and the agents visit
    their tours in order.  With 1024 samples already in the model the maximum variance is 0.2 k(0), far below the first
    periods' thresholds 0.82^p k(0) (simulator.py:1015,1037), so the period's threshold is set to 0.6 max(var) -- the state
    a long run reaches after ~11 periods."""
    import random
    from mfgp_coverage_b200 import simulator as sim
    from oracle import tsp as otsp
    n, N, A = 256, 1024, 16
    xy = synth.grid(n)
    f = synth.truth_function(xy)
    truth_arr = np.column_stack((xy, f))
    X_L, y_L, X_H, y_H = synth.training_set(xy, f, N)
    p = ogp.GPParams.from_hyp(synth.MF_HYP)
    bbox = np.array([0.0, 1.0, 0.0, 1.0])
    # --- planner + clusters + tours at the period's start state
    m = sim.init_MFGP(synth.MF_HYP, np.column_stack((X_L, y_L)))
    m.updt_info(X_L, y_L, X_H, y_H)
    om = ogp.Model(p, X_L, y_L, X_H, y_H)
    om.updt_info()
    g = sim._grid_for(xy)
    _, var0 = m.predict_device(g.xy, grid=g)
    thr = 0.6 * float(var0.max())
    pts_o, idx_o = ocov.compute_sample_points_fast(om, xy, thr)
    pts_g, idx_g = sim.compute_sample_points(m, xy, thr, False, return_indices=True)
    assert 16 <= len(idx_o) <= 500
    assert np.array_equal(idx_g, idx_o) and np.array_equal(pts_g, pts_o)
    pos = synth.agents(A, 5)
    vor_g, vor_o = sim.voronoi_bounded(pos, bbox), ocov.voronoi_bounded(pos, bbox)
    cl_g, cl_o = sim.compute_sample_clusters(vor_g, pts_g), ocov.compute_sample_clusters(vor_o, pts_o)
    tours_g, tours_o = sim.compute_sample_tsp(cl_g), otsp.compute_sample_tsp(cl_o)
    assert sum(c.shape[0] for c in cl_g) == len(idx_o)
    for a, b, c, d in zip(cl_g, cl_o, tours_g, tours_o):
        assert np.array_equal(a, b) and np.array_equal(c, d)
    # --- the whole period through simulator.choi, from a model that already holds the c3 training set
    real_init, real_thr = sim._init_models, sim.choi_threshold

    def init_with_training_set(fidelity, hyp, prior):
        model, max_var_0 = real_init(fidelity, hyp, prior)
        model.updt_info(model.X_L, model.y_L, X_H, y_H)
        return model, max_var_0
    sim._init_models = init_with_training_set
    sim.choi_threshold = lambda threshold: thr
    try:
        loss_log, agent_log, sample_log = sim.choi("choi_hmf", 0, 8, A, pos.copy(), truth_arr, 0.1, np.column_stack((X_L, y_L)),
                                                   synth.MF_HYP, False, None, True, rng=random.Random(3),
                                                   noise_rng=np.random.default_rng(3))
    finally:
        sim._init_models, sim.choi_threshold = real_init, real_thr
    assert len(loss_log) == 8 and all(r["Period"] == 0 for r in loss_log)
    X = np.array([[r["X"], r["Y"]] for r in agent_log]).reshape(8, A, 2)
    for i in range(A):                       # agent i visits its tour in order from iteration 1 on
        k = min(7, tours_o[i].shape[0])
        assert np.array_equal(X[1:1 + k, i], tours_o[i][:k]), i
    taken = np.array([[r["X"], r["Y"]] for r in sample_log]).reshape(-1, 2)
    want = [tours_o[i][t - 1] for t in range(1, 8) for i in range(A) if tours_o[i].shape[0] >= t]
    assert len(want) >= 16 and np.array_equal(taken, np.array(want).reshape(-1, 2))
    assert all(np.isfinite(r["Loss"]) and r["Loss"] > 0 for r in loss_log)
