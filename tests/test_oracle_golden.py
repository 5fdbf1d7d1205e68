"""CPU: the oracle (numpy restatement) against every committed known-answer vector: outputs of the UNMODIFIED
reference (ref_*.npz, produced by oracle/make_golden.py) and the reference's own logged CSVs (logged_*.npz).
This is what pins the oracle."""
import os
import random

import numpy as np
import pytest

from oracle import algorithms as oalg
from oracle import coverage as ocov
from oracle import gp as ogp

TOL = 1e-12


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_oracle_gp_vs_live_reference(golden_dir):
    g = _load(golden_dir, "ref_gp_cases.npz")
    xs = g["xs"]
    for k in range(int(g["ncases"])):
        p = ogp.GPParams.from_hyp(g[f"c{k}_hyp"])
        m = ogp.Model(p, g[f"c{k}_XL"], g[f"c{k}_yL"], g[f"c{k}_XH"], g[f"c{k}_yH"])
        m.updt_info()
        mu, var = m.predict(xs)
        assert np.max(np.abs(var - g[f"c{k}_var"])) <= TOL
        assert np.max(np.abs(mu - g[f"c{k}_mu"])) <= 1e-10
        if m.N:
            assert np.max(np.abs(m.L - g[f"c{k}_L"])) <= TOL


def test_oracle_gp_vs_logged_ex_gp(golden_dir):
    g = _load(golden_dir, "logged_ex_gp.npz")
    p = ogp.GPParams.from_hyp(g["hyp"], raw_means=True)
    m = ogp.Model.from_prior(p, g["prior"])
    m.updt_info()
    mu, var = m.predict(g["xs"])
    assert np.max(np.abs(var - g["var"])) <= 1e-14
    assert np.max(np.abs(mu - g["mu"])) <= 1e-13


def test_oracle_coverage_vs_live_reference(golden_dir):
    g = _load(golden_dir, "ref_coverage_cases.npz")
    truth = g["truth"]
    xs = truth[:, :2]
    bbox = ocov.bounding_box_of(xs)
    for k in range(int(g["ncases"])):
        vor = ocov.voronoi_bounded(g[f"c{k}_seeds"], bbox)
        want = np.unpackbits(g[f"c{k}_member"], axis=1)[:, :xs.shape[0]].astype(bool)
        assert np.array_equal(ocov.membership(vor, xs), want)
        assert abs(ocov.compute_loss(vor, truth) - float(g[f"c{k}_loss"])) <= 1e-15
        assert np.max(np.abs(ocov.compute_centroids(vor, xs, g[f"c{k}_mu"]) - g[f"c{k}_cent"])) <= 1e-14
        axy, mv, _ = ocov.compute_max_var(vor, truth, g[f"c{k}_var"])
        assert np.array_equal(axy, g[f"c{k}_argmax_xy"]) and np.array_equal(mv[:, 0], g[f"c{k}_maxvar"])


def test_oracle_lloyd_vs_logged_csv(golden_dir):
    g = _load(golden_dir, "logged_australia6_lloyd.npz")
    truth = _load(golden_dir, "inputs_australia6.npz")["truth"]
    pos = g["s0_pos"][0].copy()
    T, A = 40, pos.shape[0]
    loss_log, agent_log, _ = oalg.lloyd(0, T, A, pos, truth)
    loss = np.array([r["Loss"] for r in loss_log])
    cen = np.array([[r["XCentroid"], r["YCentroid"]] for r in agent_log]).reshape(T, A, 2)
    assert np.max(np.abs(loss - g["s0_loss"][:T]) / g["s0_loss"][:T]) <= 1e-12
    assert np.max(np.abs(cen - g["s0_cent"][:T])) <= 1e-13


@pytest.mark.parametrize("ds,name,algo,hyp_key,use_prior", [
    ("australia6", "lloyd", "lloyd", None, False), ("australia6", "todescato_hmf", "todescato", "mf_hyp", True),
    ("australia6", "todescato_nsf", "todescato", "sf_hyp", False), ("australia6", "periodic_hsf", "periodic", "sf_hyp", True),
    ("australia6", "choi_hmf", "choi", "mf_hyp", True),
    # (australia6 choi_nsf is left to the GPU test: with a null prior the planner's picks hinge on variances that tie
    #  bit-exactly at mirror-image points only in the reference's own LU arithmetic -- DESIGN "Arg-max ties" case (a); the
    #  product's tolerance rule reproduces them, the oracle's plain np.argmax on its triangular solves does not)
    ("australia3", "todescato_nsf", "todescato", "sf_hyp", False), ("australia3", "choi_nsf", "choi", "sf_hyp", False)])
def test_oracle_loops_vs_seeded_reference_runs(golden_dir, ds, name, algo, hyp_key, use_prior):
    g = _load(golden_dir, "ref_runs.npz")
    inp = _load(golden_dir, f"inputs_{ds}.npz")
    key = f"{ds}_{name}"
    A, T, seed = (int(v) for v in g[f"{key}_meta"])
    pos = g[f"{key}_start"].copy()
    if algo == "lloyd":
        logs = oalg.lloyd(0, T, A, pos, inp["truth"])
    else:
        r = random.Random(seed)
        for _ in range(2 * A):
            r.random()      # the reference drew the start positions from the same stream first (runner.py:41-42)
        prior = inp["prior"] if use_prior else None
        logs = getattr(oalg, algo)(0, T, A, pos, inp["truth"], 0.1, prior, inp[hyp_key], r, np.random.default_rng(seed))
    loss = np.array([r_["Loss"] for r_ in logs[0]])
    # Loss is chaotic under ulp changes of the seeds when >= 2 explorers sit on grid points (exact bisector ties decided
    # by Qhull vertex rounding): check it on the reference's OWN logged positions instead.
    bbox = ocov.bounding_box_of(inp["truth"][:, :2])
    for t in range(len(loss)):
        replay = ocov.compute_loss(ocov.voronoi_bounded(g[f"{key}_agent"][t, :, :2], bbox), inp["truth"])
        assert abs(replay - g[f"{key}_loss"][t]) <= 1e-12, t
        if g[f"{key}_agent"][t, :, 9].sum() < 2:
            assert abs(loss[t] - g[f"{key}_loss"][t]) <= 1e-12, t
    cen = np.array([[r_["XCentroid"], r_["YCentroid"], r_["VarMax"], r_["Explore"]] for r_ in logs[1]]).reshape(len(loss), A, 4)
    assert np.max(np.abs(cen - g[f"{key}_agent"][:, :, [6, 7, 4, 9]])) <= 1e-12


@pytest.mark.parametrize("fname,ds,hyp_key,raw,use_prior,means", [
    ("logged_two_corners_hmf.npz", "two_corners", "mf_hyp", True, True, True),
    ("logged_australia6_nsf.npz", "australia6", "sf_hyp", False, False, True),
    ("logged_australia3_nsf.npz", "australia3", "sf_hyp", False, False, False)])
def test_oracle_replays_logged_gp_runs(golden_dir, fname, ds, hyp_key, raw, use_prior, means):
    """Replay of the reference's own logged todescato runs: logged samples in,
    logged VarMax / XMax (and, where the run used today's mean convention, centroids) out.  australia3_todescato_nsf
    (BASELINE config 1) predates the exp(mean) convention, so it pins the variance path only."""
    g = _load(golden_dir, fname)
    inp = _load(golden_dir, f"inputs_{ds}.npz")
    truth = inp["truth"]
    xs = truth[:, :2]
    bbox = ocov.bounding_box_of(xs)
    p = ogp.GPParams.from_hyp(inp[hyp_key], raw_means=raw)
    prior = inp["prior"] if use_prior else None
    m = ogp.Model.from_prior(p, prior)
    assert abs(p.k0 - g["var0"][0, 0]) <= 1e-12 * p.k0
    T = min(g["pos"].shape[0], 12)
    seeds = g["pos"][0]
    for t in range(T):
        smp = g["samples"][g["samples"][:, 0] == t]
        m.append(smp[:, 2:4], smp[:, 4:5])
        mu, var = m.predict(xs)
        vor = ocov.voronoi_bounded(seeds, bbox)
        axy, mv, _ = ocov.compute_max_var(vor, truth, var)
        assert np.max(np.abs(mv[:, 0] - g["varmax"][t])) <= 1e-9 * p.k0, t
        tie = np.abs(mv[:, 0] - g["varmax"][t]) <= 1e-12 * p.k0
        assert np.all((axy[:, 0] == g["xmax"][t]) | tie), t
        if means:
            assert np.max(np.abs(ocov.compute_centroids(vor, xs, mu) - g["cent"][t])) <= 1e-9, t
        seeds = g["cent"][t]


def test_fast_planner_equals_literal_planner():
    from tests import synth
    xy = synth.grid(24)
    f = synth.truth_function(xy)
    X_L, y_L, X_H, y_H = synth.training_set(xy, f, 30)
    m = ogp.Model(ogp.GPParams.from_hyp(synth.MF_HYP), X_L, y_L, X_H, y_H)
    m.updt_info()
    _, var = m.predict(xy)
    a, ia = ocov.compute_sample_points(m, xy, 0.6 * var.max())
    b, ib = ocov.compute_sample_points_fast(m, xy, 0.6 * var.max())
    assert len(ia) > 3 and np.array_equal(ia, ib)
