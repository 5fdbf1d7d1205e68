"""CPU: host-side logic of the drop-in modules that needs no GPU (hyper-parameter conventions, Qhull wrapper, O(A)
finishing arithmetic, decision rules, runner dispatch)."""
import numpy as np
import pytest

from oracle import coverage as ocov
from oracle import gp as ogp
from tests import synth


def test_evaluate_hyp_matches_oracle_conventions():
    from mfgp_coverage_b200.gaussian_process import evaluate_hyp, prior_variance
    for hyp in (synth.MF_HYP, synth.SF_HYP):
        for raw in (False, True):
            p = ogp.GPParams.from_hyp(hyp, raw_means=raw)
            q = evaluate_hyp(hyp, raw)
            for k in ("s_H", "l_H", "noise_H", "mean_H"):
                assert q[k] == getattr(p, k)
            if p.multi:
                for k in ("s_L", "l_L", "rho", "noise_L", "mean_L"):
                    assert q[k] == getattr(p, k)
            assert prior_variance(q) == p.k0
    with pytest.raises(TypeError):
        evaluate_hyp(np.zeros(5))


def test_bounded_voronoi_matches_oracle():
    from mfgp_coverage_b200._coverage import BoundedVoronoi
    xy = synth.grid(11)
    bbox = ocov.bounding_box_of(xy)
    seeds = synth.agents(9, 3)
    a, b = BoundedVoronoi(seeds, bbox), ocov.voronoi_bounded(seeds, bbox)
    assert np.array_equal(a.vertices, b.vertices) and a.filtered_regions == b.filtered_regions
    areas = a.areas()
    assert abs(areas.sum() - 1.1 * 1.1) < 1e-12           # cells tile the box inflated by eps/2 per side
    s, poly, off = a.flat()
    assert off[-1] == poly.shape[0] and s.shape == (9, 2)


def test_finishing_arithmetic_matches_oracle():
    from mfgp_coverage_b200._coverage import centroids_from_partials, loss_from_partials
    xy = synth.grid(21)
    f = synth.truth_function(xy)
    truth = np.column_stack((xy, f))
    bbox = ocov.bounding_box_of(xy)
    vor = ocov.voronoi_bounded(synth.agents(6, 1), bbox)
    mem = ocov.membership(vor, xy)
    mu = np.random.default_rng(0).normal(0.3, 0.2, xy.shape[0])
    cent = np.array([[mu[m].sum(), (mu[m] * xy[m, 0]).sum(), (mu[m] * xy[m, 1]).sum(), m.sum()] for m in mem])
    lossp = np.array([[(((xy[m] - vor.filtered_points[i]) ** 2).sum(axis=1) * f[m]).sum(), m.sum()] for i, m in enumerate(mem)])
    areas = np.array([ocov.poly_area(vor.cell_vertices(i)[:, 0], vor.cell_vertices(i)[:, 1]) for i in range(6)])
    assert np.max(np.abs(centroids_from_partials(cent, areas, 0, 1, 0, 1) - ocov.compute_centroids(vor, xy, mu))) < 1e-13
    assert abs(loss_from_partials(lossp, areas) - ocov.compute_loss(vor, truth)) < 1e-15
    cent[2, 3] = 0
    cent[2, :3] = 0
    assert np.all(np.isnan(centroids_from_partials(cent, areas, 0, 1, 0, 1)[2]))     # empty cell -> NaN like np.mean([])


def test_decision_rules():
    from mfgp_coverage_b200 import simulator as sim
    assert sim.choi_threshold(1.0) == 0.82 and sim.choi_double(3) == 64
    assert [sim.periodic_decision(i) for i in (0, 4, 5, 9, 10)] == [True, True, False, False, True]
    mv = np.array([[0.02], [0.04]])
    assert np.allclose(sim.todescato_prob(mv, 0.08), np.sqrt(mv / 0.16))
    with pytest.raises(TypeError):
        sim._fidelity_of(np.zeros(5))


def test_runner_rejects_unknown_algorithm():
    from mfgp_coverage_b200 import runner
    with pytest.raises(ValueError):
        runner.run_sim(("x", "nonsense", 0, 1, 2, None, 0.1, None, None, False, None, True))
