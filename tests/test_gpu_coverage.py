"""GPU parity: coverage step (membership, loss, centroids, per-cell max variance) vs the oracle."""
import numpy as np
import pytest

from oracle import coverage as ocov
from tests import synth

pytestmark = pytest.mark.gpu


def _setup(n, A, seed, on_grid=False):
    xy = synth.grid(n)
    f = synth.truth_function(xy)
    truth = np.column_stack((xy, f))
    seeds = synth.agents(A, seed)
    if on_grid:     # agents sitting on grid points: grid points lie exactly on bisectors (tie handling)
        rng = np.random.default_rng(seed)
        seeds = xy[rng.choice(xy.shape[0], A, replace=False)].copy()
    return xy, f, truth, seeds


@pytest.mark.parametrize("n,A,seed,on_grid", [(51, 8, 1, False), (51, 8, 2, True), (51, 4, 3, True), (64, 16, 4, False),
                                              (101, 64, 5, False), (51, 16, 6, True), (33, 70, 7, False)])
def test_coverage_functions_match_oracle(n, A, seed, on_grid):
    from mfgp_coverage_b200 import simulator as sim
    xy, f, truth, seeds = _setup(n, A, seed, on_grid)
    bbox = ocov.bounding_box_of(xy)
    rng = np.random.default_rng(seed + 100)
    mu = rng.normal(0.3, 0.2, xy.shape[0])          # weights may be negative (posterior mean)
    var = rng.random(xy.shape[0])
    var[rng.choice(xy.shape[0], 50)] = var.max()    # plant exact duplicates of the maximum: first index must win
    ovor = ocov.voronoi_bounded(seeds, bbox)
    vor = sim.voronoi_bounded(seeds, bbox)
    # membership bit-exact, ties included
    want = ocov.membership(ovor, xy)
    from mfgp_coverage_b200._coverage import CoverageGrid
    res = CoverageGrid(xy, f).assign_reduce(lloyd_vor=vor, want_members=True)
    m = res["members"].cpu().numpy().view(np.uint64)
    got = np.stack([((m[:, i // 64] >> np.uint64(i % 64)) & np.uint64(1)).astype(bool) for i in range(A)])
    assert np.array_equal(got, want)
    # reductions
    loss_o = ocov.compute_loss(ovor, truth)
    loss = sim.compute_loss(vor, truth)
    assert abs(loss - loss_o) <= 1e-9 * abs(loss_o)
    cen_o = ocov.compute_centroids(ovor, xy, mu.reshape(-1, 1))
    cen = sim.compute_centroids(vor, xy, mu.reshape(-1, 1))
    assert np.max(np.abs(cen - cen_o)) <= 1e-9
    xy_o, mv_o, idx_o = ocov.compute_max_var(ovor, truth, var)
    xy_g, mv_g = sim.compute_max_var(vor, truth, var)
    assert np.array_equal(xy_g, xy_o) and np.array_equal(mv_g, mv_o)


def test_in_polygon_and_clusters():
    from mfgp_coverage_b200 import simulator as sim
    xy, f, truth, seeds = _setup(51, 8, 11, True)
    bbox = ocov.bounding_box_of(xy)
    ovor = ocov.voronoi_bounded(seeds, bbox)
    vor = sim.voronoi_bounded(seeds, bbox)
    for i in range(8):
        v = ovor.cell_vertices(i)
        assert np.array_equal(sim.in_polygon(xy[:, 0], xy[:, 1], v[:, 0], v[:, 1]),
                              ocov.in_polygon(xy[:, 0], xy[:, 1], v[:, 0], v[:, 1]))
    pts = xy[np.random.default_rng(0).choice(xy.shape[0], 40, replace=False)]
    co = ocov.compute_sample_clusters(ovor, pts)
    cg = sim.compute_sample_clusters(vor, pts)
    assert all(np.array_equal(a, b) for a, b in zip(co, cg))


def test_empty_cell_raises_like_numpy():
    from mfgp_coverage_b200 import simulator as sim
    xy = synth.grid(5)
    truth = np.column_stack((xy, np.ones(25)))
    seeds = np.array([[0.5, 0.5], [0.5001, 0.5001], [0.1, 0.9]])       # the sliver between two close seeds is empty?
    seeds = np.array([[0.1, 0.1], [0.12, 0.12], [0.125, 0.125], [0.9, 0.9]])
    bbox = ocov.bounding_box_of(xy)
    vor = sim.voronoi_bounded(seeds, bbox)
    ovor = ocov.voronoi_bounded(seeds, bbox)
    counts = ocov.membership(ovor, xy).sum(axis=1)
    if counts.min() > 0:
        pytest.skip("configuration has no empty cell")
    with pytest.raises(ValueError):
        sim.compute_max_var(vor, truth, np.ones(25))
    assert np.isnan(sim.compute_loss(vor, truth)) == np.isnan(ocov.compute_loss(ovor, truth))


@pytest.mark.parametrize("A,seed", [(4, 0), (8, 1), (16, 2), (64, 3), (200, 4)])
def test_device_voronoi_clip_matches_qhull(A, seed):
    """cov_voronoi_clip (half-plane clipping on the device) vs the reference's mirrored-seed Qhull diagram
    (simulator.py:154-191): same cells -- vertex sets to 1e-12, shoelace areas to 1e-11 relative."""
    from mfgp_coverage_b200 import _coverage as cv
    seeds = synth.agents(A, seed)
    bbox = np.array([0.0, 1.0, 0.0, 1.0])
    q = ocov.voronoi_bounded(seeds, bbox)
    c = cv.ClippedVoronoi(seeds, bbox)
    assert len(c) == len(q.filtered_regions) == A
    qa = np.array([ocov.poly_area(q.cell_vertices(i)[:, 0], q.cell_vertices(i)[:, 1]) for i in range(A)])
    assert np.max(np.abs(c.areas() - qa) / qa) <= 1e-11
    assert abs(c.areas().sum() - 1.1 ** 2) <= 1e-12                   # the cells tile the box inflated by eps/2
    for i in range(A):
        vq, vc = q.cell_vertices(i), c.cell_vertices(i)
        assert vq.shape == vc.shape, i
        d = np.sqrt(((vq[:, None, :] - vc[None, :, :]) ** 2).sum(axis=2))
        assert d.min(axis=1).max() <= 1e-12 and d.min(axis=0).max() <= 1e-12


def test_coverage_step_with_device_voronoi(monkeypatch):
    """One fused coverage pass on device-built cells + device finishing (cov_finish) vs the oracle on Qhull cells."""
    from mfgp_coverage_b200 import simulator as sim
    monkeypatch.setattr(sim, "VORONOI", "clip")
    xy, f, truth, _ = _setup(96, 16, 5, False)
    rng = np.random.default_rng(8)
    mu = f + 0.1 * rng.standard_normal(f.size)
    var = rng.random(f.size)
    pos, cen = synth.agents(16, 31), synth.agents(16, 32)
    bbox = ocov.bounding_box_of(xy)
    state = sim._Sim(truth)
    import torch
    state.mu.copy_(torch.from_numpy(mu))
    state.var.copy_(torch.from_numpy(var))
    lv, pv = sim.voronoi_bounded(cen, bbox), sim.voronoi_bounded(pos, bbox)
    res = state.grid.assign_reduce(lv, pv, w=state.mu, var=state.var)
    loss, cent, mv, idx = state.grid.finish(res, lv, pv, bbox)
    olv, opv = ocov.voronoi_bounded(cen, bbox), ocov.voronoi_bounded(pos, bbox)
    assert abs(loss - ocov.compute_loss(opv, truth)) <= 1e-9 * abs(loss)
    assert np.max(np.abs(cent - ocov.compute_centroids(olv, xy, mu.reshape(-1, 1)))) <= 1e-9
    _, mv_o, idx_o = ocov.compute_max_var(olv, truth, var)
    assert np.array_equal(idx, idx_o) and np.array_equal(mv, mv_o.reshape(-1))


@pytest.mark.parametrize("n,A,seed,on_grid", [(51, 8, 2, True), (101, 64, 5, False), (64, 16, 6, True), (130, 200, 9, False),
                                              (33, 5, 10, True)])
def test_sweep_and_generic_kernels_agree(n, A, seed, on_grid):
    """Tensor-product grids take the column-sweep kernel (cov_assign_reduce_grid), arbitrary point lists the generic one
    (cov_assign_reduce): both must give the oracle's per-cell sums (1e-9) and the same arg-max indices / counts exactly."""
    import torch
    from mfgp_coverage_b200 import _coverage as cv
    from mfgp_coverage_b200 import simulator as sim
    xy, f, truth, seeds = _setup(n, A, seed, on_grid)
    rng = np.random.default_rng(seed)
    mu = f + 0.1 * rng.standard_normal(f.size)
    var = rng.random(f.size)
    pos = seeds
    cen = synth.agents(A, seed + 50)
    bbox = ocov.bounding_box_of(xy)
    lv, pv = cv.BoundedVoronoi(cen, bbox), cv.BoundedVoronoi(pos, bbox)      # Qhull's polygons: on-grid seeds put grid points ON bisectors
    res = {}
    for sweep in (True, False):
        g = cv.CoverageGrid(xy, f)
        assert g.axes is not None
        g.use_sweep = sweep
        r = g.assign_reduce(lv, pv, w=torch.from_numpy(mu).cuda(), var=torch.from_numpy(var).cuda())
        res[sweep] = {k: v.cpu().numpy() for k, v in r.items() if torch.is_tensor(v)}
    a, b = res[True], res[False]
    assert np.array_equal(a["amax_idx"], b["amax_idx"]) and np.array_equal(a["amax_val"], b["amax_val"])
    assert np.array_equal(a["cent"][:, 3], b["cent"][:, 3]) and np.array_equal(a["lossp"][:, 1], b["lossp"][:, 1])   # counts
    assert np.max(np.abs(a["cent"] - b["cent"])) <= 1e-9 * max(1.0, np.max(np.abs(b["cent"])))
    assert np.max(np.abs(a["lossp"] - b["lossp"])) <= 1e-9 * max(1.0, np.max(np.abs(b["lossp"])))
    olv, opv = ocov.voronoi_bounded(cen, bbox), ocov.voronoi_bounded(pos, bbox)
    loss = cv.loss_from_partials(a["lossp"], pv.areas())
    assert abs(loss - ocov.compute_loss(opv, truth)) <= 1e-9 * abs(loss)
    _, mv_o, idx_o = ocov.compute_max_var(olv, truth, var)
    assert np.array_equal(a["amax_idx"], idx_o)


def test_packed_partitions_and_packed_results_match_the_plain_path():
    """N > 1 plumbing on one GPU: partitions that arrive as packed device buffers (sharding.broadcast_partitions ->
    PackedPartition) and results that leave as one packed copy (results_to_host / gather_results_to_host) give exactly
    what BoundedVoronoi objects and per-tensor copies give."""
    import torch
    from mfgp_coverage_b200 import _coverage as cv
    from mfgp_coverage_b200 import sharding
    from mfgp_coverage_b200 import simulator as sim
    xy, f, truth, seeds = _setup(64, 16, 6, True)
    rng = np.random.default_rng(11)
    mu = torch.from_numpy(f + 0.1 * rng.standard_normal(f.size)).cuda()
    var = torch.from_numpy(rng.random(f.size)).cuda()
    cen = synth.agents(16, 77)
    bbox = ocov.bounding_box_of(xy)
    g = cv.CoverageGrid(xy, f)
    lv, pv = cv.BoundedVoronoi(cen, bbox), cv.BoundedVoronoi(seeds, bbox)      # what broadcast_partitions packs: Qhull's cells
    plain = g.assign_reduce(lv, pv, w=mu, var=var)
    ref = {k: plain[k].cpu().numpy().copy() for k in ("cent", "amax_val", "amax_idx", "lossp")}
    pk_p, pk_l = sharding.broadcast_partitions([seeds, cen], bbox, g.device)
    assert isinstance(pk_p, cv.PackedPartition) and len(pk_p) == len(pv) and pk_p.seeds_inside == pv.seeds_inside
    assert np.array_equal(pk_l.areas(), lv.areas()) and np.array_equal(pk_p.areas(), pv.areas())
    host = sharding.gather_results_to_host(g.assign_reduce(pk_l, pk_p, w=mu, var=var))
    for k in ref:
        assert np.array_equal(host[k], ref[k]), k
    host2 = cv.CoverageGrid.results_to_host(plain)
    for k in ref:
        assert np.array_equal(host2[k], ref[k]), k


@pytest.mark.parametrize("on_grid", [False, True])
def test_hybrid_cells_fall_back_to_qhull_exactly_when_a_pass_meets_tie_points(on_grid):
    """simulator.voronoi_bounded (default VORONOI = "auto"): cells clipped on the device, host Qhull only when the pass
    reports grid points within TIE_TOL of a bisector (their membership hangs on the polygon vertices) -- and then the result
    is what the always-Qhull path gives, bit for bit.  Off-grid seeds: no tie point, no Qhull run, areas from the device."""
    from mfgp_coverage_b200 import _coverage as cv
    from mfgp_coverage_b200 import simulator as sim
    xy, f, truth, seeds = _setup(51, 8, 4, on_grid)
    bbox = ocov.bounding_box_of(xy)
    assert sim.VORONOI == "auto"
    vor = sim.voronoi_bounded(seeds, bbox)
    assert isinstance(vor, cv.HybridVoronoi) and vor._qhull is None
    g = sim._grid_for(truth)
    host = g.reduce_to_host(loss_vor=vor)
    q = cv.BoundedVoronoi(seeds, bbox)
    want = cv.CoverageGrid.results_to_host(g.assign_reduce(loss_vor=q))
    if on_grid:
        assert want["ties"] > 0 and vor._qhull is not None         # the pass was repeated on Qhull's polygons
        assert np.array_equal(host["lossp"], want["lossp"]) and np.array_equal(vor.areas(), q.areas())
    else:
        assert host["ties"] == 0 and vor._qhull is None             # Qhull never ran
        assert np.array_equal(host["lossp"], want["lossp"])         # nearest-seed membership: identical sums
        assert np.max(np.abs(vor.areas() - q.areas()) / q.areas()) <= 1e-11
    loss = sim.compute_loss(vor, truth)
    assert abs(loss - ocov.compute_loss(ocov.voronoi_bounded(seeds, bbox), truth)) <= 1e-9 * abs(loss)
    assert np.array_equal(vor.vertices, q.vertices) and vor.filtered_regions == q.filtered_regions    # host view = Qhull's
