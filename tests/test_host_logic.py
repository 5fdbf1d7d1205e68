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


def test_cell_areas_and_flat_layout_match_the_reference_formula():
    """BoundedVoronoi.areas(): the vectorised shoelace equals simulator.py:127-136 per cell (1e-15 absolute: the two dot
    products of the reference are summed in a different order); flat() lists every cell's vertices in region order."""
    from mfgp_coverage_b200 import _coverage as cv
    bb = np.array([0.0, 1.0, 0.0, 1.0])
    for A, seed in ((1, 0), (2, 1), (8, 2), (64, 3), (200, 4)):
        v = cv.BoundedVoronoi(synth.agents(A, seed), bb)
        seeds, poly, off = v.flat()
        assert seeds.shape == (A, 2) and off[0] == 0 and off[-1] == poly.shape[0]
        ref = np.empty(A)
        for i in range(A):
            pts = v.vertices[v.filtered_regions[i]]
            assert np.array_equal(poly[off[i]:off[i + 1]], pts)
            x, y = pts[:, 0], pts[:, 1]
            ref[i] = 0.5 * np.abs(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1)))          # poly_area, simulator.py:127-136
        assert np.max(np.abs(v.areas() - ref)) <= 4e-15
        assert abs(v.areas().sum() - 1.1 * 1.1) <= 1e-12       # the cells tile the box inflated by eps/2 on every side


def test_packed_partition_round_trip_and_seed_summary():
    """The packed cell buffer that travels between ranks (sharding.broadcast_partitions) carries exactly the partition."""
    import torch
    from mfgp_coverage_b200 import _coverage as cv
    bb = np.array([0.0, 1.0, 0.0, 1.0])
    for A, seed in ((1, 5), (5, 6), (64, 7)):
        pts = synth.agents(A, seed)
        v = cv.BoundedVoronoi(pts, bb)
        n = cv.partition_doubles(A)
        assert n % 2 == 0                                      # the next partition in a shared buffer stays 16-byte aligned
        buf = np.full(n, np.nan)
        cv.pack_partition(v, buf)
        assert np.all(np.isfinite(buf))
        p = cv.PackedPartition(torch.from_numpy(buf), A, v.seeds_inside)
        seeds, poly, off = v.flat()
        assert len(p) == A and p.nvert == cv.partition_capacity(A) >= off[-1]
        assert np.array_equal(p.seeds.numpy(), seeds.reshape(-1)) and np.array_equal(p.off.numpy()[:A + 1], off)
        assert np.array_equal(p.poly.numpy()[:2 * off[-1]], poly.reshape(-1)) and np.array_equal(p.areas(), v.areas())
        assert cv.seeds_summary(pts, bb) == (A, v.seeds_inside)
    outside = np.array([[0.5, 0.5], [1.05, 0.5], [3.0, 3.0]])     # one seed in the cushion, one beyond it (dropped)
    v = cv.BoundedVoronoi(outside, bb)
    assert cv.seeds_summary(outside, bb) == (len(v), v.seeds_inside) == (2, False)


def test_host_argmax_merge_matches_the_tensor_version():
    import torch
    from mfgp_coverage_b200 import sharding
    rng = np.random.default_rng(8)
    vals = rng.integers(0, 4, size=(5, 40)).astype(np.float64)          # many ties across ranks
    idxs = rng.permutation(5 * 40).reshape(5, 40).astype(np.int64)
    idxs[rng.random((5, 40)) < 0.3] = -1                                # empty cells on some ranks
    idxs[:, 7] = -1                                                     # ... and one cell empty everywhere
    bv, bi = sharding.merge_argmax_host(vals, idxs)
    tv, ti = sharding.merge_argmax(torch.from_numpy(vals), torch.from_numpy(idxs))
    assert np.array_equal(bv, tv.numpy()) and np.array_equal(bi, ti.numpy())
    assert bi[7] == -1 and bv[7] == -np.inf


def test_cross_rank_argmax_merge_applies_the_kernels_tie_rule():
    """Two mirror-image grid points whose variances differ by rounding noise (~1e-16) and that land on DIFFERENT ranks
    must resolve to the lower global index, as csrc/argmax.cuh does inside a shard; the 1-ulp plateaus of k0 - q far
    from all data must NOT be merged (tolerance relative to the variance reduction)."""
    import torch
    from mfgp_coverage_b200 import sharding
    k0, rel = 1.7, 1e-10
    # (a) symmetric-prior case: rank 1 holds the larger value by one ulp, rank 0 the lower index
    v0 = 0.9
    vals = np.array([[v0, 0.5], [np.nextafter(v0, 2.0), 0.7]])
    idxs = np.array([[10, 11], [900, 901]], dtype=np.int64)
    bv, bi = sharding.merge_argmax_host(vals, idxs, k0, rel)
    assert bi.tolist() == [10, 901] and bv[0] == vals[1, 0] and bv[1] == 0.7       # tie -> low index, larger value carried
    tv, ti = sharding.merge_argmax(torch.from_numpy(vals), torch.from_numpy(idxs), k0, rel)
    assert np.array_equal(bv, tv.numpy()) and np.array_equal(bi, ti.numpy())
    assert sharding.merge_argmax_host(vals, idxs)[1].tolist() == [900, 901]            # plain rule: noise decides
    # (b) plateau case: var = k0 - q with q ~ ulp(k0); one ulp apart is NOT a tie
    top = k0 - 2.0 ** -52
    vals = np.array([[np.nextafter(top, 0.0)], [top]])
    idxs = np.array([[3], [4000]], dtype=np.int64)
    assert sharding.merge_argmax_host(vals, idxs, k0, rel)[1].tolist() == [4000]
    # the fold equals the kernel's pairwise rule on random data with many near-ties and empty cells
    rng = np.random.default_rng(2)
    base = rng.random(64)
    vals = base[None, :] + rng.integers(-2, 3, size=(8, 64)) * 1e-16
    idxs = rng.permutation(8 * 64).reshape(8, 64).astype(np.int64)
    idxs[rng.random((8, 64)) < 0.25] = -1
    bv, bi = sharding.merge_argmax_host(vals, idxs, k0, rel)
    tv, ti = sharding.merge_argmax(torch.from_numpy(vals), torch.from_numpy(idxs), k0, rel)
    assert np.array_equal(bi, ti.numpy()) and np.array_equal(bv, tv.numpy())
    for a in range(64):
        live = idxs[:, a] >= 0
        assert bi[a] == (idxs[live, a].min() if live.any() else -1)                    # all within tolerance: lowest index


def test_dataflow_ticket_order_with_chain_lookahead_cannot_deadlock():
    """The order the kernel uses (chain task d drawn d // 5 block columns early): the same tasks, each once; ONLY chain
    tasks wait on later tickets; and however the resident CTAs are scheduled, at most nb // 5 + 2 of them can hold such a
    task at one time -- every other resident CTA holds a task whose dependencies have smaller tickets, so the smallest
    unfinished ticket among those always makes progress.  Simulated: tickets drawn in order by a small pool of CTAs, a task
    completes when its dependencies have; the run must finish with as few CTAs as (early bound + 1)."""
    from oracle import tiled_cholesky as tc
    for nb, nr in ((1, 0), (2, 1), (7, 2), (20, 3), (64, 2)):
        base = tc.task_order(nb, nr, chain_la=0)
        order = tc.task_order(nb, nr)
        assert sorted(map(str, order)) == sorted(map(str, base))
        early = tc.early_tasks(nb, nr)
        assert all(t[0] == "chain" for t in early)
        bound = nb // 5 + 2
        # worst case over time of chain tasks that are drawn but cannot have finished: chain d is early from its ticket to the
        # ticket of its last dependency
        pos = {t: k for k, t in enumerate(order)}
        live = [0] * (len(order) + 1)
        for t, span in early.items():
            for k in range(pos[t], pos[t] + span):
                live[k] += 1
        assert max(live, default=0) <= bound, (nb, nr, max(live))
        # event simulation with a pool of `bound + 1` CTAs
        ncta = bound + 1
        deps = {t: [tc.producer_of(d) for d in tc.task_dependencies(t)] for t in order}
        done, running, nxt = set(), [], 0
        for _ in range(10 * len(order) + 10):
            while len(running) < ncta and nxt < len(order):
                running.append(order[nxt])
                nxt += 1
            finished = [t for t in running if all(d in done for d in deps[t])]
            if not finished and not running:
                break
            assert finished, (nb, nr, "deadlock with", ncta, "CTAs", running[:4])
            done.update(finished)
            running = [t for t in running if t not in done]
            if len(done) == len(order):
                break
        assert len(done) == len(order)


def test_two_queue_gram_policy_cannot_deadlock_and_keeps_group_order():
    """mfgp_cholesky_solve_gram: the Gram tasks (M = Y^T Y per group of block rows) ride in a second ticket queue of the same
    kernel.  Restated policy under adversarial random interleavings, lost races at group boundaries included, with very few
    CTAs: every task completes (no CTA ever waits on a task nobody has drawn) and every tile of M receives its groups in
    order -- which is what makes M bitwise reproducible."""
    from oracle import tiled_cholesky as tc
    rng = np.random.default_rng(11)
    for nb, nr, mg, lead, ncta in ((4, 1, 2, 0, 2), (9, 2, 2, 3, 3), (16, 3, 4, 8, 4), (16, 2, 8, 2, 5), (24, 2, 4, 4, 7), (12, 3, 16, 1, 3)):
        order = tc.simulate_two_queues(nb, nr, mg, lead, max(ncta, nb // 5 + 3), rng)
        rows = tc.gram_group_rows(nb, mg)
        ngroups = len(rows)
        assert rows[0][0] == 0 and rows[-1][1] == nb and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        assert rows[-1][1] - rows[-1][0] == 1
        assert len(order) == len(set(order)) == nb + (nb - 1) * (nb - 2) // 2 + nb * nr + ngroups * nr * (nr + 1) // 2
        seen = {}
        for t in order:
            if t[0] == "M":
                assert seen.get(t[2:], -1) == t[1] - 1, t
                seen[t[2:]] = t[1]
        assert all(v == ngroups - 1 for v in seen.values())


def test_chebyshev_truncation_table():
    """The truncated column layout of the factored posterior (host side, _engine.chebyshev_truncation): per y term a multiple
    of 4 between 4 and kpad, non-increasing, full width for the leading terms; every dropped term's coefficient bound is
    below the tolerance and their sum stays under the 5e-15 the orders were chosen for; c4's 36 x 36 block shrinks by a third."""
    from mfgp_coverage_b200 import _engine as E
    hull = (-0.05, 1.05)
    for lH, lL, lo, hi in ((0.2, 0.58, 0.0, 1.0), (0.2, 0.58, 0.5, 0.625), (0.35, 0.9, 0.0, 1.0)):
        rH = E.chebyshev_order(lH, lo, hi, *hull)
        rL = E.chebyshev_order(lL, lo, hi, *hull)
        ryH, ryL = E.chebyshev_order(lH, 0.0, 1.0, *hull), E.chebyshev_order(lL, 0.0, 1.0, *hull)
        kp = -(-max(rH, rL) // 4) * 4
        parts = [(np.pad(E.chebyshev_envelope(lH, lo, hi, *hull, rH), (0, kp - rH)), E.chebyshev_envelope(lH, 0.0, 1.0, *hull, ryH)),
                 (np.pad(E.chebyshev_envelope(lL, lo, hi, *hull, rL), (0, kp - rL)), E.chebyshev_envelope(lL, 0.0, 1.0, *hull, ryL))]
        ry = max(ryH, ryL)
        kx = E.chebyshev_truncation(parts, kp, ry)
        assert kx.dtype == np.int32 and kx.size == ry and np.all(kx % 4 == 0) and kx.min() >= 4 and kx.max() <= kp
        assert np.all(np.diff(kx) <= 0) and kx[0] == kp
        dropped = 0.0
        for ax, ay in parts:
            for l in range(ay.size):
                tail = ay[l] * ax[kx[l]:]
                assert tail.size == 0 or tail.max() <= 1e-16
                dropped += float(tail.sum())
        assert dropped <= 5e-15
        if (lH, lo, hi) == (0.2, 0.0, 1.0):
            assert (rH, ryH) == (36, 36) and kx.sum() <= 0.7 * 36 * 36


def test_vectorised_finishing_is_bitwise_the_reference_loop():
    """loss_from_partials / centroids_from_partials are elementwise over the cells; they must give bit for bit what the
    reference's per-cell statements give (simulator.py:215-219, :256-271), empty cells (0/0 -> NaN) included."""
    from mfgp_coverage_b200 import _coverage as cv
    rng = np.random.default_rng(0)
    for _ in range(100):
        A = int(rng.integers(1, 70))
        cent = np.column_stack((rng.normal(0.3, 1.0, A), rng.normal(0, 1, A), rng.normal(0, 1, A),
                                rng.integers(0, 50, A).astype(float)))
        lossp = np.column_stack((rng.random(A), rng.integers(0, 50, A).astype(float)))
        areas = rng.random(A)
        loss = 0
        ref = np.empty((A, 2))
        with np.errstate(invalid="ignore", divide="ignore"):
            for i in range(A):
                loss += (lossp[i, 0] / lossp[i, 1]) * areas[i]                     # mean(point_loss) * area
                n = cent[i, 3]
                f_integral = (cent[i, 0] / n) * areas[i]
                c = (np.array([cent[i, 1] / n, cent[i, 2] / n]) * areas[i]) / f_integral
                c[0] = 0.0 if c[0] < 0.0 else c[0]
                c[0] = 1.0 if c[0] > 1.0 else c[0]
                c[1] = 0.0 if c[1] < 0.0 else c[1]
                c[1] = 1.0 if c[1] > 1.0 else c[1]
                ref[i] = c
        got = cv.loss_from_partials(lossp, areas)
        assert (np.isnan(got) and np.isnan(loss)) or got == loss
        assert np.array_equal(cv.centroids_from_partials(cent, areas, 0.0, 1.0, 0.0, 1.0), ref, equal_nan=True)


def test_pair_step_factorisation_matches_lapack():
    """The algorithm of potrf_diag_body (two columns per step, 2x2 pivot blocks, inverse carried along), restated in numpy:
    same factor and inverse as LAPACK, same failing-pivot index."""
    import scipy.linalg as sl
    from oracle import tiled_cholesky as tc
    rng = np.random.default_rng(4)
    for n, cond in ((2, 10.0), (8, 1e3), (64, 1e6), (64, 1e10)):
        q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        S = (q * np.logspace(0, -np.log10(cond), n)) @ q.T
        S = 0.5 * (S + S.T)
        L, W, info = tc.pair_step_factor(S)
        Lref = np.linalg.cholesky(S)
        assert info == 0
        assert np.max(np.abs(L - Lref)) <= 1e-15 * cond * np.max(np.abs(Lref)) + 1e-15
        assert np.max(np.abs(W @ Lref - np.eye(n))) <= 1e-14 * cond
        if n >= 4:                                         # the chain task's form: two half-size sweeps + three small products
            L2, W2, info2 = tc.two_level_factor(S)
            assert info2 == 0
            assert np.max(np.abs(L2 - Lref)) <= 1e-15 * cond * np.max(np.abs(Lref)) + 1e-15
            assert np.max(np.abs(W2 @ Lref - np.eye(n))) <= 1e-14 * cond
    S = (q * np.logspace(0, -3, 64)) @ q.T
    Lr = np.linalg.cholesky(S)
    for bad in (0, 5, 38, 63):
        Sb = S.copy()
        Sb[bad, bad] = float(np.sum(Lr[bad, :bad] ** 2)) - 0.25
        _, info_ref = sl.lapack.dpotrf(Sb, lower=1)
        L, W, info = tc.pair_step_factor(Sb)
        assert info == info_ref == bad + 1 and np.array_equal(L, np.eye(64)) and np.array_equal(W, np.eye(64))
        assert tc.two_level_factor(Sb)[2] == info_ref


def test_dataflow_ticket_order_is_topological():
    """chol_dataflow_kernel hands out tasks through a ticket counter and a task spins on the ready flags of other tasks; it
    cannot deadlock iff every task only waits for tasks with a SMALLER ticket (a CTA holds a ticket only while resident).
    Checked exhaustively on the restated order for every size up to 20 block columns; the count matches the kernel's."""
    from oracle import tiled_cholesky as tc
    for nb in range(1, 21):
        for nr in (0, 1, 3):
            order = tc.task_order(nb, nr, chain_la=0)     # chain tasks with their own column: strictly topological
            assert len(order) == len(set(order)) == nb + (nb - 1) * (nb - 2) // 2 + nb * nr       # a.total in gp_fit.cu
            ticket = {t: k for k, t in enumerate(order)}
            produced = set()
            for t in order:
                for dep in tc.task_dependencies(t):
                    assert ticket[tc.producer_of(dep)] < ticket[t], (nb, nr, t, dep)
                    assert dep in produced, (nb, nr, t, dep)
                if t[0] == "chain":
                    produced.add(("W", t[1]))
                    if t[1] > 0:
                        produced.add(("Ltile", t[1], t[1] - 1))
                elif t[0] == "L":
                    produced.add(("Ltile", t[1], t[2]))
                else:
                    produced.add(("Ytile", t[1], t[2]))
            # every tile of L below the diagonal, every diagonal block and every Y tile is produced exactly once
            assert produced == {("W", c) for c in range(nb)} | {("Ltile", i, c) for c in range(nb) for i in range(c + 1, nb)} | \
                {("Ytile", c, r) for c in range(nb) for r in range(nr)}


# ---- runner._map_sims: the multi-process run-sharding path (reference runner.py:131-141) ------------------------------

def _sim_ok(a):
    return ([{"sim": a}], [], [])


def _sim_raises(a):
    if a == 3:
        raise np.linalg.LinAlgError("Matrix is not positive definite (pivot 7)")
    return ([{"sim": a}], [], [])


def _sim_dies(a):
    import os
    if a == 2:
        os._exit(17)              # hard crash: no report ever reaches the queue
    return ([{"sim": a}], [], [])


def test_map_sims_shards_runs_and_propagates_worker_failures():
    from mfgp_coverage_b200 import runner
    out = runner._map_sims(list(range(7)), 3, seed=1, fn=_sim_ok)
    assert [o[0][0]["sim"] for o in out] == list(range(7))                 # original order, whatever worker ran what
    with pytest.raises(np.linalg.LinAlgError, match="pivot 7"):            # Pool.map re-raises worker exceptions
        runner._map_sims(list(range(7)), 3, fn=_sim_raises)
    with pytest.raises(runner.WorkerError, match="exited with code 17"):   # ... and a dead worker must not hang the parent
        runner._map_sims(list(range(7)), 3, fn=_sim_dies)
    with pytest.raises(ValueError, match="Invalid simulation algorithm"):  # the real run_sim, failing before any GPU call
        runner._map_sims([("o", "nonsense", 0, 1, 2, None, 0.1, None, None, False, None, False)] * 2, 2)


def test_bordered_inverse_append_equals_a_refit():
    """oracle.tiled_cholesky.bordered_inverse_append (the algorithm of the batched stepper's append kernel) against the
    reference's refit-from-scratch: same W = L^-1 and z after every append, on an MF model that grows by 1..8 hifi samples."""
    from oracle import tiled_cholesky as tc
    xy = synth.grid(24)
    f = synth.truth_function(xy)
    X_L, y_L, X_H, y_H = synth.training_set(xy, f, 120)
    p = ogp.GPParams.from_hyp(synth.MF_HYP)
    n0 = 20
    K = ogp.train_cov(p, X_L, X_H[:n0])
    W = np.linalg.inv(np.linalg.cholesky(K))
    yc = ogp.centered_y(p, y_L, y_H[:n0]).reshape(-1)
    z = W @ yc
    nh = n0
    for q in (1, 8, 3, 8, 5):
        Kfull = ogp.train_cov(p, X_L, X_H[:nh + q])
        N = K.shape[0]
        ycn = ogp.centered_y(p, y_L, y_H[:nh + q]).reshape(-1)
        W, z = tc.bordered_inverse_append(W, z, Kfull[:N, N:], Kfull[N:, N:], ycn)
        Lref = np.linalg.cholesky(Kfull)
        Wref = np.linalg.inv(Lref)
        assert np.max(np.abs(W - Wref)) <= 1e-11 * np.max(np.abs(Wref))
        assert np.max(np.abs(z - Wref @ ycn)) <= 1e-11 * max(1.0, np.max(np.abs(z)))
        assert np.allclose(np.triu(W, 1), 0.0)
        K, nh = Kfull, nh + q
