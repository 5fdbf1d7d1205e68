"""Oracle: the deterministic tour planner that stands in for the reference's mlrose genetic algorithm (TEST INFRASTRUCTURE).

The reference orders every agent's sample points with `mlrose.TSPOpt` + `mlrose.genetic_alg(problem, mutation_prob=0.2,
max_attempts=100, random_state=2)` (/root/reference/simulator.py:415-454, call at :435-438).  mlrose is a third-party
dependency that is neither vendored under /root/reference nor pinned (no requirements file) nor installed here, and a
genetic algorithm's tour depends on its implementation's RNG consumption -- the tour ORDER is "parity unpinned"
(SURVEY.md section 8c).  As SURVEY 8(c) item 3 prescribes, both sides therefore use one deterministic replacement with
the same objective mlrose minimises (length of the CLOSED tour through the cluster, `TravellingSales` fitness):

  1. nearest-neighbour construction from point 0 of the cluster (clusters keep the greedy selection order of
     compute_sample_points, so point 0 is the cluster's highest-variance pick); ties -> lowest index;
  2. best-improvement 2-opt with position 0 fixed: among all segment reversals tour[i..j], 1 <= i < j <= n-1, apply the
     one with the most negative length change (ties -> lowest i, then lowest j) until no reversal improves the tour by
     more than IMPROVE_TOL.

This file is the literal, loop-level statement; `choi_tsp_tours` in mfgp-coverage_b200/csrc/tsp.cu is the device
implementation (one CTA per cluster, all reversals evaluated in parallel), and oracle/refshim/mlrose routes the LIVE
reference's `genetic_alg` call here, so seeded reference runs, the oracle loops and the product follow identical tours.
All arithmetic is IEEE fp64 with one rounding per operation (no fused multiply-add), in the order written below, so the
decisions are bit-reproducible on the device.
"""
import math

IMPROVE_TOL = 1e-12


def _dist(p, a, b):
    dx = p[a][0] - p[b][0]
    dy = p[a][1] - p[b][1]
    return math.sqrt(dx * dx + dy * dy)


def nearest_neighbour(points):
    p = [(float(x), float(y)) for x, y in points]
    n = len(p)
    tour, used = [0], [False] * n
    used[0] = True
    for _ in range(n - 1):
        cur = tour[-1]
        best, best_d2 = -1, math.inf
        for k in range(n):
            if used[k]:
                continue
            dx = p[cur][0] - p[k][0]
            dy = p[cur][1] - p[k][1]
            d2 = dx * dx + dy * dy
            if d2 < best_d2:              # strict: the lowest index wins ties
                best, best_d2 = k, d2
        tour.append(best)
        used[best] = True
    return tour


def two_opt(points, tour, max_moves=None):
    p = [(float(x), float(y)) for x, y in points]
    n = len(tour)
    tour = list(tour)
    if n < 4:
        return tour, 0
    max_moves = 20 * n + 100 if max_moves is None else max_moves
    moves = 0
    while moves < max_moves:
        e = [_dist(p, tour[k], tour[(k + 1) % n]) for k in range(n)]      # e[k]: edge tour[k] -> tour[k+1]
        best = (-IMPROVE_TOL, -1, -1)
        for i in range(1, n - 1):
            a, b = tour[i - 1], tour[i]
            for j in range(i + 1, n):
                c, d = tour[j], tour[(j + 1) % n]
                delta = (_dist(p, a, c) + _dist(p, b, d)) - (e[i - 1] + e[j])
                if delta < best[0]:       # strict: lowest (i, j) wins ties
                    best = (delta, i, j)
        if best[1] < 0:
            break
        _, i, j = best
        tour[i:j + 1] = tour[i:j + 1][::-1]
        moves += 1
    return tour, moves


def plan_tour(points):
    """Visiting order (indices into `points`) of the closed tour; `points`: sequence of (x, y)."""
    n = len(points)
    if n == 0:
        return []
    tour = nearest_neighbour(points)
    tour, _ = two_opt(points, tour)
    return tour


def tour_length(points, tour):
    p = [(float(x), float(y)) for x, y in points]
    n = len(tour)
    return sum(_dist(p, tour[k], tour[(k + 1) % n]) for k in range(n))


def compute_sample_tsp(clusters):
    """/root/reference/simulator.py:415-454 with the planner above in place of mlrose."""
    import numpy as np
    tours = []
    for cluster in clusters:
        tour = np.empty((0, 2))
        if cluster.shape[0] > 0:
            tour = cluster[plan_tour([tuple(c) for c in cluster])]
        tours.append(tour)
    return tours
