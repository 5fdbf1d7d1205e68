"""Oracle: the coverage step (numpy/scipy restatement; TEST INFRASTRUCTURE).

Follows /root/reference/simulator.py:
  in_polygon              <- :105-124  (matplotlib Path.contains_points, radius 0 == the crossings test restated in
                                        SURVEY.md Appendix A.1; here vectorised over the query points)
  poly_area               <- :127-136
  in_box, voronoi_bounded <- :139-191  (scipy.spatial.Voronoi == Qhull, mirrored seeds, first A regions)
  compute_loss            <- :194-228
  compute_centroids       <- :231-283
  compute_max_var         <- :286-323  (takes the variance vector, i.e. np.diag of the reference's covariance)
  compute_sample_points   <- :326-374  (greedy planner; `_fast` is the bordered/V-cached restatement of SURVEY.md
                                        section 7 step 5, numerically equivalent, used where the slow one takes minutes)
  compute_sample_clusters <- :377-412
A grid point may lie in 0, 1 or 2+ cells (exact ties); every reduction uses whatever the crossings test returns.
"""
import numpy as np
from scipy.linalg import solve_triangular
from scipy.spatial import Voronoi

from . import gp as ogp

EPS = 0.1   # simulator.py:33


def in_polygon(xq, yq, xv, yv):
    xq = np.asarray(xq, dtype=np.float64).reshape(-1)
    yq = np.asarray(yq, dtype=np.float64).reshape(-1)
    xv = np.asarray(xv, dtype=np.float64).reshape(-1)
    yv = np.asarray(yv, dtype=np.float64).reshape(-1)
    n = xv.shape[0]
    inside = np.zeros(xq.shape[0], dtype=bool)
    if n < 3:
        return inside
    for i in range(n):          # edges v_i -> v_{i+1}; the last one is the implicit closing edge
        x0, y0 = xv[i], yv[i]
        x1, y1 = xv[(i + 1) % n], yv[(i + 1) % n]
        f0 = y0 >= yq
        f1 = y1 >= yq
        cross = ((y1 - yq) * (x0 - x1) >= (x1 - xq) * (y0 - y1)) == f1
        inside ^= (f0 != f1) & cross
    return inside


def poly_area(x, y):
    return 0.5 * np.abs(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1)))


def in_box(points, bounding_box):
    return np.logical_and(np.logical_and(bounding_box[0] - EPS <= points[:, 0], points[:, 0] <= bounding_box[1] + EPS),
                          np.logical_and(bounding_box[2] - EPS <= points[:, 1], points[:, 1] <= bounding_box[3] + EPS))


class BoundedVoronoi:
    """What the reference keeps of the scipy object: vertices, the A bounded regions (vertex-id lists in Qhull's
    order) and the seeds (`filtered_points`)."""

    def __init__(self, vertices, regions, seeds):
        self.vertices = vertices
        self.filtered_regions = regions
        self.filtered_points = seeds

    def cell_vertices(self, i):
        return self.vertices[self.filtered_regions[i], :]


def voronoi_bounded(points, bounding_box):
    points = np.asarray(points, dtype=np.float64)
    i = in_box(points, bounding_box)
    c = points[i, :]
    left = np.copy(c)
    left[:, 0] = bounding_box[0] - (left[:, 0] - bounding_box[0] + EPS)
    right = np.copy(c)
    right[:, 0] = bounding_box[1] + (bounding_box[1] - right[:, 0] + EPS)
    down = np.copy(c)
    down[:, 1] = bounding_box[2] - (down[:, 1] - bounding_box[2] + EPS)
    up = np.copy(c)
    up[:, 1] = bounding_box[3] + (bounding_box[3] - up[:, 1] + EPS)
    pts = np.append(c, np.append(np.append(left, right, axis=0), np.append(down, up, axis=0), axis=0), axis=0)
    vor = Voronoi(pts)
    regions = [list(vor.regions[r]) for r in vor.point_region[:vor.npoints // 5]]
    return BoundedVoronoi(vor.vertices, regions, c)


def bounding_box_of(x_star):
    return np.array([np.amin(x_star[:, 0]), np.amax(x_star[:, 0]), np.amin(x_star[:, 1]), np.amax(x_star[:, 1])])


def membership(vor, x_star):
    """[A, G] boolean: crossings test of every grid point against every cell."""
    return np.stack([in_polygon(x_star[:, 0], x_star[:, 1], vor.cell_vertices(i)[:, 0], vor.cell_vertices(i)[:, 1])
                     for i in range(len(vor.filtered_regions))])


def compute_loss(vor, truth_arr):
    loss = 0
    for i in range(len(vor.filtered_regions)):
        vertices = vor.cell_vertices(i)
        inn = in_polygon(truth_arr[:, 0], truth_arr[:, 1], vertices[:, 0], vertices[:, 1])
        pts = truth_arr[inn, :]
        center = vor.filtered_points[i, :]
        distances = np.sum((pts[:, [0, 1]] - center) ** 2, axis=1)
        point_loss = distances * pts[:, 2]
        with np.errstate(invalid="ignore"), _quiet():
            loss += np.mean(point_loss) * poly_area(vertices[:, 0], vertices[:, 1])
    return loss


def compute_centroids(vor, x_star, mu_star):
    mu_star = np.asarray(mu_star, dtype=np.float64).reshape(-1, 1)
    xmin, xmax = np.amin(x_star[:, 0]), np.amax(x_star[:, 0])
    ymin, ymax = np.amin(x_star[:, 1]), np.amax(x_star[:, 1])
    out = np.empty((0, 2))
    for i in range(len(vor.filtered_regions)):
        vertices = vor.cell_vertices(i)
        inn = in_polygon(x_star[:, 0], x_star[:, 1], vertices[:, 0], vertices[:, 1])
        pts = x_star[inn, :]
        w = mu_star[inn, :]
        area = poly_area(vertices[:, 0], vertices[:, 1])
        with np.errstate(invalid="ignore", divide="ignore"), _quiet():
            f_integral = np.mean(w[:, 0]) * area
            weighted = np.multiply(np.column_stack((w[:, 0], w[:, 0])), pts[:, [0, 1]])
            w_integral = np.mean(weighted, axis=0) * area
            c = w_integral / f_integral
        if c[0] < xmin:
            c[0] = xmin
        if c[0] > xmax:
            c[0] = xmax
        if c[1] < ymin:
            c[1] = ymin
        if c[1] > ymax:
            c[1] = ymax
        out = np.vstack((out, c))
    return out


def compute_max_var(vor, truth_arr, var):
    """Returns (argmax_xy [A,2], max_var [A,1], argmax_index [A]).  Empty cell -> ValueError like np.amax([])."""
    var = np.asarray(var, dtype=np.float64).reshape(-1)
    arg_xy = np.empty((0, 2))
    mx = np.empty((0, 1))
    idx = []
    for i in range(len(vor.filtered_regions)):
        vertices = vor.cell_vertices(i)
        inn = in_polygon(truth_arr[:, 0], truth_arr[:, 1], vertices[:, 0], vertices[:, 1])
        ids = np.nonzero(inn)[0]
        in_var = var[inn]
        m = np.amax(in_var)
        j = int(np.argmax(in_var))
        arg_xy = np.vstack((arg_xy, truth_arr[ids[j], [0, 1]]))
        mx = np.vstack((mx, m))
        idx.append(int(ids[j]))
    return arg_xy, mx, np.asarray(idx, dtype=np.int64)


def compute_sample_points(model, x_star, threshold, max_points=None):
    """The reference's greedy planner, literally: refit + full predict per pick.  Returns (points [k,2], indices [k])."""
    tmp = model.copy()
    mu, var = tmp.predict(x_star)
    pts = np.empty((0, 2))
    idx = []
    while np.amax(var) > threshold:
        j = int(np.argmax(var))
        pts = np.vstack((pts, x_star[j].reshape(1, -1)))
        idx.append(j)
        tmp.append(x_star[j].reshape(1, -1), np.array(mu[j]).reshape(1, -1))
        mu, var = tmp.predict(x_star)
        if max_points is not None and len(idx) >= max_points:
            break
    return pts, np.asarray(idx, dtype=np.int64)


def compute_sample_points_fast(model, x_star, threshold, max_points=None):
    """Same selection by a bordered Cholesky append on the cached V = L^-1 Psi^T:
    new row v = (k_HH(x*, x_j) - l^T V)/d with l = V[:, j], d = sqrt(k(0) + noise_H + jitter - l.l); var -= v^2.
    The posterior mean is unchanged because the pseudo-observation equals the current mean."""
    p = model.p
    x_star = np.asarray(x_star, dtype=np.float64)
    G = x_star.shape[0]
    if model.N:
        psi = ogp.cross_cov(p, x_star, model.X_L, model.X_H)
        V = solve_triangular(model.L, psi.T, lower=True)
        var = p.k0 - np.einsum("ij,ij->j", V, V)
    else:
        V = np.empty((0, G))
        var = np.full(G, p.k0)
    rows = [V]
    pts = np.empty((0, 2))
    idx = []
    while np.amax(var) > threshold:
        j = int(np.argmax(var))
        Vall = np.vstack(rows) if len(rows) > 1 else rows[0]
        rows = [Vall]
        l = Vall[:, j]
        xj = x_star[j].reshape(1, 2)
        if p.multi:
            kx = p.rho ** 2 * ogp.rbf(xj, x_star, p.s_L, p.l_L) + ogp.rbf(xj, x_star, p.s_H, p.l_H)
        else:
            kx = ogp.rbf(xj, x_star, p.s_H, p.l_H)
        d = np.sqrt(p.k0 + p.noise_H + ogp.JITTER - l @ l)
        v = (kx[0] - l @ Vall) / d
        rows.append(v.reshape(1, -1))
        var = var - v * v
        pts = np.vstack((pts, xj))
        idx.append(j)
        if max_points is not None and len(idx) >= max_points:
            break
    return pts, np.asarray(idx, dtype=np.int64)


def compute_sample_clusters(vor, sample_points):
    clusters = [np.empty((0, 2)) for _ in range(len(vor.filtered_regions))]
    if sample_points.shape[0] == 0:
        return clusters
    for i in range(len(vor.filtered_regions)):
        vertices = vor.cell_vertices(i)
        inn = in_polygon(sample_points[:, 0], sample_points[:, 1], vertices[:, 0], vertices[:, 1])
        clusters[i] = sample_points[inn, :]
    return clusters


class _quiet:
    """np.mean([]) warns 'Mean of empty slice'; the reference lets it through (NaN result)."""

    def __enter__(self):
        import warnings
        self._c = warnings.catch_warnings()
        self._c.__enter__()
        warnings.simplefilter("ignore")

    def __exit__(self, *a):
        self._c.__exit__(*a)
