"""CPU restatement of matplotlib `src/_path.h::point_in_path_impl` for radius 0 and no path codes (the routine behind
`simulator.py:123-124`); SURVEY.md Appendix A.1.  Pure fp64 subtract / multiply / compare."""
import numpy as np


class Path:
    def __init__(self, vertices):
        self.vertices = np.asarray(vertices, dtype=np.float64).reshape(-1, 2)

    def contains_points(self, points):
        pts = np.asarray(points, dtype=np.float64).reshape(-1, 2)
        v = self.vertices
        n = v.shape[0]
        inside = np.zeros(pts.shape[0], dtype=bool)
        if n < 3:
            return inside
        tx, ty = pts[:, 0], pts[:, 1]
        # edges v_i -> v_{i+1}, closing edge included (the degenerate first pass v0->v0 never toggles)
        for i in range(n):
            x0, y0 = v[i]
            x1, y1 = v[(i + 1) % n]
            f0 = y0 >= ty
            f1 = y1 >= ty
            cross = ((y1 - ty) * (x0 - x1) >= (x1 - tx) * (y0 - y1)) == f1
            inside ^= (f0 != f1) & cross
        return inside
