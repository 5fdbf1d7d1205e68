// Device helpers shared by the coverage kernels (coverage.cu) and the batched replicate stepper (batched.cu): the
// reference's crossings test and the warp-per-cell bounded-Voronoi clipper.
#pragma once
#include "common.cuh"

namespace mfgp {

// matplotlib _path.h point_in_path_impl, radius 0, no codes (implicitly closed polygon): SURVEY.md Appendix A.1
__device__ __forceinline__ bool crossings_inside(const double* __restrict__ pv, int n, double tx, double ty) {
    if (n < 3) return false;
    bool inside = false;
    double x0 = pv[2 * (n - 1)], y0 = pv[2 * (n - 1) + 1];   // closing edge v_{n-1} -> v_0 first; toggles commute
    bool f0 = y0 >= ty;
    for (int i = 0; i < n; i++) {
        const double x1 = pv[2 * i], y1 = pv[2 * i + 1];
        const bool f1 = y1 >= ty;
        if (f0 != f1) {
            const double lhs = __dmul_rn(__dsub_rn(y1, ty), __dsub_rn(x0, x1));
            const double rhs = __dmul_rn(__dsub_rn(x1, tx), __dsub_rn(y0, y1));
            if ((lhs >= rhs) == f1) inside = !inside;
        }
        f0 = f1; x0 = x1; y0 = y1;
    }
    return inside;
}

constexpr int VC_MAXV = 48;
constexpr int VC_WARPS = 16;

// One WARP per cell, lane = polygon vertex (two slots per lane: up to 64 >= VC_MAXV vertices).  A clip against one bisector
// is: signed distances of every vertex and of its predecessor in parallel, two ballots (vertex kept / edge crosses), output
// slots by prefix popcount -- the same vertices, from the same fp operations, in the same order as the sequential
// Sutherland-Hodgman sweep (crossing point first, then the kept vertex).  Most bisectors do not touch the cell once its
// nearest neighbours have been applied: one ballot rejects them.  Pass 1 counts the vertices (-> packed offsets), pass 2
// clips again and writes vertices + shoelace area (a cell costs a few thousand cycles; a stash would not pay).
__device__ __forceinline__ int vc_clip_cell(const double* __restrict__ s_seeds, int A, int i, double x0, double x1, double y0,
                                            double y1, double* __restrict__ buf, int lane, bool& overflow, int& which) {
    // buf: [2 ping-pong][2 coords][64]
    double* P = buf;
    double* Q = buf + 128;
    if (lane < 4) {
        P[lane] = (lane == 0 || lane == 3) ? x0 : x1;
        P[64 + lane] = (lane < 2) ? y0 : y1;
    }
    __syncwarp();
    int n = 4;
    which = 0;
    const double sx = s_seeds[2 * i], sy = s_seeds[2 * i + 1];
    const unsigned lt = (1u << lane) - 1u;
    // r2max: squared distance of the farthest polygon vertex from the seed.  The bisector with seed j passes at distance
    // |s_j - s_i| / 2 from s_i, so it cannot cut the polygon when |s_j - s_i|^2 > 4 r2max: those seeds are skipped by a
    // scalar test (exact: the clip would have left the polygon unchanged).  After the first few neighbours r2max is a few
    // cell radii and most seeds fall out here.
    double r2max = 0.0;
    {
        const double ex = fmax(fabs(x0 - sx), fabs(x1 - sx)), ey = fmax(fabs(y0 - sy), fabs(y1 - sy));
        r2max = ex * ex + ey * ey;
    }
    for (int j = 0; j < A && n > 0; j++) {
        if (j == i) continue;
        const double tx = s_seeds[2 * j], ty = s_seeds[2 * j + 1];
        const double nx = tx - sx, ny = ty - sy;
        if (nx == 0.0 && ny == 0.0) continue;                      // coincident seeds share one cell
        if (nx * nx + ny * ny > 4.0 * r2max * (1.0 + 1e-9)) continue;
        const double c = 0.5 * ((tx * tx + ty * ty) - (sx * sx + sy * sy));
        double bx[2], by[2], ax[2], ay[2], da[2], db[2];
        bool valid[2], inb[2], cross[2];
        bool any_out = false;
#pragma unroll
        for (int sl = 0; sl < 2; sl++) {
            const int v = lane + 32 * sl;
            valid[sl] = v < n;
            const int u = valid[sl] ? (v + n - 1) % n : 0, vv = valid[sl] ? v : 0;
            bx[sl] = P[vv]; by[sl] = P[64 + vv]; ax[sl] = P[u]; ay[sl] = P[64 + u];
            db[sl] = nx * bx[sl] + ny * by[sl] - c;
            da[sl] = nx * ax[sl] + ny * ay[sl] - c;
            inb[sl] = valid[sl] && db[sl] <= 0.0;
            cross[sl] = valid[sl] && ((da[sl] <= 0.0) != (db[sl] <= 0.0));
            any_out = any_out || (valid[sl] && !(db[sl] <= 0.0));
        }
        if (!__any_sync(0xffffffffu, any_out)) continue;           // the whole polygon lies on the near side: unchanged
        const unsigned c0 = __ballot_sync(0xffffffffu, cross[0]), i0 = __ballot_sync(0xffffffffu, inb[0]);
        const unsigned c1 = __ballot_sync(0xffffffffu, cross[1]), i1 = __ballot_sync(0xffffffffu, inb[1]);
        const int tot0 = __popc(c0) + __popc(i0);
        const int m = tot0 + __popc(c1) + __popc(i1);
        int pos[2] = {__popc(c0 & lt) + __popc(i0 & lt), tot0 + __popc(c1 & lt) + __popc(i1 & lt)};
#pragma unroll
        for (int sl = 0; sl < 2; sl++) {
            int o = pos[sl];
            if (cross[sl]) {
                const double t = da[sl] / (da[sl] - db[sl]);
                if (o < VC_MAXV) { Q[o] = ax[sl] + (bx[sl] - ax[sl]) * t; Q[64 + o] = ay[sl] + (by[sl] - ay[sl]) * t; }
                o++;
            }
            if (inb[sl] && o < VC_MAXV) { Q[o] = bx[sl]; Q[64 + o] = by[sl]; }
        }
        n = m;
        if (n > VC_MAXV) { overflow = true; n = VC_MAXV; }
        double* T = P; P = Q; Q = T;
        which ^= 1;
        __syncwarp();
        double r2 = 0.0;                                           // new extent of the polygon around its seed
#pragma unroll
        for (int sl = 0; sl < 2; sl++) {
            const int v = lane + 32 * sl;
            if (v < n) { const double dx = P[v] - sx, dy = P[64 + v] - sy; r2 = fmax(r2, dx * dx + dy * dy); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, o));
        r2max = r2;
    }
    return n;
}


}  // namespace mfgp
