// Batched replicate stepper: many INDEPENDENT coverage runs on small grids advance one iteration per launch sequence.
// Replaces, for the replicate sweeps of the reference (runner.py:131-147: Pool.map over 100+ simulations of the same
// experiment; BASELINE config 5: 512 periodic_hmf runs on 64x64 grids), the per-run loop body of simulator.py
// periodic :618-785 / todescato :788-954 / lloyd :508-616:
//     take samples (:698-713)  ->  updt_hifi / updt (:716-721)  ->  predict (:723)  ->  compute_loss, compute_centroids,
//     compute_max_var (:725-733)  ->  log rows (:740-768)  ->  decision (:771-775)  ->  move (:777-782).
// One run at a time this loop is launch-bound (~20 small kernels + host Python per iteration, ~0.5 ms); here THREE launches
// step every run of the batch, the loop state never leaves the device, and the host is not involved between iterations
// (the three launches of an iteration can be replayed from a CUDA graph):
//   batch_append_kernel     one CTA per run: the exploring agents' samples (truth at the position's grid index + the run's
//                           pre-drawn noise), block-bordered update of W = L^-1 and z (the reference refits from scratch,
//                           gaussian_process.py:266-268, :540-542; appended rows only change the trailing block), axis-factor
//                           table rows of the new points;
//   batch_posterior_kernel  the standing posterior takes the new rows' contribution: mu += v_new . z_new, var -= |v_new|^2,
//                           v_new = W_new psi with the separable cross-covariance read from the per-run axis tables;
//   batch_coverage_kernel   one CTA per run: both bounded-Voronoi partitions clipped in place (cov_device.cuh), membership
//                           of every grid point (nearest seed; the reference's crossings test within TIE_TOL of a bisector),
//                           per-cell sums / arg-max with the tie rule of argmax.cuh in a fixed order, O(A) finishing with the
//                           reference's arithmetic, log rows, explore decision, position update.
// Randomness is drawn on the host in the order the reference consumes it and uploaded once (uniforms of todescato's
// Bernoulli draws :943, the N(0, sigma_n) sample noise :707).  Dynamics (centroids, arg-max points, decisions, samples) do
// not depend on how exact bisector ties are resolved (the Lloyd partition is seeded by centroids); the LOSS of an iteration
// whose agents sit on grid points may hold grid points exactly on a bisector, which the reference resolves through Qhull's
// vertex rounding -- those (run, iteration) pairs are counted in `ties` for the host to re-evaluate if it needs them bit-exact.
#include <cfloat>

#include "common.cuh"
#include "argmax.cuh"
#include "cov_device.cuh"

namespace mfgp {

constexpr int BT_THREADS = 256;
constexpr int BT_MAXA = 16;          // agents per run
constexpr int BT_LOGC = 11;          // X, Y, XMax, YMax, VarMax, Var0, XCentroid, YCentroid, ProbExplore, Explore, Distance

struct BatchArgs {
    int runs, G, nx, ny, A, NL, cap, algo, iterations, max_samples;      // algo: 0 lloyd, 1 periodic, 2 todescato
    double xmin, xmax, ymin, ymax, eps, tie_tol, amax_rel;
    DevParams p;
    const double* xy; const double* f; const double* ux; const double* uy;
    double* Xt; double* y; double* W; double* z;                         // [runs][cap(,2 | ,cap)]
    double* TxL; double* TyL; double* TxH; double* TyH;                  // [runs][cap][nx | ny]
    double* mu; double* var;                                             // [runs][G]
    double* pos; double* prev; double* cen;                              // [runs][A][2]
    long long* pos_idx;                                                  // [runs][A] grid index of the position, or -1
    double* prob; int* explore;                                          // [runs][A]
    int* Ncur; int* knew; int* status; int* noise_used; int* nsamples;   // [runs]
    int* ties;                                                           // [runs][iterations]
    const double* noise; const double* unif;                             // [runs][max_samples], [runs][iterations][A]
    double* log_loss; double* log_agent; double* log_sample;             // [runs][it], [runs][it][A][BT_LOGC], [runs][max_samples][5]
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// training covariance entry between a NEW hifi point (x, y) and training point n of the run -- the operations of
// build_train_cov_kernel (gp_fit.cu), i.e. of gaussian_process.py:523-528 / :253
__device__ __forceinline__ double bt_cov(const DevParams& p, double xi, double yi, double xj, double yj, bool jL) {
    if (p.multi) {
        const double kL = rbf_scaled(xi / p.l_L, yi / p.l_L, xj / p.l_L, yj / p.l_L, p.s_L);
        if (jL) return p.rho * kL;
        const double kH = rbf_scaled(xi / p.l_H, yi / p.l_H, xj / p.l_H, yj / p.l_H, p.s_H);
        return __dadd_rn(__dmul_rn(p.rho2, kL), kH);
    }
    return rbf_scaled(xi / p.l_H, yi / p.l_H, xj / p.l_H, yj / p.l_H, p.s_H);
}

// ---- step 1: samples + block-bordered update of W = L^-1 ---------------------------------------------------------------------
// With K_new = [[K, k], [k^T, kk]], l = W k (N x q), S = kk - l^T l = C C^T:   L_new = [[L, 0], [l^T, C]],
// W_new = [[W, 0], [-C^-1 l^T W, C^-1]],  z_new = W_new,rows (y - m).  q <= A new points per iteration.
__global__ void __launch_bounds__(BT_THREADS) batch_append_kernel(BatchArgs a, int it) {
    extern __shared__ __align__(16) double bsm[];
    __shared__ int s_agent[BT_MAXA];
    __shared__ int s_q, s_N;
    __shared__ double s_S[BT_MAXA][BT_MAXA], s_Ci[BT_MAXA][BT_MAXA], s_xn[BT_MAXA][2];
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = BT_THREADS / 32;
    const int cap = a.cap, A = a.A;
    if (tid == 0) {
        int q = 0;
        if (a.status[r] == 0 && a.algo != 0)
            for (int i = 0; i < A; i++)
                if (a.explore[r * A + i]) s_agent[q++] = i;
        s_N = a.Ncur[r];
        if (q > 0 && (s_N + q > cap || a.noise_used[r] + q > a.max_samples)) { a.status[r] = -1; q = 0; }
        s_q = q;
        a.knew[r] = q;
    }
    __syncthreads();
    const int q = s_q, N = s_N;
    if (q == 0) return;
    double* Xt = a.Xt + (int64_t)r * cap * 2;
    double* yv = a.y + (int64_t)r * cap;
    double* W = a.W + (int64_t)r * cap * cap;
    double* z = a.z + (int64_t)r * cap;
    double* kv = bsm;                       // [q][cap]  covariance of the new points with the old ones
    double* lv = bsm + (int64_t)A * cap;            // [q][cap]  l = W k
    if (tid < q) {          // reference simulator.py:698-713: sample = truth at the position's grid point + N(0, sigma_n)
        const int ag = s_agent[tid];
        const long long gi = a.pos_idx[r * A + ag];
        const int used = a.noise_used[r];
        const double sx = a.xy[2 * gi], sy = a.xy[2 * gi + 1];
        const double sample = a.f[gi] + a.noise[(int64_t)r * a.max_samples + used + tid];
        Xt[2 * (N + tid)] = sx; Xt[2 * (N + tid) + 1] = sy;
        yv[N + tid] = sample;
        s_xn[tid][0] = sx; s_xn[tid][1] = sy;
        double* row = a.log_sample + ((int64_t)r * a.max_samples + used + tid) * 5;
        row[0] = (double)it; row[1] = (double)ag; row[2] = sx; row[3] = sy; row[4] = sample;
    }
    __syncthreads();
    if (tid == 0) { a.noise_used[r] += q; a.nsamples[r] += q; }
    for (int e = tid; e < q * N; e += BT_THREADS) {
        const int j = e / N, n = e % N;
        kv[j * cap + n] = bt_cov(a.p, s_xn[j][0], s_xn[j][1], Xt[2 * n], Xt[2 * n + 1], n < a.NL);
    }
    if (tid < q * q) {
        const int j = tid / q, j2 = tid % q;
        double v = a.p.multi ? __dadd_rn(__dmul_rn(a.p.rho2, rbf_scaled(s_xn[j][0] / a.p.l_L, s_xn[j][1] / a.p.l_L, s_xn[j2][0] / a.p.l_L,
                                                                        s_xn[j2][1] / a.p.l_L, a.p.s_L)),
                                         rbf_scaled(s_xn[j][0] / a.p.l_H, s_xn[j][1] / a.p.l_H, s_xn[j2][0] / a.p.l_H, s_xn[j2][1] / a.p.l_H, a.p.s_H))
                                : rbf_scaled(s_xn[j][0] / a.p.l_H, s_xn[j][1] / a.p.l_H, s_xn[j2][0] / a.p.l_H, s_xn[j2][1] / a.p.l_H, a.p.s_H);
        if (j == j2) { v = v + a.p.noise_H; v = v + a.p.jitter; }
        s_S[j][j2] = v;
    }
    __syncthreads();
    // l[n][j] = sum_{m <= n} W[n][m] k[m][j]: a warp per row, lanes over m (coalesced), fixed butterfly
    for (int n = warp; n < N; n += nwarp) {
        double acc[BT_MAXA];
#pragma unroll
        for (int j = 0; j < BT_MAXA; j++) acc[j] = 0.0;
        const double* wr = W + (int64_t)n * cap;
#pragma unroll 4
        for (int m = lane; m <= n; m += 32) {         // (unrolled: four loads of W in flight per lane; same order of the sums)
            const double w = wr[m];
#pragma unroll
            for (int j = 0; j < BT_MAXA; j++)
                if (j < q) acc[j] = fma(w, kv[j * cap + m], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < BT_MAXA; j++)
            if (j < q) {
                const double s = warp_sum_d(acc[j]);
                if (lane == 0) lv[j * cap + n] = s;
            }
    }
    __syncthreads();
    // S = kk - l^T l
    for (int e = warp; e < q * q; e += nwarp) {
        const int j = e / q, j2 = e % q;
        double s = 0.0;
        for (int n = lane; n < N; n += 32) s = fma(lv[j * cap + n], lv[j2 * cap + n], s);
        s = warp_sum_d(s);
        if (lane == 0) s_S[j][j2] -= s;
    }
    __syncthreads();
    if (tid == 0) {          // C = chol(S) (lower), Ci = C^-1; np.linalg.cholesky raises on a non-positive pivot (:254, :529)
        int bad = 0;
        for (int j = 0; j < q && !bad; j++) {
            for (int j2 = 0; j2 <= j; j2++) {
                double s = s_S[j][j2];
                for (int t = 0; t < j2; t++) s -= s_S[j][t] * s_S[j2][t];
                if (j == j2) {
                    if (!(s > 0.0)) { bad = N + j + 1; break; }
                    s_S[j][j] = sqrt(s);
                } else {
                    s_S[j][j2] = s / s_S[j2][j2];
                }
            }
        }
        if (bad) a.status[r] = bad;
        for (int j = 0; j < q; j++)
            for (int j2 = 0; j2 < q; j2++) s_Ci[j][j2] = 0.0;
        if (!bad)
            for (int c = 0; c < q; c++) {          // column c of C^-1 by forward substitution
                for (int j = c; j < q; j++) {
                    double s = (j == c) ? 1.0 : 0.0;
                    for (int t = c; t < j; t++) s -= s_S[j][t] * s_Ci[t][c];
                    s_Ci[j][c] = s / s_S[j][j];
                }
            }
        s_q = bad ? 0 : q;
    }
    __syncthreads();
    if (s_q == 0) { if (tid == 0) a.knew[r] = 0; return; }
    // t[j][m] = sum_{n >= m} l[n][j] W[n][m]  (a thread per column m: coalesced over m for every n), then
    // W_new[j][m] = -sum_{j2 <= j} Ci[j][j2] t[j2][m]
    for (int m = tid; m < N; m += BT_THREADS) {
        double t[BT_MAXA];
#pragma unroll
        for (int j = 0; j < BT_MAXA; j++) t[j] = 0.0;
#pragma unroll 4
        for (int n = m; n < N; n++) {
            const double w = W[(int64_t)n * cap + m];
#pragma unroll
            for (int j = 0; j < BT_MAXA; j++)
                if (j < q) t[j] = fma(lv[j * cap + n], w, t[j]);
        }
#pragma unroll
        for (int j = 0; j < BT_MAXA; j++)
            if (j < q) {
                double s = 0.0;
#pragma unroll
                for (int j2 = 0; j2 < BT_MAXA; j2++)
                    if (j2 <= j) s = fma(s_Ci[j][j2], t[j2], s);
                W[(int64_t)(N + j) * cap + m] = -s;
            }
    }
    if (tid < q * q) {
        const int j = tid / q, j2 = tid % q;
        W[(int64_t)(N + j) * cap + N + j2] = (j2 <= j) ? s_Ci[j][j2] : 0.0;
    }
    // axis-factor table rows of the new points: e(u, U) = exp(-0.5 ((u - U) / l)^2) with the division first (:77-78)
    for (int e = tid; e < q * (a.nx + a.ny); e += BT_THREADS) {
        const int j = e / (a.nx + a.ny), i = e % (a.nx + a.ny);
        const bool isx = i < a.nx;
        const double u = isx ? a.ux[i] : a.uy[i - a.nx], U = s_xn[j][isx ? 0 : 1];
        const double dH = u / a.p.l_H - U / a.p.l_H;
        const int64_t row = (int64_t)r * cap + N + j;
        if (isx) a.TxH[row * a.nx + i] = exp(-0.5 * (dH * dH)); else a.TyH[row * a.ny + (i - a.nx)] = exp(-0.5 * (dH * dH));
        if (a.p.multi) {
            const double dL = u / a.p.l_L - U / a.p.l_L;
            if (isx) a.TxL[row * a.nx + i] = exp(-0.5 * (dL * dL)); else a.TyL[row * a.ny + (i - a.nx)] = exp(-0.5 * (dL * dL));
        }
    }
    __syncthreads();
    // z_new[j] = W_new[j][:] (y - m): a warp per new row
    for (int j = warp; j < q; j += nwarp) {
        const double* wr = W + (int64_t)(N + j) * cap;
        double s = 0.0;
        for (int m = lane; m < N + q; m += 32) s = fma(wr[m], yv[m] - (m < a.NL ? a.p.mean_L : a.p.mean_H), s);
        s = warp_sum_d(s);
        if (lane == 0) z[N + j] = s;
    }
    if (tid == 0) a.Ncur[r] = N + q;
}

// ---- step 2: the standing posterior takes the new rows ---------------------------------------------------------------------
__global__ void __launch_bounds__(BT_THREADS) batch_posterior_kernel(BatchArgs a) {
    const int r = blockIdx.y;
    const int q = a.knew[r];
    if (q == 0 || a.status[r] != 0) return;
    const int g = blockIdx.x * BT_THREADS + threadIdx.x;
    if (g >= a.G) return;
    const int Nn = a.Ncur[r], N0 = Nn - q, cap = a.cap;
    const int ix = g / a.ny, iy = g % a.ny;
    const double* W = a.W + (int64_t)r * cap * cap + (int64_t)N0 * cap;
    const double* TxH = a.TxH + (int64_t)r * cap * a.nx;
    const double* TyH = a.TyH + (int64_t)r * cap * a.ny;
    const double* TxL = a.TxL + (int64_t)r * cap * a.nx;
    const double* TyL = a.TyL + (int64_t)r * cap * a.ny;
    double v[BT_MAXA];
#pragma unroll
    for (int j = 0; j < BT_MAXA; j++) v[j] = 0.0;
    const double cLL = a.p.rho * a.p.s_L, cLH = a.p.rho2 * a.p.s_L;      // gaussian_process.py:426-429
    for (int n = 0; n < Nn; n++) {
        double psi = 0.0;
        if (n >= a.NL) psi = a.p.s_H * (__ldg(TxH + (int64_t)n * a.nx + ix) * __ldg(TyH + (int64_t)n * a.ny + iy));
        if (a.p.multi) psi = fma(n < a.NL ? cLL : cLH, __ldg(TxL + (int64_t)n * a.nx + ix) * __ldg(TyL + (int64_t)n * a.ny + iy), psi);
#pragma unroll
        for (int j = 0; j < BT_MAXA; j++)
            if (j < q) v[j] = fma(__ldg(W + (int64_t)j * cap + n), psi, v[j]);
    }
    const double* z = a.z + (int64_t)r * cap + N0;
    double dm = 0.0, dq = 0.0;
#pragma unroll
    for (int j = 0; j < BT_MAXA; j++)
        if (j < q) { dm = fma(v[j], z[j], dm); dq = fma(v[j], v[j], dq); }
    a.mu[(int64_t)r * a.G + g] += dm;
    a.var[(int64_t)r * a.G + g] -= dq;
}

// The same update with PT consecutive grid points (same ix) per thread: the q loads of W_new[:, n] and the two x-table entries of
// a training point serve PT points, the y-table entries come as 16-byte loads -- 14 loads for 44 FMAs (q = 8, PT = 4) instead of
// 12 for 11.  Per point the same operations in the same order as batch_posterior_kernel: bitwise the same results.
template <int QM, int PT>          // q <= QM appended rows, PT in {2, 4} points per thread (ny % PT == 0)
__global__ void __launch_bounds__(BT_THREADS) batch_posterior_tiled_kernel(BatchArgs a) {
    const int r = blockIdx.y;
    const int q = a.knew[r];
    if (q == 0 || a.status[r] != 0) return;
    const int g = (blockIdx.x * BT_THREADS + threadIdx.x) * PT;
    if (g >= a.G) return;
    const int Nn = a.Ncur[r], N0 = Nn - q, cap = a.cap;
    const int ix = g / a.ny, iy = g % a.ny;
    const double* W = a.W + (int64_t)r * cap * cap + (int64_t)N0 * cap;
    const double* TxH = a.TxH + (int64_t)r * cap * a.nx + ix;
    const double* TyH = a.TyH + (int64_t)r * cap * a.ny + iy;
    const double* TxL = a.TxL + (int64_t)r * cap * a.nx + ix;
    const double* TyL = a.TyL + (int64_t)r * cap * a.ny + iy;
    double v[QM][PT];
#pragma unroll
    for (int j = 0; j < QM; j++)
#pragma unroll
        for (int p = 0; p < PT; p++) v[j][p] = 0.0;
    const double cLL = a.p.rho * a.p.s_L, cLH = a.p.rho2 * a.p.s_L;      // gaussian_process.py:426-429
    for (int n = 0; n < Nn; n++) {
        double psi[PT];
#pragma unroll
        for (int p = 0; p < PT; p++) psi[p] = 0.0;
        if (n >= a.NL) {
            const double tx = __ldg(TxH + (int64_t)n * a.nx);
#pragma unroll
            for (int p = 0; p < PT; p += 2) {
                const double2 ty = __ldg(reinterpret_cast<const double2*>(TyH + (int64_t)n * a.ny + p));
                psi[p] = a.p.s_H * (tx * ty.x); psi[p + 1] = a.p.s_H * (tx * ty.y);
            }
        }
        if (a.p.multi) {
            const double c = n < a.NL ? cLL : cLH, tx = __ldg(TxL + (int64_t)n * a.nx);
#pragma unroll
            for (int p = 0; p < PT; p += 2) {
                const double2 ty = __ldg(reinterpret_cast<const double2*>(TyL + (int64_t)n * a.ny + p));
                psi[p] = fma(c, tx * ty.x, psi[p]); psi[p + 1] = fma(c, tx * ty.y, psi[p + 1]);
            }
        }
#pragma unroll
        for (int j = 0; j < QM; j++)
            if (j < q) {
                const double w = __ldg(W + (int64_t)j * cap + n);
#pragma unroll
                for (int p = 0; p < PT; p++) v[j][p] = fma(w, psi[p], v[j][p]);
            }
    }
    const double* z = a.z + (int64_t)r * cap + N0;
#pragma unroll
    for (int p = 0; p < PT; p++) {
        double dm = 0.0, dq = 0.0;
#pragma unroll
        for (int j = 0; j < QM; j++)
            if (j < q) { dm = fma(v[j][p], z[j], dm); dq = fma(v[j][p], v[j][p], dq); }
        a.mu[(int64_t)r * a.G + g + p] += dm;
        a.var[(int64_t)r * a.G + g + p] -= dq;
    }
}

// ---- step 3: coverage step, finishing, log, decision, move -------------------------------------------------------------------
constexpr int BC_SLOTS = 8;          // per cell: sum w, sum w x, sum w y, count, max var, arg-max index, loss sum, loss count
constexpr int BC_MAXV = 24;          // a cell of <= 16 seeds in a box has at most 4 + 15 vertices

__global__ void __launch_bounds__(BT_THREADS) batch_coverage_kernel(BatchArgs a, int it) {
    extern __shared__ __align__(16) unsigned short s_mask[];    // [2][G] membership bit masks (Lloyd partition, loss partition)
    __shared__ double s_seed[2][BT_MAXA][2];                    // [0]: Lloyd partition (centroids), [1]: loss partition (positions)
    __shared__ double s_poly[2][BT_MAXA][2 * BC_MAXV];
    __shared__ int s_cnt[2][BT_MAXA];
    __shared__ double s_area[2][BT_MAXA];
    __shared__ double s_buf[BT_THREADS / 32][256];
    __shared__ double s_tot[BT_MAXA][BC_SLOTS];
    __shared__ int s_ties, s_bad;
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = BT_THREADS / 32;
    const int A = a.A;
    if (a.status[r] != 0) return;
    double* pos = a.pos + (int64_t)r * A * 2;
    double* prev = a.prev + (int64_t)r * A * 2;
    double* cen = a.cen + (int64_t)r * A * 2;
    if (tid < 2 * A) { s_seed[0][tid >> 1][tid & 1] = cen[tid]; s_seed[1][tid >> 1][tid & 1] = pos[tid]; }
    if (tid == 0) { s_ties = 0; s_bad = 0; }
    __syncthreads();
    // bounded Voronoi cells of both partitions: the box inflated by eps/2 clipped by the bisectors (simulator.py:154-191)
    const double h = 0.5 * a.eps;
    for (int c = warp; c < 2 * A; c += nwarp) {
        const int part = c / A, i = c % A;
        bool overflow = false;
        int which = 0;
        const int n = vc_clip_cell(&s_seed[part][0][0], A, i, a.xmin - h, a.xmax + h, a.ymin - h, a.ymax + h, s_buf[warp], lane, overflow, which);
        const double* px = s_buf[warp] + 128 * which;
        const double* py = px + 64;
        if (n > BC_MAXV) overflow = true;
        for (int v = lane; v < n && v < BC_MAXV; v += 32) { s_poly[part][i][2 * v] = px[v]; s_poly[part][i][2 * v + 1] = py[v]; }
        if (lane == 0) {
            double s1 = 0.0, s2 = 0.0;
            for (int v = 0; v < n; v++) {
                const int u = (v + n - 1) % n;
                s1 += px[v] * py[u];
                s2 += py[v] * px[u];
            }
            s_cnt[part][i] = n < BC_MAXV ? n : BC_MAXV;
            s_area[part][i] = 0.5 * fabs(s1 - s2);
            if (overflow) s_bad = 1;
        }
        __syncwarp();
    }
    __syncthreads();
    const double* w = (a.algo == 0) ? a.f : a.mu + (int64_t)r * a.G;
    const double* var = a.var + (int64_t)r * a.G;
    const bool with_var = a.algo != 0;
    const TieRule tol{a.p.k0, a.amax_rel};
    // pass 1: membership masks of every grid point in both partitions
    int hard = 0;
    for (int g = tid; g < a.G; g += BT_THREADS) {
        const double x = a.xy[2 * g], y = a.xy[2 * g + 1];
#pragma unroll
        for (int part = 0; part < 2; part++) {
            double best = DBL_MAX, second = DBL_MAX;
            int bi = 0;
            for (int c = 0; c < A; c++) {
                const double dx = x - s_seed[part][c][0], dy = y - s_seed[part][c][1];
                const double d = fma(dx, dx, dy * dy);
                if (d < best) { second = best; best = d; bi = c; }
                else if (d < second) second = d;
            }
            const double gap = second - best;
            unsigned mask = 0;
            if (gap > a.tie_tol) {
                mask = 1u << bi;
            } else {                                   // within TIE_TOL of a bisector: the reference's crossings test decides
                if (!(gap > 0.1 * a.tie_tol)) hard++;
                for (int c = 0; c < A; c++)
                    if (crossings_inside(s_poly[part][c], s_cnt[part][c], x, y)) mask |= 1u << c;
            }
            s_mask[part * a.G + g] = (unsigned short)mask;
        }
    }
    if (hard) atomicAdd(&s_ties, hard);
    __syncthreads();
    // pass 2: a warp per cell; lanes stride over the grid in index order, one fixed butterfly per quantity
    for (int c = warp; c < A; c += nwarp) {
        double sw = 0.0, swx = 0.0, swy = 0.0, sl = 0.0;
        int cn = 0, ln = 0;
        ArgMax am{0.0, -1};
        const unsigned short bit = (unsigned short)(1u << c);
        const double sx = s_seed[1][c][0], sy = s_seed[1][c][1];
        for (int g = lane; g < a.G; g += 32) {
            if (s_mask[g] & bit) {
                const double x = a.xy[2 * g], y = a.xy[2 * g + 1], wv = w[g];
                sw += wv; swx += wv * x; swy += wv * y; cn++;
                if (with_var) am = argmax_combine(am, ArgMax{var[g], (long long)g}, tol);
            }
            if (s_mask[a.G + g] & bit) {
                const double dx = a.xy[2 * g] - sx, dy = a.xy[2 * g + 1] - sy;
                sl += __dmul_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), a.f[g]);     // simulator.py:215-216
                ln++;
            }
        }
        const double t0 = warp_sum_d(sw), t1 = warp_sum_d(swx), t2 = warp_sum_d(swy), t3 = warp_sum_d(sl);
        const int n0 = __reduce_add_sync(0xffffffffu, cn), n1 = __reduce_add_sync(0xffffffffu, ln);
        const ArgMax m = with_var ? argmax_warp(am, tol) : ArgMax{0.0, -1};
        if (lane == 0) {
            double* o = s_tot[c];
            o[0] = t0; o[1] = t1; o[2] = t2; o[3] = (double)n0; o[4] = m.v; o[5] = __longlong_as_double(m.i); o[6] = t3; o[7] = (double)n1;
            if (with_var && m.i < 0) s_bad = 2;          // np.amax([]) raises ValueError (simulator.py:312)
        }
    }
    __syncthreads();
    if (tid == 0) {
        a.ties[(int64_t)r * a.iterations + it] = s_ties;
        if (s_bad) a.status[r] = (s_bad == 2) ? -4 : -5;
        double loss = 0.0;                                 // simulator.py:215-219, cell order
        for (int c = 0; c < A; c++) loss += (s_tot[c][6] / s_tot[c][7]) * s_area[1][c];
        a.log_loss[(int64_t)r * a.iterations + it] = loss;
    }
    if (tid < A) {          // centroid (simulator.py:256-271), log row (:740-768), decision (:771-775 / :942-943), move (:777-782)
        const int i = tid;
        const double n = s_tot[i][3], ar = s_area[0][i];
        const double f_int = (s_tot[i][0] / n) * ar;
        double cx = ((s_tot[i][1] / n) * ar) / f_int, cy = ((s_tot[i][2] / n) * ar) / f_int;
        if (cx < a.xmin) cx = a.xmin;
        if (cx > a.xmax) cx = a.xmax;
        if (cy < a.ymin) cy = a.ymin;
        if (cy > a.ymax) cy = a.ymax;
        const double px = pos[2 * i], py = pos[2 * i + 1];
        const double ddx = px - prev[2 * i], ddy = py - prev[2 * i + 1];
        const double dist = sqrt(ddx * ddx + ddy * ddy);
        const long long gi = __double_as_longlong(s_tot[i][5]);
        const double mv = with_var ? s_tot[i][4] : 0.0;
        const double axm = (with_var && gi >= 0) ? a.xy[2 * gi] : 0.0, aym = (with_var && gi >= 0) ? a.xy[2 * gi + 1] : 0.0;
        double* row = a.log_agent + (((int64_t)r * a.iterations + it) * A + i) * BT_LOGC;
        row[0] = px; row[1] = py; row[2] = axm; row[3] = py; row[4] = mv; row[5] = with_var ? a.p.k0 : 0.0;
        row[6] = cx; row[7] = cy; row[8] = a.prob[r * A + i]; row[9] = (double)a.explore[r * A + i]; row[10] = dist;
        int ex = 0;
        double pr = 0.0;
        if (a.algo == 1) { ex = ((it / 5) % 2 == 0) ? 1 : 0; pr = (double)ex; }
        else if (a.algo == 2) {
            pr = sqrt(mv / (a.p.k0 * (double)A));              // todescato_prob, simulator.py:457-467
            ex = a.unif[((int64_t)r * a.iterations + it) * A + i] < pr ? 1 : 0;
        }
        a.prob[r * A + i] = pr;
        a.explore[r * A + i] = ex;
        prev[2 * i] = px; prev[2 * i + 1] = py;
        if (a.algo == 0 || !ex) { pos[2 * i] = cx; pos[2 * i + 1] = cy; a.pos_idx[r * A + i] = -1; }
        else { pos[2 * i] = axm; pos[2 * i + 1] = aym; a.pos_idx[r * A + i] = gi; }
        cen[2 * i] = cx; cen[2 * i + 1] = cy;
    }
}

}  // namespace mfgp

using namespace mfgp;

// Layout of the caller-provided state block (doubles unless noted), per run r and in this order:
//   Xt[cap*2] y[cap] W[cap*cap] z[cap] TxL[cap*nx] TyL[cap*ny] TxH[cap*nx] TyH[cap*ny] mu[G] var[G] pos[2A] prev[2A] cen[2A]
//   prob[A] | int64: pos_idx[A] | int32: explore[A] Ncur knew status noise_used nsamples ties[iterations]
// (the host module mfgp-coverage_b200/_batched.py owns the buffers; this file only receives pointers)
extern "C" {

/* struct mfgp_batch: include/mfgp_b200.h */

int mfgp_batch_step(const mfgp_batch* b, const mfgp_params* p_host, int64_t iteration, void* stream) {
    if (!b || !p_host || b->runs <= 0 || b->A <= 0 || b->A > BT_MAXA || b->G <= 0 || b->nx * b->ny != b->G || b->cap <= 0 ||
        iteration < 0 || iteration >= b->iterations)
        return MFGP_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BatchArgs a{};
    a.runs = (int)b->runs; a.G = (int)b->G; a.nx = (int)b->nx; a.ny = (int)b->ny; a.A = (int)b->A; a.NL = (int)b->NL;
    a.cap = (int)b->cap; a.algo = (int)b->algo; a.iterations = (int)b->iterations; a.max_samples = (int)b->max_samples;
    a.xmin = b->xmin; a.xmax = b->xmax; a.ymin = b->ymin; a.ymax = b->ymax; a.eps = b->eps; a.tie_tol = b->tie_tol;
    a.amax_rel = b->amax_rel; a.p = make_dev_params(*p_host);
    a.xy = b->xy; a.f = b->f; a.ux = b->ux; a.uy = b->uy; a.Xt = b->Xt; a.y = b->y; a.W = b->W; a.z = b->z;
    a.TxL = b->TxL; a.TyL = b->TyL; a.TxH = b->TxH; a.TyH = b->TyH; a.mu = b->mu; a.var = b->var; a.pos = b->pos; a.prev = b->prev;
    a.cen = b->cen; a.pos_idx = reinterpret_cast<long long*>(b->pos_idx); a.prob = b->prob; a.explore = b->explore;
    a.Ncur = b->Ncur; a.knew = b->knew; a.status = b->status; a.noise_used = b->noise_used; a.nsamples = b->nsamples;
    a.ties = b->ties; a.noise = b->noise; a.unif = b->unif; a.log_loss = b->log_loss; a.log_agent = b->log_agent;
    a.log_sample = b->log_sample;
    if (a.algo != 0) {
        const size_t smem = sizeof(double) * 2 * (size_t)a.A * (size_t)a.cap;
        if (smem > 200 * 1024) return MFGP_ERR_INVALID;
        MFGP_CUDA_CHECK(cudaFuncSetAttribute(batch_append_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        batch_append_kernel<<<a.runs, BT_THREADS, smem, st>>>(a, (int)iteration);
        MFGP_LAUNCH_CHECK();
        // (the tables' rows must be 16-byte aligned for the tiled kernels: ny even; their point tiles must not straddle a column)
        if (a.A <= 8 && a.ny % 4 == 0 && a.G % 4 == 0)
            batch_posterior_tiled_kernel<8, 4><<<dim3((unsigned)((a.G / 4 + BT_THREADS - 1) / BT_THREADS), (unsigned)a.runs), BT_THREADS, 0, st>>>(a);
        else if (a.ny % 2 == 0 && a.G % 2 == 0)
            batch_posterior_tiled_kernel<BT_MAXA, 2><<<dim3((unsigned)((a.G / 2 + BT_THREADS - 1) / BT_THREADS), (unsigned)a.runs), BT_THREADS, 0, st>>>(a);
        else
            batch_posterior_kernel<<<dim3((unsigned)((a.G + BT_THREADS - 1) / BT_THREADS), (unsigned)a.runs), BT_THREADS, 0, st>>>(a);
        MFGP_LAUNCH_CHECK();
    }
    const size_t msmem = sizeof(unsigned short) * 2 * (size_t)a.G;
    if (msmem > 160 * 1024) return MFGP_ERR_INVALID;
    MFGP_CUDA_CHECK(cudaFuncSetAttribute(batch_coverage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    batch_coverage_kernel<<<a.runs, BT_THREADS, msmem, st>>>(a, (int)iteration);
    MFGP_LAUNCH_CHECK();
    return MFGP_OK;
}

}  // extern "C"
