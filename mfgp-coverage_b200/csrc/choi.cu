// Choi greedy sample planner on the cached V = L^-1 Psi^T (replaces compute_sample_points, reference
// simulator.py:326-374, which refits and re-predicts the whole grid for every pick).
//
// Every candidate is a grid point, so appending pick j to the training set is a bordered Cholesky step whose new
// factor row is the cached column l = V[:, j]:  d = sqrt(k(0) + noise_H + jitter - l.l),
// v[g] = (k_HH(x_g, x_j) - sum_n l[n] V[n, g]) / d,  var[g] -= v[g]^2.  The pseudo-observation equals the current
// mean, so the posterior mean does not change.  One pick = one pass over V: 8*n*G bytes,
// HBM-bound.  The same kernel produces the block-level candidates of the next first-index argmax.
#include <cfloat>

#include "common.cuh"
#include "argmax.cuh"

namespace mfgp {

constexpr int CH_THREADS = 128;
constexpr int CH_COLS = CH_THREADS * 2;

struct ChoiArgs {
    const double* Xs; int64_t G;
    double* Vc; int64_t ldv; int n;          // rows [0,n) valid; row n is written
    double* var;
    const long long* pick;                   // device: grid index of the point being appended
    DevParams p;
    TieRule tol;                             // arg-max tie rule
    double* q;                               // running sum of v^2 per grid point (variance reduction)
    double* pv; long long* pi;               // per-block argmax candidates of the updated variance
};

__global__ void __launch_bounds__(CH_THREADS) choi_append_kernel(ChoiArgs a) {
    extern __shared__ __align__(16) double l[];   // [n]
    __shared__ double red[CH_THREADS / 32];
    __shared__ double sv[CH_THREADS / 32];
    __shared__ long long si[CH_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t j = *a.pick;
    double part = 0.0;
    for (int n = tid; n < a.n; n += CH_THREADS) {
        const double v = a.Vc[(int64_t)n * a.ldv + j];
        l[n] = v;
        part += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    double ll = 0.0;
#pragma unroll
    for (int w = 0; w < CH_THREADS / 32; w++) ll += red[w];
    const DevParams& p = a.p;
    const double d = sqrt(p.k0 + p.noise_H + p.jitter - ll);
    const double xj = a.Xs[2 * j], yj = a.Xs[2 * j + 1];

    const int64_t g = (int64_t)blockIdx.x * CH_COLS + tid * 2;
    ArgMax best{0.0, -1};
    if (g < a.G) {
        const bool pair = (g + 1 < a.G) && ((a.ldv & 1) == 0);
        double s0 = 0.0, s1 = 0.0;
        if (pair) {
            const double* col = a.Vc + g;
            int n = 0;
            for (; n + 8 <= a.n; n += 8) {
                double2 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = *reinterpret_cast<const double2*>(col + (int64_t)(n + u) * a.ldv);
#pragma unroll
                for (int u = 0; u < 8; u++) { s0 += l[n + u] * v[u].x; s1 += l[n + u] * v[u].y; }
            }
            for (; n < a.n; n++) {
                const double2 v = *reinterpret_cast<const double2*>(col + (int64_t)n * a.ldv);
                s0 += l[n] * v.x; s1 += l[n] * v.y;
            }
        } else {
            for (int n = 0; n < a.n; n++) {
                s0 += l[n] * a.Vc[(int64_t)n * a.ldv + g];
                if (g + 1 < a.G) s1 += l[n] * a.Vc[(int64_t)n * a.ldv + g + 1];
            }
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int64_t gg = g + c;
            if (gg < a.G) {
                const double x = a.Xs[2 * gg], y = a.Xs[2 * gg + 1];
                double k;
                const double kH = rbf_scaled(xj / p.l_H, yj / p.l_H, x / p.l_H, y / p.l_H, p.s_H);
                if (p.multi) {
                    const double kL = rbf_scaled(xj / p.l_L, yj / p.l_L, x / p.l_L, y / p.l_L, p.s_L);
                    k = __dadd_rn(__dmul_rn(p.rho2, kL), kH);
                } else {
                    k = kH;
                }
                const double v = (k - (c ? s1 : s0)) / d;
                a.Vc[(int64_t)a.n * a.ldv + gg] = v;
                const double nq = a.q[gg] + v * v;         // var is recomputed from the accumulated reduction, as the
                a.q[gg] = nq;                              // reference's full predict does (k** - psi beta), never decremented
                const double nv = p.k0 - nq;
                a.var[gg] = nv;
                best = argmax_combine(best, ArgMax{nv, (long long)gg}, a.tol);
            }
        }
    }
    best = argmax_warp(best, a.tol);
    if (lane == 0) { sv[warp] = best.v; si[warp] = best.i; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < CH_THREADS / 32; w++) best = argmax_combine(best, ArgMax{sv[w], si[w]}, a.tol);
        a.pv[blockIdx.x] = best.v; a.pi[blockIdx.x] = best.i;
    }
}

}  // namespace mfgp

using namespace mfgp;

extern "C" int64_t choi_greedy(const double* Xs, int64_t G, double* Vc, int64_t ldv, int64_t n0, int64_t cap, double* var,
                               double* q, const mfgp_params* p_host, double threshold, double tie_rel, int64_t max_picks,
                               int64_t* picks_host,
                               void* work, int64_t work_bytes, void* stream) {
    if (!Xs || !Vc || !var || !q || !p_host || !picks_host || !work || G <= 0 || ldv < G || n0 < 0 || cap < n0) return MFGP_ERR_INVALID;
    const int nblocks = (int)((G + CH_COLS - 1) / CH_COLS);
    const int64_t need = (int64_t)nblocks * 16 + 64 + cov_workspace_bytes(G, 1, 0);
    if (work_bytes < need) return MFGP_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* pv = static_cast<double*>(work);
    long long* pi = reinterpret_cast<long long*>(pv + nblocks);
    double* d_val = reinterpret_cast<double*>(pi + nblocks);
    int64_t* d_idx = reinterpret_cast<int64_t*>(d_val + 1);
    void* awork = d_idx + 7;
    const int64_t awork_bytes = work_bytes - ((char*)awork - (char*)work);
    const DevParams dp0 = make_dev_params(*p_host);
    const TieRule tol{dp0.k0, tie_rel > 0.0 ? tie_rel : 0.0};
    int rc = cov_argmax(var, G, 0, tol.k0, tol.rel, d_val, d_idx, awork, awork_bytes, st);
    if (rc) return rc;
    struct { double val; int64_t idx; } h;
    const DevParams dp = make_dev_params(*p_host);
    int64_t n = n0, picks = 0;
    static_assert(sizeof(h) == 16, "layout");
    while (true) {
        MFGP_CUDA_CHECK(cudaMemcpyAsync(&h, d_val, 16, cudaMemcpyDeviceToHost, st));
        MFGP_CUDA_CHECK(cudaStreamSynchronize(st));
        if (!(h.val > threshold) || picks >= max_picks) break;    // simulator.py:345 `while max_var > threshold`
        if (n >= cap) return MFGP_ERR_INVALID;                    // V cache exhausted
        picks_host[picks++] = h.idx;
        ChoiArgs a;
        a.Xs = Xs; a.G = G; a.Vc = Vc; a.ldv = ldv; a.n = (int)n; a.var = var;
        a.pick = reinterpret_cast<const long long*>(d_idx); a.p = dp; a.tol = tol; a.q = q; a.pv = pv; a.pi = pi;
        const size_t smem = sizeof(double) * (size_t)(n > 0 ? n : 1);
        if (smem > 48 * 1024)
            MFGP_CUDA_CHECK(cudaFuncSetAttribute(choi_append_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        choi_append_kernel<<<nblocks, CH_THREADS, smem, st>>>(a);
        MFGP_LAUNCH_CHECK();
        argmax_final_kernel<<<1, 32, 0, st>>>(pv, pi, nblocks, tol, d_val, d_idx);
        MFGP_LAUNCH_CHECK();
        n++;
    }
    return picks;
}
