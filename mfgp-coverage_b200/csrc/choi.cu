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

// Device-resident loop state: the host launches a BATCH of (append, arg-max) pairs without synchronising; each pair
// checks `done` / the threshold on the device, so the greedy loop costs one host round trip per batch, not per pick.
struct ChoiState {
    long long n;          // valid rows of the V cache
    long long picks;      // picks recorded so far (all batches)
    long long done;       // 1: max var <= threshold or a limit was reached -> remaining launches of the batch are no-ops
    long long pick;       // grid index to append next (the current first-index arg-max)
    double val;           // its variance
};

struct ChoiArgs {
    const double* Xs; int64_t G;
    double* Vc; int64_t ldv;                 // rows [0, state->n) valid; row state->n is written
    double* var;
    ChoiState* state;
    DevParams p;
    TieRule tol;                             // arg-max tie rule
    double* q;                               // running sum of v^2 per grid point (variance reduction)
    double* pv; long long* pi;               // per-block argmax candidates of the updated variance
    double threshold; long long cap; long long max_picks;
    long long* picks_dev;                    // [max_picks] grid indices in selection order
};

__global__ void __launch_bounds__(CH_THREADS) choi_append_kernel(ChoiArgs a) {
    extern __shared__ __align__(16) double l[];   // [n]
    __shared__ double red[CH_THREADS / 32];
    __shared__ double sv[CH_THREADS / 32];
    __shared__ long long si[CH_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (a.state->done) return;                    // uniform for the whole grid
    const int an = (int)a.state->n;
    const int64_t j = a.state->pick;
    double part = 0.0;
    for (int n = tid; n < an; n += CH_THREADS) {
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
            for (; n + 16 <= an; n += 16) {       // 16 x 16 B in flight per thread: the pass over V is HBM-bound
                double2 v[16];
#pragma unroll
                for (int u = 0; u < 16; u++) v[u] = __ldcs(reinterpret_cast<const double2*>(col + (int64_t)(n + u) * a.ldv));
#pragma unroll
                for (int u = 0; u < 16; u++) { s0 += l[n + u] * v[u].x; s1 += l[n + u] * v[u].y; }
            }
            for (; n < an; n++) {
                const double2 v = *reinterpret_cast<const double2*>(col + (int64_t)n * a.ldv);
                s0 += l[n] * v.x; s1 += l[n] * v.y;
            }
        } else {
            for (int n = 0; n < an; n++) {
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
                a.Vc[(int64_t)an * a.ldv + gg] = v;
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

// One warp: the row just written becomes valid, the block candidates give the next arg-max, and the loop condition of
// simulator.py:345 (`while max_var > threshold`) is evaluated on the device.
__global__ void choi_advance_kernel(ChoiArgs a, int nblocks, int first) {
    ChoiState* st = a.state;
    if (st->done) return;
    ArgMax best{0.0, -1};
    if (first) {            // the arg-max of the incoming variance was computed by cov_argmax into (val, pick)
        best = ArgMax{st->val, st->pick};
    } else {
        for (int i = threadIdx.x; i < nblocks; i += 32) best = argmax_combine(best, ArgMax{a.pv[i], a.pi[i]}, a.tol);
        best = argmax_warp(best, a.tol);
    }
    if (threadIdx.x == 0) {
        if (!first) st->n += 1;
        st->val = best.v;
        st->pick = best.i;
        if (!(best.v > a.threshold) || st->picks >= a.max_picks || st->n >= a.cap) st->done = 1;
        else a.picks_dev[st->picks++] = best.i;
    }
}

}  // namespace mfgp

using namespace mfgp;

extern "C" int64_t choi_greedy(const double* Xs, int64_t G, double* Vc, int64_t ldv, int64_t n0, int64_t cap, double* var,
                               double* q, const mfgp_params* p_host, double threshold, double tie_rel, int64_t max_picks,
                               int64_t* picks_host,
                               void* work, int64_t work_bytes, void* stream) {
    if (!Xs || !Vc || !var || !q || !p_host || !picks_host || !work || G <= 0 || ldv < G || n0 < 0 || cap < n0) return MFGP_ERR_INVALID;
    if (max_picks > cap - n0) max_picks = cap - n0;           // the V cache holds cap rows: the caller grows it and resumes
    const int nblocks = (int)((G + CH_COLS - 1) / CH_COLS);
    const int64_t need = (int64_t)nblocks * 16 + 128 + (max_picks + 1) * 8 + cov_workspace_bytes(G, 1, 0);
    if (work_bytes < need) return MFGP_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* pv = static_cast<double*>(work);
    long long* pi = reinterpret_cast<long long*>(pv + nblocks);
    ChoiState* state = reinterpret_cast<ChoiState*>(pi + nblocks);
    long long* picks_dev = reinterpret_cast<long long*>(state + 2);
    void* awork = picks_dev + max_picks + 1;
    const int64_t awork_bytes = work_bytes - ((char*)awork - (char*)work);
    const DevParams dp = make_dev_params(*p_host);
    const TieRule tol{dp.k0, tie_rel > 0.0 ? tie_rel : 0.0};
    ChoiState h{n0, 0, 0, -1, 0.0};
    MFGP_CUDA_CHECK(cudaMemcpyAsync(state, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    int rc = cov_argmax(var, G, 0, tol.k0, tol.rel, &state->val, reinterpret_cast<int64_t*>(&state->pick), awork, awork_bytes, st);
    if (rc) return rc;
    ChoiArgs a;
    a.Xs = Xs; a.G = G; a.Vc = Vc; a.ldv = ldv; a.var = var; a.state = state; a.p = dp; a.tol = tol; a.q = q; a.pv = pv; a.pi = pi;
    a.threshold = threshold; a.cap = cap; a.max_picks = max_picks; a.picks_dev = picks_dev;
    choi_advance_kernel<<<1, 32, 0, st>>>(a, nblocks, 1);
    MFGP_LAUNCH_CHECK();
    constexpr int BATCH = 16;
    int64_t n_hi = n0;                // upper bound of state->n known to the host (sizes the shared-memory column)
    while (true) {
        for (int b = 0; b < BATCH; b++) {
            n_hi++;
            const size_t smem = sizeof(double) * (size_t)n_hi;
            if (smem > 48 * 1024)
                MFGP_CUDA_CHECK(cudaFuncSetAttribute(choi_append_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            choi_append_kernel<<<nblocks, CH_THREADS, smem, st>>>(a);
            MFGP_LAUNCH_CHECK();
            choi_advance_kernel<<<1, 32, 0, st>>>(a, nblocks, 0);
            MFGP_LAUNCH_CHECK();
        }
        MFGP_CUDA_CHECK(cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, st));
        MFGP_CUDA_CHECK(cudaStreamSynchronize(st));
        if (h.done) break;
        n_hi = h.n;
    }
    if (h.picks > 0) {
        MFGP_CUDA_CHECK(cudaMemcpyAsync(picks_host, picks_dev, sizeof(long long) * h.picks, cudaMemcpyDeviceToHost, st));
        MFGP_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    return h.picks;
}
