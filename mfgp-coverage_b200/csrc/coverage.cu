// Coverage step on the device: bounded-Voronoi membership of every grid point + per-cell reductions, one pass.
// Replaces, in the reference's simulator.py: in_polygon :105-124, compute_loss :194-228, compute_centroids :231-283,
// compute_max_var :286-323, compute_sample_clusters :377-412 (the polygons themselves come from the caller).
//
// Membership rule.  The reference tests every point against every Qhull cell polygon with matplotlib's crossings
// test.  Away from cell borders that is the nearest-seed rule, so the fast path is an fp64 argmin over the seeds with
// the runner-up tracked; only when the two smallest squared distances are within `tie_tol` (the point is within
// ~tie_tol/(2*seed distance) of a bisector) is the reference's test evaluated literally -- same fp64 subtract /
// multiply / compare sequence, no FMA contraction -- against all cell polygons, and the point then lands in 0, 1 or
// several cells exactly as it does in the reference.
//
// Reductions are deterministic: a fixed butterfly inside the warp per distinct cell, per-warp shared-memory slots,
// per-block partials in global memory, and a second kernel that adds the block partials in block order.
#include <cfloat>

#include "common.cuh"
#include "argmax.cuh"

namespace mfgp {

constexpr int COV_WARPS = COV_THREADS / 32;
constexpr int COV_MAX_CELLS = 256;
constexpr int C_SLOTS = 6;   // sum w, sum w x, sum w y, count, max var, argmax index (int64 bits)
constexpr int P_SLOTS = 2;   // sum d^2 f, count

struct CovPartition {
    const double* seeds; int A;
    const double* poly_xy; const int32_t* poly_off;
};

struct CovArgs {
    const double* xy; const double* w; const double* var; const double* f;
    int64_t G; int64_t base_index;
    CovPartition C, P;
    double tie_tol;
    TieRule amax_tol;     // tie rule of the per-cell arg-max (see argmax_combine)
    uint64_t* member_c;
    double* partials;     // [nblocks][C.A*C_SLOTS + P.A*P_SLOTS]
};

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

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// membership words of one point in one partition
template <int WORDS>
__device__ __forceinline__ void point_membership(const CovPartition& part, const double* __restrict__ s_seeds,
                                                 const double* __restrict__ s_poly, const int* __restrict__ s_off,
                                                 double x, double y, double tie_tol, uint64_t (&m)[WORDS]) {
#pragma unroll
    for (int k = 0; k < WORDS; k++) m[k] = 0;
    double best = DBL_MAX, second = DBL_MAX;
    int bi = -1;
    for (int c = 0; c < part.A; c++) {
        const double dx = x - s_seeds[2 * c], dy = y - s_seeds[2 * c + 1];
        const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (d < best) { second = best; best = d; bi = c; }
        else if (d < second) second = d;
    }
    if (second - best > tie_tol) {
#pragma unroll
        for (int k = 0; k < WORDS; k++)
            if ((bi >> 6) == k) m[k] = 1ull << (bi & 63);
    } else {
        for (int c = 0; c < part.A; c++) {
            const int o = s_off[c], n = s_off[c + 1] - o;
            if (crossings_inside(s_poly + 2 * o, n, x, y)) {
#pragma unroll
                for (int k = 0; k < WORDS; k++)
                    if ((c >> 6) == k) m[k] |= 1ull << (c & 63);
            }
        }
    }
}

template <int WORDS>
__global__ void __launch_bounds__(COV_THREADS) cov_assign_reduce_kernel(CovArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int Ac = a.C.A, Ap = a.P.A;
    const int nvc = Ac ? a.C.poly_off[Ac] : 0, nvp = Ap ? a.P.poly_off[Ap] : 0;
    double* s_seed_c = sm;                       // [Ac*2]
    double* s_seed_p = s_seed_c + 2 * Ac;        // [Ap*2]
    double* s_poly_c = s_seed_p + 2 * Ap;        // [nvc*2]
    double* s_poly_p = s_poly_c + 2 * nvc;       // [nvp*2]
    double* s_acc = s_poly_p + 2 * nvp;          // [COV_WARPS][Ac*C_SLOTS + Ap*P_SLOTS]
    const int stride = Ac * C_SLOTS + Ap * P_SLOTS;
    int* s_off_c = reinterpret_cast<int*>(s_acc + COV_WARPS * stride);   // [Ac+1]
    int* s_off_p = s_off_c + (Ac + 1);                                    // [Ap+1]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 2 * Ac; i += COV_THREADS) s_seed_c[i] = a.C.seeds[i];
    for (int i = tid; i < 2 * Ap; i += COV_THREADS) s_seed_p[i] = a.P.seeds[i];
    for (int i = tid; i < 2 * nvc; i += COV_THREADS) s_poly_c[i] = a.C.poly_xy[i];
    for (int i = tid; i < 2 * nvp; i += COV_THREADS) s_poly_p[i] = a.P.poly_xy[i];
    for (int i = tid; i <= Ac && Ac; i += COV_THREADS) s_off_c[i] = a.C.poly_off[i];
    for (int i = tid; i <= Ap && Ap; i += COV_THREADS) s_off_p[i] = a.P.poly_off[i];
    for (int i = tid; i < COV_WARPS * stride; i += COV_THREADS) s_acc[i] = 0.0;
    __syncthreads();
    double* wacc = s_acc + warp * stride;
    if (lane == 0)
        for (int c = 0; c < Ac; c++) {
            wacc[c * C_SLOTS + 4] = -DBL_MAX;
            reinterpret_cast<long long*>(wacc)[c * C_SLOTS + 5] = -1;
        }
    __syncwarp();

    for (int64_t base = (int64_t)blockIdx.x * COV_THREADS; base < a.G; base += (int64_t)gridDim.x * COV_THREADS) {
        const int64_t g = base + tid;
        const bool valid = g < a.G;
        double x = 0, y = 0, wv = 0, vv = 0, fv = 0;
        if (valid) {
            const double2 p = reinterpret_cast<const double2*>(a.xy)[g];
            x = p.x; y = p.y;
            if (a.w) wv = a.w[g];
            if (a.var) vv = a.var[g];
            if (a.f) fv = a.f[g];
        }
        if (Ac) {
            uint64_t m[WORDS];
            point_membership<WORDS>(a.C, s_seed_c, s_poly_c, s_off_c, x, y, a.tie_tol, m);
            if (!valid) {
#pragma unroll
                for (int k = 0; k < WORDS; k++) m[k] = 0;
            }
            if (a.member_c && valid) {
#pragma unroll
                for (int k = 0; k < WORDS; k++) a.member_c[g * WORDS + k] = m[k];
            }
            const double wx = wv * x, wy = wv * y;
#pragma unroll
            for (int k = 0; k < WORDS; k++) {
                unsigned pending;
                while ((pending = __ballot_sync(0xffffffffu, m[k] != 0)) != 0) {
                    const int leader = __ffs(pending) - 1;
                    const uint64_t lm = __shfl_sync(0xffffffffu, m[k], leader);
                    const int bit = __ffsll((long long)lm) - 1;
                    const bool mine = (m[k] >> bit) & 1ull;
                    const int c = k * 64 + bit;
                    const double s0 = warp_sum(mine ? wv : 0.0);
                    const double s1 = warp_sum(mine ? wx : 0.0);
                    const double s2 = warp_sum(mine ? wy : 0.0);
                    const int cnt = __popc(__ballot_sync(0xffffffffu, mine));
                    const ArgMax am = argmax_warp(ArgMax{vv, mine ? (long long)(a.base_index + g) : -1LL}, a.amax_tol);
                    if (lane == 0) {
                        double* slot = wacc + c * C_SLOTS;
                        slot[0] += s0; slot[1] += s1; slot[2] += s2; slot[3] += (double)cnt;
                        long long* islot = reinterpret_cast<long long*>(slot);
                        if (a.var) {
                            const ArgMax r = argmax_combine(ArgMax{slot[4], islot[5]}, am, a.amax_tol);
                            slot[4] = r.v; islot[5] = r.i;
                        }
                    }
                    m[k] &= ~(1ull << bit);
                }
            }
        }
        if (Ap) {
            uint64_t m[WORDS];
            point_membership<WORDS>(a.P, s_seed_p, s_poly_p, s_off_p, x, y, a.tie_tol, m);
            if (!valid) {
#pragma unroll
                for (int k = 0; k < WORDS; k++) m[k] = 0;
            }
#pragma unroll
            for (int k = 0; k < WORDS; k++) {
                unsigned pending;
                while ((pending = __ballot_sync(0xffffffffu, m[k] != 0)) != 0) {
                    const int leader = __ffs(pending) - 1;
                    const uint64_t lm = __shfl_sync(0xffffffffu, m[k], leader);
                    const int bit = __ffsll((long long)lm) - 1;
                    const bool mine = (m[k] >> bit) & 1ull;
                    const int c = k * 64 + bit;
                    // simulator.py:215-216: (dx^2 + dy^2) * f
                    const double dx = x - s_seed_p[2 * c], dy = y - s_seed_p[2 * c + 1];
                    const double pl = __dmul_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), fv);
                    const double s0 = warp_sum(mine ? pl : 0.0);
                    const int cnt = __popc(__ballot_sync(0xffffffffu, mine));
                    if (lane == 0) {
                        double* slot = wacc + Ac * C_SLOTS + c * P_SLOTS;
                        slot[0] += s0; slot[1] += (double)cnt;
                    }
                    m[k] &= ~(1ull << bit);
                }
            }
        }
    }
    __syncthreads();
    // block partial = warp slots combined in warp order
    double* out = a.partials + (int64_t)blockIdx.x * stride;
    for (int i = tid; i < Ac * C_SLOTS; i += COV_THREADS) {
        const int slot = i % C_SLOTS;
        if (slot < 4) {
            double s = 0.0;
            for (int w = 0; w < COV_WARPS; w++) s += s_acc[w * stride + i];
            out[i] = s;
        } else if (slot == 4) {
            ArgMax best{0.0, -1};
            for (int w = 0; w < COV_WARPS; w++)
                best = argmax_combine(best, ArgMax{s_acc[w * stride + i], reinterpret_cast<const long long*>(s_acc)[w * stride + i + 1]},
                                      a.amax_tol);
            out[i] = best.v;
            reinterpret_cast<long long*>(out)[i + 1] = best.i;
        }
    }
    for (int i = tid; i < Ap * P_SLOTS; i += COV_THREADS) {
        double s = 0.0;
        for (int w = 0; w < COV_WARPS; w++) s += s_acc[w * stride + Ac * C_SLOTS + i];
        out[Ac * C_SLOTS + i] = s;
    }
}

__global__ void cov_finalize_kernel(const double* __restrict__ partials, int nblocks, int Ac, int Ap, TieRule amax_tol, double* __restrict__ cent,
                                    double* __restrict__ amax_val, int64_t* __restrict__ amax_idx, double* __restrict__ lossp) {
    const int stride = Ac * C_SLOTS + Ap * P_SLOTS;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Ac * 4) {
        const int c = i / 4, s = i % 4;
        double acc = 0.0;
        for (int b = 0; b < nblocks; b++) acc += partials[(int64_t)b * stride + c * C_SLOTS + s];
        if (cent) cent[i] = acc;
    } else if (i < Ac * 5) {
        const int c = i - Ac * 4;
        ArgMax best{0.0, -1};
        for (int b = 0; b < nblocks; b++)
            best = argmax_combine(best, ArgMax{partials[(int64_t)b * stride + c * C_SLOTS + 4],
                                               reinterpret_cast<const long long*>(partials)[(int64_t)b * stride + c * C_SLOTS + 5]},
                                  amax_tol);
        if (amax_val) amax_val[c] = best.v;
        if (amax_idx) amax_idx[c] = best.i;
    } else if (i < Ac * 5 + Ap * 2) {
        const int j = i - Ac * 5;
        double acc = 0.0;
        for (int b = 0; b < nblocks; b++) acc += partials[(int64_t)b * stride + Ac * C_SLOTS + j];
        if (lossp) lossp[j] = acc;
    }
}

}  // namespace mfgp

using namespace mfgp;

extern "C" int64_t cov_workspace_bytes(int64_t G, int64_t Ac, int64_t Ap) {
    const int64_t stride = Ac * C_SLOTS + Ap * P_SLOTS;
    int64_t a = (int64_t)cov_blocks(G) * stride * 8;
    int64_t b = (int64_t)cov_blocks(G) * 16;
    return (a > b ? a : b) + 256;
}

extern "C" int cov_assign_reduce(const double* xy, const double* w, const double* var, const double* f, int64_t G,
                                 int64_t base_index, const double* seeds_c, int64_t Ac, const double* poly_xy_c,
                                 const int32_t* poly_off_c, int64_t nvert_c, const double* seeds_p, int64_t Ap,
                                 const double* poly_xy_p, const int32_t* poly_off_p, int64_t nvert_p, double tie_tol, double amax_k0, double amax_rel, double* cent, double* amax_val,
                                 int64_t* amax_idx, double* lossp, uint64_t* member_c, void* work, int64_t work_bytes,
                                 void* stream) {
    if (!xy || G <= 0 || Ac < 0 || Ap < 0 || Ac + Ap == 0 || Ac > COV_MAX_CELLS || Ap > COV_MAX_CELLS) return MFGP_ERR_INVALID;
    if (Ac && (!seeds_c || !poly_xy_c || !poly_off_c)) return MFGP_ERR_INVALID;
    if (Ap && (!seeds_p || !poly_xy_p || !poly_off_p || !f)) return MFGP_ERR_INVALID;
    if (!work || work_bytes < cov_workspace_bytes(G, Ac, Ap)) return MFGP_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (nvert_c < 0 || nvert_p < 0) return MFGP_ERR_INVALID;
    const int64_t nvc = Ac ? nvert_c : 0, nvp = Ap ? nvert_p : 0;   // == poly_off[A], passed by the host to size shared memory
    CovArgs a;
    a.xy = xy; a.w = w; a.var = var; a.f = f; a.G = G; a.base_index = base_index;
    a.C = {seeds_c, (int)Ac, poly_xy_c, poly_off_c};
    a.P = {seeds_p, (int)Ap, poly_xy_p, poly_off_p};
    a.tie_tol = tie_tol; a.amax_tol = TieRule{amax_k0, amax_rel > 0.0 ? amax_rel : 0.0}; a.member_c = member_c; a.partials = static_cast<double*>(work);
    const int stride = (int)(Ac * C_SLOTS + Ap * P_SLOTS);
    const size_t smem = sizeof(double) * (2 * Ac + 2 * Ap + 2 * (size_t)nvc + 2 * (size_t)nvp + (size_t)COV_WARPS * stride) +
                        sizeof(int) * (Ac + Ap + 2);
    if (smem > 200 * 1024) return MFGP_ERR_INVALID;
    const int nblocks = cov_blocks(G);
    const int words = (int)((Ac > Ap ? Ac : Ap) + 63) / 64;
    auto launch = [&](auto kern) -> int {
        MFGP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<nblocks, COV_THREADS, smem, st>>>(a);
        MFGP_LAUNCH_CHECK();
        return MFGP_OK;
    };
    int rc;
    if (words <= 1) rc = launch(cov_assign_reduce_kernel<1>);
    else if (words == 2) rc = launch(cov_assign_reduce_kernel<2>);
    else rc = launch(cov_assign_reduce_kernel<4>);
    if (rc) return rc;
    const int nfin = (int)(Ac * 5 + Ap * 2);
    cov_finalize_kernel<<<(nfin + 127) / 128, 128, 0, st>>>(a.partials, nblocks, (int)Ac, (int)Ap, a.amax_tol, cent, amax_val, amax_idx, lossp);
    MFGP_LAUNCH_CHECK();
    return MFGP_OK;
}

extern "C" int cov_argmax(const double* v, int64_t G, int64_t base_index, double k0, double rel, double* out_val,
                          int64_t* out_idx, void* work, int64_t work_bytes, void* stream) {
    const TieRule tol{k0, rel > 0.0 ? rel : 0.0};
    if (!v || G <= 0 || !out_val || !out_idx || !work) return MFGP_ERR_INVALID;
    const int nblocks = cov_blocks(G);
    if (work_bytes < (int64_t)nblocks * 16) return MFGP_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* pv = static_cast<double*>(work);
    long long* pi = reinterpret_cast<long long*>(pv + nblocks);
    argmax_partial_kernel<<<nblocks, 256, 0, st>>>(v, G, base_index, tol, pv, pi);
    MFGP_LAUNCH_CHECK();
    argmax_final_kernel<<<1, 32, 0, st>>>(pv, pi, nblocks, tol, out_val, out_idx);
    MFGP_LAUNCH_CHECK();
    return MFGP_OK;
}
