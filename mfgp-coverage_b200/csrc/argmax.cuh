// First-index argmax (np.argmax semantics: the LOWEST index attaining the maximum) in two deterministic stages.
#pragma once
#include <cfloat>

#include "common.cuh"

namespace mfgp {

constexpr int COV_THREADS = 256;

// Tie-aware (value, index) combine.  np.argmax returns the FIRST index of the maximum, and the reference's variances
// tie in two different ways (hard part 6; DESIGN.md "Arg-max ties"):
//  (a) symmetric priors: mirror-image grid points carry bit-identical variances in the reference, ~1e-16 apart in any
//      other operation order;
//  (b) far from all data var = k0 - q with q ~ ulp(k0): a staircase of 1-ulp plateaus, first index of the TOP plateau.
// Rule: two values are tied when they differ by at most rel * (k0 - smaller value), i.e. a tolerance relative to the
// variance REDUCTION q.  In (a) q is large and the tolerance dwarfs arithmetic noise; in (b) it is far below one ulp, so
// the comparison is exact and plateaus are never merged.  Ties go to the lower index; the larger value is carried.
// rel = 0 gives plain first-index argmax.
struct ArgMax {
    double v; long long i;
};
struct TieRule {
    double k0, rel;
};
__device__ __forceinline__ ArgMax argmax_combine(ArgMax a, ArgMax b, TieRule t) {
    if (a.i < 0) return b;
    if (b.i < 0) return a;
    const double lo = a.v < b.v ? a.v : b.v;
    const double tol = t.rel > 0.0 ? fmax(t.rel * (t.k0 - lo), 0.0) : 0.0;
    const double d = a.v - b.v;
    if (d > tol) return a;
    if (-d > tol) return b;
    ArgMax r;
    r.v = a.v > b.v ? a.v : b.v;
    r.i = a.i < b.i ? a.i : b.i;
    return r;
}
__device__ __forceinline__ ArgMax argmax_warp(ArgMax x, TieRule tol) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax y;
        y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
        y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
        x = argmax_combine(x, y, tol);
    }
    return x;
}

// ---- global first-index argmax -------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) argmax_partial_kernel(const double* __restrict__ v, int64_t G, int64_t base_index,
                                                                    TieRule tol, double* __restrict__ pv, long long* __restrict__ pi) {
    ArgMax best{0.0, -1};
    for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < G; g += (int64_t)gridDim.x * 256)
        best = argmax_combine(best, ArgMax{v[g], (long long)(base_index + g)}, tol);
    __shared__ double sv[8]; __shared__ long long si[8];
    best = argmax_warp(best, tol);
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best.v; si[threadIdx.x >> 5] = best.i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) best = argmax_combine(best, ArgMax{sv[w], si[w]}, tol);
        pv[blockIdx.x] = best.v; pi[blockIdx.x] = best.i;
    }
}

static __global__ void argmax_final_kernel(const double* __restrict__ pv, const long long* __restrict__ pi, int n, TieRule tol,
                                           double* __restrict__ out_val, int64_t* __restrict__ out_idx) {
    ArgMax best{0.0, -1};
    for (int i = threadIdx.x; i < n; i += 32) best = argmax_combine(best, ArgMax{pv[i], pi[i]}, tol);
    best = argmax_warp(best, tol);
    if (threadIdx.x == 0) { *out_val = best.v; *out_idx = best.i; }
}

inline int cov_blocks(int64_t G) {
    int64_t b = (G + COV_THREADS - 1) / COV_THREADS;
    const int64_t cap = 148 * 4;   // persistent-style: a few CTAs per SM, grid-stride over the points
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace mfgp
