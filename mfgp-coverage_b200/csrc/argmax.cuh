// First-index argmax (np.argmax semantics: the LOWEST index attaining the maximum) in two deterministic stages.
#pragma once
#include <cfloat>

#include "common.cuh"

namespace mfgp {

constexpr int COV_THREADS = 256;

// ---- global first-index argmax -------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) argmax_partial_kernel(const double* __restrict__ v, int64_t G, int64_t base_index,
                                                             double* __restrict__ pv, long long* __restrict__ pi) {
    double bv = -DBL_MAX; long long bi = 0x7fffffffffffffffLL;
    for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < G; g += (int64_t)gridDim.x * 256) {
        const double x = v[g];
        if (x > bv) { bv = x; bi = base_index + g; }   // ascending g: strict '>' keeps the first index
    }
    __shared__ double sv[8]; __shared__ long long si[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++)
            if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
        pv[blockIdx.x] = bv; pi[blockIdx.x] = bi;
    }
}

static __global__ void argmax_final_kernel(const double* __restrict__ pv, const long long* __restrict__ pi, int n,
                                    double* __restrict__ out_val, int64_t* __restrict__ out_idx) {
    double bv = -DBL_MAX; long long bi = 0x7fffffffffffffffLL;
    for (int i = threadIdx.x; i < n; i += 32)
        if (pv[i] > bv || (pv[i] == bv && pi[i] < bi)) { bv = pv[i]; bi = pi[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (threadIdx.x == 0) { *out_val = bv; *out_idx = bi; }
}

inline int cov_blocks(int64_t G) {
    int64_t b = (G + COV_THREADS - 1) / COV_THREADS;
    const int64_t cap = 148 * 4;   // persistent-style: a few CTAs per SM, grid-stride over the points
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace mfgp
