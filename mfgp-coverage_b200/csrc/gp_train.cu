// Hyper-parameter training on the device: negative log marginal likelihood and its analytic gradient.
// Replaces SFGP.likelihood / MFGP.likelihood (reference gaussian_process.py:81-105, :344-384) and the autograd
// differentiation behind SFGP.train / MFGP.train (:107-119, :386-399; value_and_grad at :118, :397):
//     NLML = 1/2 z^T z + sum log diag L + 1/2 N log 2 pi,     z = L^-1 (y - m),
//     dNLML/dh = 1/2 sum_ij (K^-1 - alpha alpha^T)_ij dK_ij/dh - alpha^T dm/dh,     alpha = K^-1 (y - m) = W^T z,
// on top of the fit the posterior path already has (mfgp_build_train_cov -> mfgp_cholesky -> mfgp_tri_inverse ->
// mfgp_whiten): K^-1 = W^T W is one DMMA tile product over the lower tiles, and the nine (four) gradient components come
// out of ONE coalesced sweep over the lower triangle that re-evaluates the kernel parts on the fly (fused pairwise
// distance + exp + fidelity-scale blocks, like build_train_cov_kernel) -- no dK/dh matrix is ever stored.
// Deterministic: per-block partial sums in a fixed order, then a single-block final pass.
#include "common.cuh"
#include "gemm_f64.cuh"

namespace mfgp {

constexpr int TR_THREADS = 256;
constexpr int TR_SUMS = 7;    // S1 .. S7 below

__global__ void transpose_kernel(const double* __restrict__ W, int64_t ldw, double* __restrict__ Wt, int64_t ldt, int n) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = by + r, j = bx + threadIdx.x;
        tile[r][threadIdx.x] = (i < n && j < n) ? W[(int64_t)i * ldw + j] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = bx + r, j = by + threadIdx.x;
        if (i < n && j < n) Wt[(int64_t)i * ldt + j] = tile[threadIdx.x][r];
    }
}

// alpha[i] = sum_{k >= i} Wt[i][k] z[k]  (Wt = W^T is upper triangular): one warp per row
__global__ void alpha_kernel(const double* __restrict__ Wt, int64_t ldt, const double* __restrict__ z, int npad,
                             double* __restrict__ alpha) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= npad) return;
    double s = 0.0;
    for (int k = row + lane; k < npad; k += 32) s += Wt[(int64_t)row * ldt + k] * z[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) alpha[row] = s;
}

// One sweep over the lower triangle (i >= j, both < N): q = Kinv_ij - alpha_i alpha_j, weight 2 off the diagonal.
//   S1 = sum q c kL        S2 = sum q c kL r^2/l_L^2     (c = 1 / rho / rho^2 in the LL / LH / HH block)
//   S3 = sum_HH q kH       S4 = sum_HH q kH r^2/l_H^2
//   S5 = sum_LH q rho kL + sum_HH q 2 rho^2 kL
//   S6 = tr_LL q           S7 = tr_HH q
// One CTA per 64-row block i0 and a strided set of 64-column blocks; partial[block][TR_SUMS].
__global__ void __launch_bounds__(TR_THREADS) nlml_grad_sweep_kernel(const double* __restrict__ Kinv, int64_t ldk,
                                                                     const double* __restrict__ alpha, const double* __restrict__ Tt,
                                                                     int NL, int NH, DevParams p, double* __restrict__ partial) {
    __shared__ double red[TR_THREADS / 32][TR_SUMS];
    const int N = NL + NH;
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double s[TR_SUMS];
#pragma unroll
    for (int k = 0; k < TR_SUMS; k++) s[k] = 0.0;
    if (j0 <= i0) {
        for (int e = tid; e < 64 * 64; e += TR_THREADS) {
            const int i = i0 + (e >> 6), j = j0 + (e & 63);        // consecutive threads: consecutive columns (coalesced)
            if (i >= N || j > i) continue;
            const double q = (Kinv[(int64_t)i * ldk + j] - alpha[i] * alpha[j]) * (i == j ? 1.0 : 2.0);
            const double4 ti = reinterpret_cast<const double4*>(Tt)[i], tj = reinterpret_cast<const double4*>(Tt)[j];
            const bool iL = i < NL, jL = j < NL;
            if (p.multi) {
                const double d0 = ti.x - tj.x, d1 = ti.y - tj.y;
                const double qL = __dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1));
                const double kL = p.s_L * exp(-0.5 * qL);
                const double c = (iL && jL) ? 1.0 : ((iL != jL) ? p.rho : p.rho2);
                s[0] += q * c * kL;
                s[1] += q * c * kL * qL;
                if (iL != jL) s[4] += q * p.rho * kL;
                if (!iL && !jL) s[4] += q * 2.0 * p.rho2 * kL;
            }
            if (!iL && !jL) {
                const double d0 = ti.z - tj.z, d1 = ti.w - tj.w;
                const double qH = __dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1));
                const double kH = p.s_H * exp(-0.5 * qH);
                s[2] += q * kH;
                s[3] += q * kH * qH;
            }
            if (i == j) {
                if (iL) s[5] += q; else s[6] += q;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < TR_SUMS; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if (lane == 0) red[warp][k] = s[k];
    }
    __syncthreads();
    if (tid < TR_SUMS) {
        double t = 0.0;
        for (int w = 0; w < TR_THREADS / 32; w++) t += red[w][tid];
        partial[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * TR_SUMS + tid] = t;
    }
}

// single block: add the partials in block order, the O(N) terms, and assemble out[0] = NLML, out[1 ..] = gradient in the
// reference's hyper-parameter order ([mu_lo, s^2_lo, L_lo, mu_hi, s^2_hi, L_hi, rho, noise_lo, noise_hi] or
// [mu, s^2, L, noise]); every hyper-parameter is log-scaled, so d/dh = value * d/dvalue.
__global__ void __launch_bounds__(TR_THREADS) nlml_grad_final_kernel(const double* __restrict__ partial, int nblocks,
                                                                     const double* __restrict__ Lf, int64_t ld,
                                                                     const double* __restrict__ z, const double* __restrict__ alpha,
                                                                     int NL, int NH, DevParams p, double* __restrict__ out) {
    __shared__ double red[TR_THREADS / 32][TR_SUMS + 4];
    __shared__ double tot[TR_SUMS + 4];
    const int N = NL + NH, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double s[TR_SUMS + 4];
#pragma unroll
    for (int k = 0; k < TR_SUMS + 4; k++) s[k] = 0.0;
    for (int b = tid; b < nblocks; b += TR_THREADS)
#pragma unroll
        for (int k = 0; k < TR_SUMS; k++) s[k] += partial[(int64_t)b * TR_SUMS + k];
    for (int i = tid; i < N; i += TR_THREADS) {
        s[TR_SUMS + 0] += log(Lf[(int64_t)i * (ld + 1)]);
        s[TR_SUMS + 1] += z[i] * z[i];
        if (i < NL) s[TR_SUMS + 2] += alpha[i]; else s[TR_SUMS + 3] += alpha[i];
    }
#pragma unroll
    for (int k = 0; k < TR_SUMS + 4; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if (lane == 0) red[warp][k] = s[k];
    }
    __syncthreads();
    if (tid < TR_SUMS + 4) {
        double t = 0.0;
        for (int w = 0; w < TR_THREADS / 32; w++) t += red[w][tid];
        tot[tid] = t;
    }
    __syncthreads();
    if (tid == 0) {
        const double logdet = tot[TR_SUMS], zz = tot[TR_SUMS + 1], aL = tot[TR_SUMS + 2], aH = tot[TR_SUMS + 3];
        out[0] = 0.5 * zz + logdet + 0.5 * log(2.0 * 3.14159265358979323846) * (double)N;
        if (p.multi) {
            const double eH = p.mean_H - p.rho * p.mean_L;       // exp(hyp[3])
            out[1] = -(p.mean_L * aL + p.rho * p.mean_L * aH);
            out[2] = 0.5 * tot[0];
            out[3] = 0.5 * tot[1];
            out[4] = -eH * aH;
            out[5] = 0.5 * tot[2];
            out[6] = 0.5 * tot[3];
            out[7] = 0.5 * tot[4] - p.rho * p.mean_L * aH;
            out[8] = 0.5 * p.noise_L * tot[5];
            out[9] = 0.5 * p.noise_H * tot[6];
        } else {
            out[1] = -p.mean_H * aH;
            out[2] = 0.5 * tot[2];
            out[3] = 0.5 * tot[3];
            out[4] = 0.5 * p.noise_H * tot[6];
#pragma unroll
            for (int k = 5; k < 10; k++) out[k] = 0.0;
        }
    }
}

}  // namespace mfgp

using namespace mfgp;

extern "C" int64_t mfgp_nlml_workspace_bytes(int64_t npad) {
    const int64_t nb = npad / MFGP_TILE;
    return (2 * npad * npad + npad + nb * nb * TR_SUMS) * 8 + 256;
}

extern "C" int mfgp_nlml_grad(const double* L, int64_t npad, int64_t ld, const double* W, int64_t ldw, const double* z,
                              const double* Tt, int64_t NL, int64_t NH, const mfgp_params* p_host, double* out, void* work,
                              int64_t work_bytes, void* stream) {
    if (!L || !W || !z || !Tt || !p_host || !out || !work || npad <= 0 || npad % MFGP_TILE || ld < npad || ldw < npad ||
        NL < 0 || NH < 0 || NL + NH <= 0 || NL + NH > npad || work_bytes < mfgp_nlml_workspace_bytes(npad))
        return MFGP_ERR_INVALID;
    if (!p_host->multi && NL != 0) return MFGP_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n = (int)npad, nb = n / MFGP_TILE;
    double* Wt = static_cast<double*>(work);
    double* Kinv = Wt + npad * npad;
    double* alpha = Kinv + npad * npad;
    double* partial = alpha + npad;
    {
        dim3 grid((n + 31) / 32, (n + 31) / 32), block(32, 8);
        transpose_kernel<<<grid, block, 0, st>>>(W, ldw, Wt, npad, n);
        MFGP_LAUNCH_CHECK();
    }
    GemmArgs g{};          // K^-1 = W^T W = Wt Wt^T, lower tiles; Wt is upper triangular: k >= m0 contributes
    g.A = Wt; g.lda = npad; g.B = Wt; g.ldb = npad; g.C = Kinv; g.ldc = npad;
    g.M = n; g.N = n; g.K = n; g.alpha = 1.0; g.beta = 0.0; g.mode = GEMM_SYRK_LOWER_AUPPER;
    int rc = launch_gemm(g, true, 1, st);
    if (rc) return rc;
    alpha_kernel<<<(n + 7) / 8, 256, 0, st>>>(Wt, npad, z, n, alpha);
    MFGP_LAUNCH_CHECK();
    const DevParams dp = make_dev_params(*p_host);
    nlml_grad_sweep_kernel<<<dim3(nb, nb), TR_THREADS, 0, st>>>(Kinv, npad, alpha, Tt, (int)NL, (int)NH, dp, partial);
    MFGP_LAUNCH_CHECK();
    nlml_grad_final_kernel<<<1, TR_THREADS, 0, st>>>(partial, nb * nb, L, ld, z, alpha, (int)NL, (int)NH, dp, out);
    MFGP_LAUNCH_CHECK();
    return MFGP_OK;
}
