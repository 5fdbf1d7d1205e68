// GP fit on the device: training covariance, blocked Cholesky, triangular inverse, whitened observations.
// Replaces SFGP.updt_info (reference gaussian_process.py:229-255) and MFGP.updt_info (:493-529).
#include "common.cuh"
#include "gemm_f64.cuh"

namespace mfgp {

// ---- K assembly ------------------------------------------------------------------------------------------------------
// K_LL = k_L + noise_L I, K_LH = rho k_L, K_HH = rho^2 k_L + k_H + noise_H I, then + jitter I (gaussian_process.py:
// 523-529); SF: k + noise I + jitter I (:253-254).  Padding rows/cols [N, npad) carry the identity.
__global__ void build_train_cov_kernel(const double* __restrict__ Xt, int NL, int NH, DevParams p, double* __restrict__ K,
                                       int npad, int64_t ld, double* __restrict__ Tt, int row_begin) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = row_begin + blockIdx.y * blockDim.y + threadIdx.y;
    const int N = NL + NH;
    if (i >= npad || j >= npad) return;
    double v;
    if (i >= N || j >= N) {
        v = (i == j) ? 1.0 : 0.0;
    } else {
        const double xi = Xt[2 * i], yi = Xt[2 * i + 1], xj = Xt[2 * j], yj = Xt[2 * j + 1];
        const bool iL = i < NL, jL = j < NL;
        if (p.multi) {
            const double kL = rbf_scaled(xi / p.l_L, yi / p.l_L, xj / p.l_L, yj / p.l_L, p.s_L);
            if (iL && jL) {
                v = kL;
                if (i == j) v = v + p.noise_L;
            } else if (iL != jL) {
                v = p.rho * kL;
            } else {
                const double kH = rbf_scaled(xi / p.l_H, yi / p.l_H, xj / p.l_H, yj / p.l_H, p.s_H);
                v = __dadd_rn(__dmul_rn(p.rho2, kL), kH);
                if (i == j) v = v + p.noise_H;
            }
        } else {
            v = rbf_scaled(xi / p.l_H, yi / p.l_H, xj / p.l_H, yj / p.l_H, p.s_H);
            if (i == j) v = v + p.noise_H;
        }
        if (i == j) v = v + p.jitter;
    }
    K[(int64_t)i * ld + j] = v;
    if (j == 0 && Tt != nullptr) {
        double4 t = make_double4(0.0, 0.0, 0.0, 0.0);
        if (i < N) {
            const double xi = Xt[2 * i], yi = Xt[2 * i + 1];
            t = make_double4(xi / p.l_L, yi / p.l_L, xi / p.l_H, yi / p.l_H);
        }
        reinterpret_cast<double4*>(Tt)[i] = t;
    }
}

// ---- 64x64 diagonal block: Cholesky + inverse in one CTA --------------------------------------------------------------
// 256 threads as a 16x16 grid; thread (ti, tc) keeps the 4x4 cyclic sub-block S[ti+16a][tc+16b] in REGISTERS.
// Factor: right-looking, ONE barrier per column and no divide / sqrt on the critical path -- column j is left
// UNSCALED (u = L sqrt(p_j)); its owners publish it to a double-buffered shared column, every thread takes
// r = rsqrt(p_j) and applies a[i][c] -= u_i u_c r^2 to its 16 registers; L = u r is applied once at the end (r_j is also
// 1 / L[j][j], so the inverse needs no divide).  The j loop is 4 (unrolled: static register index) x 16 (rolled).
// Inverse: block doubling inside the CTA, X21 = -X22 (L21 X11) for block sizes 8 -> 16 -> 32, all pairs of a level in
// parallel, the product L21 X11 parked in the (finally zero) upper-right block of X.
constexpr int PB = 64;
constexpr int PLD = PB + 1;

__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(double* __restrict__ A, int64_t ld, double* __restrict__ Winv,
                                                         int64_t ldw, int32_t* __restrict__ info, int jblk) {
    extern __shared__ __align__(16) double potrf_smem[];   // 2 x 64x65 doubles: above the 48 KB static limit
    double* S = potrf_smem;
    double* X = potrf_smem + PB * PLD;
    __shared__ double rs[PB];        // 1/sqrt(pivot) == 1/L[j][j]
    __shared__ double colbuf[2][PB];
    __shared__ int bad;
    const int tid = threadIdx.x;
    const int ti = tid >> 4, tc = tid & 15;
    if (tid == 0) bad = 0;
    double s[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int r = ti + 16 * a, c = tc + 16 * b;
            s[a][b] = (c <= r) ? A[(int64_t)r * ld + c] : 0.0;
        }
    for (int e = tid; e < PB * PB; e += 256) X[(e >> 6) * PLD + (e & 63)] = 0.0;
    __syncthreads();
#pragma unroll
    for (int jb = 0; jb < 4; jb++) {
#pragma unroll 1
        for (int jj = 0; jj < 16; jj++) {
            const int j = jb * 16 + jj;
            double* col = colbuf[j & 1];
            if (tc == jj) {          // owners of column j publish it (rows above j are never read)
#pragma unroll
                for (int a = 0; a < 4; a++) col[ti + 16 * a] = s[a][jb];
            }
            __syncthreads();
            double piv = col[j];
            if (!(piv > 0.0)) {     // uniform: np.linalg.cholesky raises here (gaussian_process.py:254 / :529)
                if (tid == 0 && !bad) {
                    bad = 1;
                    atomicCAS(info, 0, jblk * PB + j + 1);
                }
                piv = 1.0;
            }
            const double r = rsqrt(piv);
            const double ip = r * r;
            if (tid == 0) rs[j] = r;
            double ui[4], uc[4];
#pragma unroll
            for (int a = 0; a < 4; a++) ui[a] = col[ti + 16 * a] * ip;
#pragma unroll
            for (int b = 0; b < 4; b++) uc[b] = col[tc + 16 * b];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                if (b < jb) continue;                       // columns of earlier 16-groups are final
                const bool live = tc + 16 * b > j;          // only columns right of j change
#pragma unroll
                for (int a = 0; a < 4; a++)
                    if (live) s[a][b] = fma(-ui[a], uc[b], s[a][b]);
            }
        }
    }
    __syncthreads();
    const bool failed = bad != 0;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int r = ti + 16 * a, c = tc + 16 * b;
            // on failure leave a harmless identity so later kernels stay finite; the host raises on `info`
            const double v = failed ? ((r == c) ? 1.0 : 0.0) : ((c <= r) ? s[a][b] * rs[c] : 0.0);
            S[r * PLD + c] = v;
            if (c <= r) A[(int64_t)r * ld + c] = v;
        }
    __syncthreads();
    if (failed && tid < PB) rs[tid] = 1.0;
    __syncthreads();
    if (Winv == nullptr) return;
    // level 0: the eight 8x8 diagonal blocks, one warp each, lane c < 8 solves column c by forward substitution
    {
        const int b0 = (tid >> 5) * 8, c = tid & 31;
        if (c < 8) {
            for (int i = c; i < 8; i++) {
                double s = (i == c) ? 1.0 : 0.0;
                for (int k = c; k < i; k++) s = fma(-S[(b0 + i) * PLD + b0 + k], X[(b0 + k) * PLD + b0 + c], s);
                X[(b0 + i) * PLD + b0 + c] = s * rs[b0 + i];
            }
        }
    }
    for (int bs = 8, sh = 3; bs < PB; bs <<= 1, sh++) {
        const int total = 32 * bs;                  // (64 / 2bs) pairs x bs^2 elements
        __syncthreads();
        for (int e = tid; e < total; e += 256) {    // T = L21 X11 (X11 lower triangular: k >= c)
            const int t = e >> (2 * sh), rem = e & (bs * bs - 1), i = rem >> sh, c = rem & (bs - 1);
            const int top = 2 * bs * t;
            double s = 0.0;
            for (int k = c; k < bs; k++) s = fma(S[(top + bs + i) * PLD + top + k], X[(top + k) * PLD + top + c], s);
            X[(top + i) * PLD + top + bs + c] = s;
        }
        __syncthreads();
        for (int e = tid; e < total; e += 256) {    // X21 = -X22 T (X22 lower triangular: k <= i)
            const int t = e >> (2 * sh), rem = e & (bs * bs - 1), i = rem >> sh, c = rem & (bs - 1);
            const int top = 2 * bs * t;
            double s = 0.0;
            for (int k = 0; k <= i; k++) s = fma(X[(top + bs + i) * PLD + top + bs + k], X[(top + k) * PLD + top + bs + c], s);
            X[(top + bs + i) * PLD + top + c] = -s;
        }
        __syncthreads();
        for (int e = tid; e < total; e += 256) {    // clear the parked product
            const int t = e >> (2 * sh), rem = e & (bs * bs - 1), i = rem >> sh, c = rem & (bs - 1);
            X[(2 * bs * t + i) * PLD + 2 * bs * t + bs + c] = 0.0;
        }
    }
    __syncthreads();
    for (int e = tid; e < PB * PB; e += 256) {
        const int r = e >> 6, c = e & 63;
        Winv[(int64_t)r * ldw + c] = X[r * PLD + c];
    }
}

__global__ void zero_offdiag_blocks_kernel(double* __restrict__ W, int npad, int64_t ldw) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= npad) return;
    if ((i >> 6) != (j >> 6)) W[(int64_t)i * ldw + j] = 0.0;
}

__global__ void zero_cols_kernel(double* __restrict__ Wcol, int64_t ldw) { Wcol[(int64_t)blockIdx.y * ldw + threadIdx.x] = 0.0; }

// z = W (y - mean): one warp per row; W is lower triangular so only k <= row contributes
__global__ void whiten_kernel(const double* __restrict__ W, int npad, int64_t ldw, const double* __restrict__ y, int NL,
                              int NH, double mean_L, double mean_H, double* __restrict__ z, int row_begin) {
    const int row = row_begin + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= npad) return;
    const int N = NL + NH;
    const int kmax = min(row + 1, N);
    double s = 0.0;
    for (int k = lane; k < kmax; k += 32) {
        const double yc = y[k] - (k < NL ? mean_L : mean_H);
        s += W[(int64_t)row * ldw + k] * yc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) z[row] = s;
}

}  // namespace mfgp

using namespace mfgp;

extern "C" int64_t mfgp_npad(int64_t n) {
    if (n <= 0) return MFGP_TILE;
    return (n + MFGP_TILE - 1) / MFGP_TILE * MFGP_TILE;
}

extern "C" int64_t mfgp_workspace_bytes(int64_t npad) {
    return npad * npad * 2 + (int64_t)PB * PB * 8 + 256;   // T blocks of the inverse (npad^2/4 doubles) + one 64x64 block
}

extern "C" int mfgp_build_train_cov(const double* Xt, int64_t NL, int64_t NH, const mfgp_params* p_host, double* K,
                                    int64_t npad, int64_t ld, double* Tt, void* stream) {
    if (!p_host || !K || NL < 0 || NH < 0 || npad < NL + NH || npad % MFGP_TILE || ld < npad) return MFGP_ERR_INVALID;
    if (NL + NH > 0 && !Xt) return MFGP_ERR_INVALID;
    if (!p_host->multi && NL != 0) return MFGP_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 block(32, 8), grid((unsigned)((npad + 31) / 32), (unsigned)((npad + 7) / 8));
    build_train_cov_kernel<<<grid, block, 0, st>>>(Xt, (int)NL, (int)NH, make_dev_params(*p_host), K, (int)npad, ld, Tt, 0);
    MFGP_LAUNCH_CHECK();
    return MFGP_OK;
}

extern "C" int mfgp_cholesky(double* K, int64_t npad, int64_t ld, double* W, int64_t ldw, int32_t* info, void* work,
                             void* stream) {
    if (!K || !info || npad <= 0 || npad % MFGP_TILE || ld < npad) return MFGP_ERR_INVALID;
    if (!W && !work) return MFGP_ERR_INVALID;
    if (W && ldw < npad) return MFGP_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MFGP_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
    const int nb = (int)(npad / PB);
    constexpr int POTRF_SMEM = 2 * PB * PLD * sizeof(double);
    MFGP_CUDA_CHECK(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_SMEM));
    for (int j = 0; j < nb; j++) {
        double* Ajj = K + (int64_t)j * PB * (ld + 1);
        double* Wjj = W ? W + (int64_t)j * PB * (ldw + 1) : static_cast<double*>(work);
        const int64_t ldi = W ? ldw : PB;
        potrf_diag_kernel<<<1, 256, POTRF_SMEM, st>>>(Ajj, ld, Wjj, ldi, info, j);
        MFGP_LAUNCH_CHECK();
        const int rem = (int)(npad - (int64_t)(j + 1) * PB);
        if (rem <= 0) break;
        double* A21 = K + (int64_t)(j + 1) * PB * ld + (int64_t)j * PB;
        GemmArgs t{};   // L21 = A21 * inv(L11)^T, in place
        t.A = A21; t.lda = ld; t.B = Wjj; t.ldb = ldi; t.C = A21; t.ldc = ld;
        t.M = rem; t.N = PB; t.K = PB; t.alpha = 1.0; t.beta = 0.0; t.mode = GEMM_GENERAL;
        int rc = launch_gemm(t, true, 1, st);
        if (rc) return rc;
        GemmArgs s{};   // A22 -= L21 L21^T (lower tiles)
        s.A = A21; s.lda = ld; s.B = A21; s.ldb = ld; s.C = K + (int64_t)(j + 1) * PB * (ld + 1); s.ldc = ld;
        s.M = rem; s.N = rem; s.K = PB; s.alpha = -1.0; s.beta = 1.0; s.mode = GEMM_SYRK_LOWER;
        rc = launch_gemm(s, true, 1, st);
        if (rc) return rc;
    }
    return MFGP_OK;
}

// Cholesky of K fused with the forward substitution of R right-hand sides: on return K holds L (lower), the diagonal
// blocks of W hold the inverses of L's diagonal blocks, and Bm[npad, R] holds L^-1 Bm.  The panel chain (potrf -> panel solve
// -> trailing update: three small dependent kernels per 64 columns) leaves most of the GPU idle, so the right-hand-side
// work -- Y_j = W_jj B_j and B_{>j} -= L_{>j,j} Y_j, one pair of tile GEMMs per panel, N^2 R / 2 MACs in total -- runs on
// the caller's stream while the chain itself runs on an internal high-priority stream, and hides behind it: the explicit inverse (mfgp_tri_inverse) and the product W B are not needed
// by the factored posterior.  Cross-stream order: Y_j waits for potrf(j), the update waits for the panel solve of panel j;
// the caller's stream waits for the side stream before the call returns control of Bm.
namespace {
struct SideStream {
    cudaStream_t st = nullptr;
    cudaEvent_t ev_potrf[2] = {nullptr, nullptr}, ev_trsm[2] = {nullptr, nullptr}, ev_begin = nullptr, ev_end = nullptr;
    int device = -1;
};
SideStream g_side[16];

int side_for_current_device(SideStream** out) {
    int dev = 0;
    MFGP_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) return MFGP_ERR_INVALID;
    SideStream& s = g_side[dev];
    if (!s.st) {
        // the latency-bound panel chain runs on this stream at the HIGHEST priority, so that its small kernels are placed
        // ahead of the pending CTAs of the bulk right-hand-side GEMMs (which stay on the caller's stream)
        int least = 0, greatest = 0;
        MFGP_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        MFGP_CUDA_CHECK(cudaStreamCreateWithPriority(&s.st, cudaStreamNonBlocking, greatest));
        for (int i = 0; i < 2; i++) {
            MFGP_CUDA_CHECK(cudaEventCreateWithFlags(&s.ev_potrf[i], cudaEventDisableTiming));
            MFGP_CUDA_CHECK(cudaEventCreateWithFlags(&s.ev_trsm[i], cudaEventDisableTiming));
        }
        MFGP_CUDA_CHECK(cudaEventCreateWithFlags(&s.ev_begin, cudaEventDisableTiming));
        MFGP_CUDA_CHECK(cudaEventCreateWithFlags(&s.ev_end, cudaEventDisableTiming));
        s.device = dev;
    }
    *out = &s;
    return MFGP_OK;
}
}  // namespace

extern "C" int mfgp_cholesky_solve(double* K, int64_t npad, int64_t ld, double* W, int64_t ldw, int32_t* info, double* Bm,
                                   int64_t ldb, int64_t R, void* stream) {
    if (!K || !W || !info || !Bm || npad <= 0 || npad % MFGP_TILE || ld < npad || ldw < npad || R <= 0 || R % GT || ldb < R)
        return MFGP_ERR_INVALID;
    cudaStream_t caller = static_cast<cudaStream_t>(stream);
    SideStream* side = nullptr;
    int rcs = side_for_current_device(&side);
    if (rcs) return rcs;
    cudaStream_t st = side->st;        // panel chain: internal high-priority stream
    cudaStream_t sb = caller;          // right-hand-side GEMMs: the caller's stream
    MFGP_CUDA_CHECK(cudaEventRecord(side->ev_begin, caller));         // Bm and K were produced on the caller's stream
    MFGP_CUDA_CHECK(cudaStreamWaitEvent(st, side->ev_begin, 0));
    MFGP_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
    const int nb = (int)(npad / PB);
    constexpr int POTRF_SMEM = 2 * PB * PLD * sizeof(double);
    MFGP_CUDA_CHECK(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_SMEM));
    for (int j = 0; j < nb; j++) {
        double* Ajj = K + (int64_t)j * PB * (ld + 1);
        double* Wjj = W + (int64_t)j * PB * (ldw + 1);
        potrf_diag_kernel<<<1, 256, POTRF_SMEM, st>>>(Ajj, ld, Wjj, ldw, info, j);
        MFGP_LAUNCH_CHECK();
        MFGP_CUDA_CHECK(cudaEventRecord(side->ev_potrf[j & 1], st));
        const int rem = (int)(npad - (int64_t)(j + 1) * PB);
        double* A21 = K + (int64_t)(j + 1) * PB * ld + (int64_t)j * PB;
        if (rem > 0) {
            GemmArgs t{};   // L21 = A21 * inv(L11)^T, in place
            t.A = A21; t.lda = ld; t.B = Wjj; t.ldb = ldw; t.C = A21; t.ldc = ld;
            t.M = rem; t.N = PB; t.K = PB; t.alpha = 1.0; t.beta = 0.0; t.mode = GEMM_GENERAL;
            int rc = launch_gemm(t, true, 1, st);
            if (rc) return rc;
            MFGP_CUDA_CHECK(cudaEventRecord(side->ev_trsm[j & 1], st));
        }
        // side stream: Y_j = W_jj B_j (in place: every CTA reads only the column tile it writes)
        double* Bj = Bm + (int64_t)j * PB * ldb;
        MFGP_CUDA_CHECK(cudaStreamWaitEvent(sb, side->ev_potrf[j & 1], 0));
        GemmArgs y{};
        y.A = Wjj; y.lda = ldw; y.B = Bj; y.ldb = ldb; y.C = Bj; y.ldc = ldb;
        y.M = PB; y.N = (int)R; y.K = PB; y.alpha = 1.0; y.beta = 0.0; y.mode = GEMM_GENERAL;
        int rc = launch_gemm(y, false, 1, sb);
        if (rc) return rc;
        if (rem <= 0) break;
        MFGP_CUDA_CHECK(cudaStreamWaitEvent(sb, side->ev_trsm[j & 1], 0));
        GemmArgs u{};       // B_{>j} -= L_{>j,j} Y_j
        u.A = A21; u.lda = ld; u.B = Bj; u.ldb = ldb; u.C = Bm + (int64_t)(j + 1) * PB * ldb; u.ldc = ldb;
        u.M = rem; u.N = (int)R; u.K = PB; u.alpha = -1.0; u.beta = 1.0; u.mode = GEMM_GENERAL;
        rc = launch_gemm(u, false, 1, sb);
        if (rc) return rc;
        GemmArgs s2{};  // A22 -= L21 L21^T (lower tiles)
        s2.A = A21; s2.lda = ld; s2.B = A21; s2.ldb = ld; s2.C = K + (int64_t)(j + 1) * PB * (ld + 1); s2.ldc = ld;
        s2.M = rem; s2.N = rem; s2.K = PB; s2.alpha = -1.0; s2.beta = 1.0; s2.mode = GEMM_SYRK_LOWER;
        rc = launch_gemm(s2, true, 1, st);
        if (rc) return rc;
    }
    MFGP_CUDA_CHECK(cudaEventRecord(side->ev_end, st));               // the caller's stream resumes when the chain is done too
    MFGP_CUDA_CHECK(cudaStreamWaitEvent(caller, side->ev_end, 0));
    return MFGP_OK;
}

extern "C" int mfgp_tri_inverse(const double* L, int64_t npad, int64_t ld, double* W, int64_t ldw, void* work,
                                void* stream) {
    if (!L || !W || !work || npad <= 0 || npad % MFGP_TILE || ld < npad || ldw < npad) return MFGP_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* T = static_cast<double*>(work);
    {
        dim3 grid((unsigned)((npad + 255) / 256), (unsigned)npad);
        zero_offdiag_blocks_kernel<<<grid, 256, 0, st>>>(W, (int)npad, ldw);
        MFGP_LAUNCH_CHECK();
    }
    // block doubling: [[W11,0],[W21,W22]] with W21 = -W22 (L21 W11); all pairs of one level are independent
    for (int64_t b = PB; b < npad; b *= 2) {
        const int nfull = (int)(npad / (2 * b));
        const int64_t rem = npad - 2 * b * nfull;
        for (int pass = 0; pass < 2; pass++) {
            int batch;
            int64_t top, h;
            if (pass == 0) { batch = nfull; top = 0; h = b; }
            else { batch = (rem > b) ? 1 : 0; top = 2 * b * nfull; h = rem - b; }
            if (batch == 0) continue;
            GemmArgs g1{};   // T = L21 * W11   (W11 lower: k >= n0)
            g1.A = L + (top + b) * ld + top; g1.lda = ld; g1.strideA = 2 * b * (ld + 1);
            g1.B = W + top * (ldw + 1); g1.ldb = ldw; g1.strideB = 2 * b * (ldw + 1);
            g1.C = T; g1.ldc = b; g1.strideC = b * b;
            g1.M = (int)h; g1.N = (int)b; g1.K = (int)b; g1.alpha = 1.0; g1.beta = 0.0; g1.mode = GEMM_B_LOWER;
            int rc = launch_gemm(g1, false, batch, st);
            if (rc) return rc;
            GemmArgs g2{};   // W21 = -W22 * T  (W22 lower: k < m0 + 64)
            g2.A = W + (top + b) * (ldw + 1); g2.lda = ldw; g2.strideA = 2 * b * (ldw + 1);
            g2.B = T; g2.ldb = b; g2.strideB = b * b;
            g2.C = W + (top + b) * ldw + top; g2.ldc = ldw; g2.strideC = 2 * b * (ldw + 1);
            g2.M = (int)h; g2.N = (int)b; g2.K = (int)h; g2.alpha = -1.0; g2.beta = 0.0; g2.mode = GEMM_A_LOWER;
            rc = launch_gemm(g2, false, batch, st);
            if (rc) return rc;
        }
    }
    return MFGP_OK;
}

extern "C" int mfgp_whiten(const double* W, int64_t npad, int64_t ldw, const double* y, int64_t NL, int64_t NH,
                           const mfgp_params* p_host, double* z, void* stream) {
    if (!W || !z || !p_host || npad <= 0 || npad % MFGP_TILE || ldw < npad || NL + NH > npad) return MFGP_ERR_INVALID;
    if (NL + NH > 0 && !y) return MFGP_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int wpb = 8;
    whiten_kernel<<<(unsigned)((npad + wpb - 1) / wpb), wpb * 32, 0, st>>>(W, (int)npad, ldw, y, (int)NL, (int)NH,
                                                                          p_host->mean_L, p_host->mean_H, z, 0);
    MFGP_LAUNCH_CHECK();
    return MFGP_OK;
}

// Bordered (append-only) update of the factor: the reference appends new samples at the END of [X_L; X_H]
// (gaussian_process.py:266-268, :540-542) and refactors from scratch; algebraically the leading block of L is unchanged,
// so only the block rows that hold new points are recomputed, one 64-row block at a time (left-looking):
//   K_b   = covariance rows of the block                       (build_train_cov_kernel, rows [rb, rb+64))
//   L_b,left = K_b[:, 0:rb] W11^T                              (W11 = inverse of the already-factored leading block)
//   S     = K_b[:, rb:rb+64] - L_b,left L_b,left^T ;  L_bb = chol(S), W_bb = L_bb^-1      (potrf_diag_kernel)
//   W_b,left = -W_bb (L_b,left W11)
//   z_b   = W_b,: (y - mean)
// The skinny products (64 rows, K up to N) are split over K so they fill the GPU; partial tiles are added in fixed order.
extern "C" int mfgp_cholesky_append(const double* Xt, int64_t NL, int64_t NH_old, int64_t NH_new, const mfgp_params* p_host,
                                    double* K, int64_t ld, double* W, int64_t ldw, const double* y, double* z, double* Tt,
                                    int32_t* info, void* work, int64_t work_bytes, void* stream) {
    if (!Xt || !p_host || !K || !W || !y || !z || !info || !work || NL < 0 || NH_old < 0 || NH_new < NH_old) return MFGP_ERR_INVALID;
    const int64_t N_old = NL + NH_old, N_new = NL + NH_new;
    if (N_new == N_old) return MFGP_OK;
    const int64_t npad = mfgp_npad(N_new);
    if (ld < npad || ldw < npad) return MFGP_ERR_INVALID;
    if (!p_host->multi && NL != 0) return MFGP_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const DevParams dp = make_dev_params(*p_host);
    MFGP_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
    constexpr int POTRF_SMEM = 2 * PB * PLD * sizeof(double);
    MFGP_CUDA_CHECK(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_SMEM));
    double* part = static_cast<double*>(work);
    for (int64_t rb = N_old / PB * PB; rb < npad; rb += PB) {
        {   // covariance rows of the block: columns [0, rb+64) are used, the rest of the row is rewritten too (harmless)
            dim3 block(32, 8), grid((unsigned)((rb + PB + 31) / 32), PB / 8);
            build_train_cov_kernel<<<grid, block, 0, st>>>(Xt, (int)NL, (int)NH_new, dp, K, (int)(rb + PB), ld, Tt, (int)rb);
            MFGP_LAUNCH_CHECK();
        }
        if (rb > 0) {     // the column block above the new diagonal block of W must read as zero (upper triangle)
            dim3 zgrid(1, (unsigned)rb);
            zero_cols_kernel<<<zgrid, PB, 0, st>>>(W + rb, ldw);
            MFGP_LAUNCH_CHECK();
        }
        double* Kb = K + rb * ld;              // block rows of K / L
        double* Wb = W + rb * ldw;
        double* Kbb = Kb + rb;
        double* Wbb = Wb + rb;
        if (rb > 0) {
            int nsplit = (int)(rb / 256);
            if (nsplit < 1) nsplit = 1;
            if (nsplit > 16) nsplit = 16;
            int kchunk = (int)((rb / PB + nsplit - 1) / nsplit) * PB;
            nsplit = (int)((rb + kchunk - 1) / kchunk);
            const int64_t pstride = (int64_t)PB * rb;
            if (work_bytes < (int64_t)nsplit * pstride * 8) return MFGP_ERR_INVALID;
            dim3 rgrid((unsigned)((rb + 127) / 128), PB);
            GemmArgs g1{};      // part[z] = K_b[:, 0:rb] * W11^T over the z-th k range (W11 lower: k < n0 + 64)
            g1.A = Kb; g1.lda = ld; g1.B = W; g1.ldb = ldw; g1.C = part; g1.ldc = rb; g1.strideC = pstride;
            g1.M = PB; g1.N = (int)rb; g1.K = (int)rb; g1.alpha = 1.0; g1.beta = 0.0; g1.mode = GEMM_BT_LOWER; g1.kchunk = kchunk;
            int rc = launch_gemm(g1, true, nsplit, st);
            if (rc) return rc;
            splitk_reduce_kernel<<<rgrid, 128, 0, st>>>(part, nsplit, pstride, (int)rb, Kb, ld, PB, (int)rb, 1.0, 0.0);   // L_b,left
            MFGP_LAUNCH_CHECK();
            GemmArgs g2{};      // part[z] = L_b,left L_b,left^T over the z-th k range
            g2.A = Kb; g2.lda = ld; g2.B = Kb; g2.ldb = ld; g2.C = part; g2.ldc = PB; g2.strideC = PB * PB;
            g2.M = PB; g2.N = PB; g2.K = (int)rb; g2.alpha = 1.0; g2.beta = 0.0; g2.mode = GEMM_GENERAL; g2.kchunk = kchunk;
            rc = launch_gemm(g2, true, nsplit, st);
            if (rc) return rc;
            splitk_reduce_kernel<<<dim3(1, PB), 64, 0, st>>>(part, nsplit, PB * PB, PB, Kbb, ld, PB, PB, -1.0, 1.0);        // S
            MFGP_LAUNCH_CHECK();
        }
        potrf_diag_kernel<<<1, 256, POTRF_SMEM, st>>>(Kbb, ld, Wbb, ldw, info, (int)(rb / PB));
        MFGP_LAUNCH_CHECK();
        if (rb > 0) {
            int nsplit = (int)(rb / 256);
            if (nsplit < 1) nsplit = 1;
            if (nsplit > 16) nsplit = 16;
            int kchunk = (int)((rb / PB + nsplit - 1) / nsplit) * PB;
            nsplit = (int)((rb + kchunk - 1) / kchunk);
            const int64_t pstride = (int64_t)PB * rb;
            double* T2 = part + (int64_t)nsplit * pstride;      // reduced L_b,left W11 lives behind the partials
            if (work_bytes < ((int64_t)nsplit + 1) * pstride * 8) return MFGP_ERR_INVALID;
            GemmArgs g3{};      // part[z] = L_b,left * W11 over the z-th k range (W11 lower: k >= n0)
            g3.A = Kb; g3.lda = ld; g3.B = W; g3.ldb = ldw; g3.C = part; g3.ldc = rb; g3.strideC = pstride;
            g3.M = PB; g3.N = (int)rb; g3.K = (int)rb; g3.alpha = 1.0; g3.beta = 0.0; g3.mode = GEMM_B_LOWER; g3.kchunk = kchunk;
            int rc = launch_gemm(g3, false, nsplit, st);
            if (rc) return rc;
            dim3 rgrid((unsigned)((rb + 127) / 128), PB);
            splitk_reduce_kernel<<<rgrid, 128, 0, st>>>(part, nsplit, pstride, (int)rb, T2, rb, PB, (int)rb, 1.0, 0.0);
            MFGP_LAUNCH_CHECK();
            GemmArgs g4{};      // W_b,left = -W_bb * T2
            g4.A = Wbb; g4.lda = ldw; g4.B = T2; g4.ldb = rb; g4.C = Wb; g4.ldc = ldw;
            g4.M = PB; g4.N = (int)rb; g4.K = PB; g4.alpha = -1.0; g4.beta = 0.0; g4.mode = GEMM_GENERAL;
            rc = launch_gemm(g4, false, 1, st);
            if (rc) return rc;
        }
        whiten_kernel<<<PB / 8, 256, 0, st>>>(W, (int)(rb + PB), ldw, y, (int)NL, (int)NH_new, p_host->mean_L, p_host->mean_H, z, (int)rb);
        MFGP_LAUNCH_CHECK();
    }
    return MFGP_OK;
}

extern "C" int64_t mfgp_append_workspace_bytes(int64_t npad) { return (int64_t)17 * PB * npad * 8 + 256; }
