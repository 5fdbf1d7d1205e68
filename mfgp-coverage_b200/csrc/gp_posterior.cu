// Fused GP posterior over a tile of grid points (replaces SFGP.predict, reference gaussian_process.py:121-148, and
// MFGP.predict, :401-438, diagonal only -- the reference's callers take np.diag of the covariance).
//
//   v = W psi(x*)          W = L^-1 (lower triangular, from mfgp_tri_inverse), psi = cross-covariance column
//   mu  = mean_H + v . z   z = W (y - mean)
//   var = k(0)  - v . v
//
// One CTA owns BN grid points and walks the row blocks of W top to bottom.  Per K slab it stages a BM x BK tile of W
// with cp.async and GENERATES the BK x BN tile of psi on chip (exp on the FP64 pipe), multiplies them on the FP64
// tensor cores (DMMA 8x8x4) and, per row block, folds v into the two running column reductions.  psi and V never
// reach HBM (unless the caller asks for the V cache of the Choi planner).
//
// Why the tile is tall (BM = 512 rows x BN = 32 points): DMMA and DFMA share one datapath on sm_100a and an fp64 exp
// costs ~23 FP64 lane-slots (profiles/r01_fp64_pipes.log), so every regenerated psi element is paid in MMA time; a
// psi element is reused BM times per generation, i.e. the exp overhead is ~2*23/BM of the MMA work (9% at 512).
// Triangular structure: slabs right of the diagonal are never visited, and inside the diagonal square each warp
// skips the 8-row tiles that lie entirely above the diagonal (rows are interleaved over warps so this stays balanced).
#include "common.cuh"

namespace mfgp {

struct PostArgs {
    const double* Xs; int64_t G;
    const double* Tt; int NL, NH;
    const double* W; int npad; int64_t ldw;
    const double* z;
    double* mu; double* var;
    double* Vc; int64_t ldv;
    DevParams p;
};

template <int WARPS_M, int MI, int NI, int BK>
struct PostCfg {
    static constexpr int BM = WARPS_M * MI * 8;
    static constexpr int BN = NI * 8;
    static constexpr int THREADS = WARPS_M * 32;
    static constexpr int LDS = BK + 4;     // padded smem row in doubles (conflict-free fragment reads)
    static constexpr size_t smem_bytes() {
        return sizeof(double) * (2 * BM * LDS + 2 * BN * LDS + BN * 4 + WARPS_M * BN * 2);
    }
};

template <int WARPS_M, int MI, int NI, int BK>
__global__ void __launch_bounds__(WARPS_M * 32, 1) posterior_kernel(PostArgs a) {
    using Cfg = PostCfg<WARPS_M, MI, NI, BK>;
    constexpr int BM = Cfg::BM, BN = Cfg::BN, THREADS = Cfg::THREADS, LDS = Cfg::LDS;
    extern __shared__ __align__(16) double smem[];
    double* Ws = smem;                          // [2][BM][LDS]
    double* Ps = Ws + 2 * BM * LDS;             // [2][BN][LDS]   psi tile, point-major, k contiguous
    double* xs = Ps + 2 * BN * LDS;             // [BN][4]        grid coords / l_L, / l_H
    double* colacc = xs + BN * 4;               // [WARPS_M][BN][2]  per-warp running (sum v^2, sum v z)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int64_t g0 = (int64_t)blockIdx.x * BN;
    const int N = a.NL + a.NH;
    const DevParams& p = a.p;

    for (int e = tid; e < BN; e += THREADS) {
        int64_t g = g0 + e;
        if (g >= a.G) g = a.G - 1;
        const double x = a.Xs[2 * g], y = a.Xs[2 * g + 1];
        xs[e * 4 + 0] = x / p.l_L;
        xs[e * 4 + 1] = y / p.l_L;
        xs[e * 4 + 2] = x / p.l_H;
        xs[e * 4 + 3] = y / p.l_H;
    }
    for (int e = tid; e < WARPS_M * BN * 2; e += THREADS) colacc[e] = 0.0;
    __syncthreads();

    const int nrb = (a.npad + BM - 1) / BM;
    for (int rb = 0; rb < nrb; rb++) {
        const int row0 = rb * BM;
        const int kmax = min(a.npad, row0 + BM);
        const int nslab = kmax / BK;

        double acc[MI][NI][2];
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

        auto stage = [&](int buf, int k0) {
            // W tile: BM rows x BK doubles, 16-byte chunks; rows past npad are zero-filled
            double* wdst = Ws + buf * BM * LDS;
#pragma unroll 4
            for (int c = tid; c < BM * (BK / 2); c += THREADS) {
                const int r = c / (BK / 2), q = c % (BK / 2);
                const int row = row0 + r;
                const bool ok = row < a.npad;
                cp_async16(wdst + r * LDS + q * 2, a.W + (int64_t)(ok ? row : 0) * a.ldw + k0 + q * 2, ok);
            }
            cp_async_commit();
            // psi tile: element (point gi, training column k0 + k); gaussian_process.py:426-429 / :139
            double* pdst = Ps + buf * BN * LDS;
            for (int e = tid; e < BN * BK; e += THREADS) {
                const int k = e % BK, gi = e / BK;
                const int n = k0 + k;
                double v = 0.0;
                if (n < N) {
                    const double4 t = reinterpret_cast<const double4*>(a.Tt)[n];
                    if (p.multi) {
                        const double kL = rbf_scaled(xs[gi * 4 + 0], xs[gi * 4 + 1], t.x, t.y, p.s_L);
                        if (n < a.NL) {
                            v = p.rho * kL;
                        } else {
                            const double kH = rbf_scaled(xs[gi * 4 + 2], xs[gi * 4 + 3], t.z, t.w, p.s_H);
                            v = __dadd_rn(__dmul_rn(p.rho2, kL), kH);
                        }
                    } else {
                        v = rbf_scaled(xs[gi * 4 + 2], xs[gi * 4 + 3], t.z, t.w, p.s_H);
                    }
                }
                pdst[gi * LDS + k] = v;
            }
        };

        // last k (inclusive) at which the 8-row tile mi of this warp still has a non-zero W entry; -1 = tile unused
        int klim[MI];
#pragma unroll
        for (int i = 0; i < MI; i++) {
            const int r = row0 + (i * WARPS_M + warp) * 8;
            klim[i] = (r < a.npad) ? r + 7 : -1;
        }

        stage(0, 0);
        for (int s = 0; s < nslab; s++) {
            const int buf = s & 1;
            const int k0 = s * BK;
            if (s + 1 < nslab) {
                stage(buf ^ 1, k0 + BK);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const double* wsrc = Ws + buf * BM * LDS;
            const double* psrc = Ps + buf * BN * LDS;
            const bool full = (k0 + BK <= row0) && (row0 + BM <= a.npad);
            if (full) {
#pragma unroll
                for (int kk = 0; kk < BK; kk += 4) {
                    double af[MI], bf[NI];
#pragma unroll
                    for (int i = 0; i < MI; i++) af[i] = wsrc[((i * WARPS_M + warp) * 8 + gq) * LDS + kk + tq];
#pragma unroll
                    for (int j = 0; j < NI; j++) bf[j] = psrc[(j * 8 + gq) * LDS + kk + tq];
#pragma unroll
                    for (int i = 0; i < MI; i++)
#pragma unroll
                        for (int j = 0; j < NI; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                }
            } else {
#pragma unroll
                for (int kk = 0; kk < BK; kk += 4) {
                    double bf[NI];
#pragma unroll
                    for (int j = 0; j < NI; j++) bf[j] = psrc[(j * 8 + gq) * LDS + kk + tq];
#pragma unroll
                    for (int i = 0; i < MI; i++) {
                        if (k0 + kk <= klim[i]) {   // warp-uniform
                            const double af = wsrc[((i * WARPS_M + warp) * 8 + gq) * LDS + kk + tq];
#pragma unroll
                            for (int j = 0; j < NI; j++) dmma884(acc[i][j][0], acc[i][j][1], af, bf[j]);
                        }
                    }
                }
            }
            __syncthreads();
        }

        // fold this row block of V into the column reductions (and the V cache if requested)
        double sq[NI][2], dt[NI][2];
#pragma unroll
        for (int j = 0; j < NI; j++) sq[j][0] = sq[j][1] = dt[j][0] = dt[j][1] = 0.0;
#pragma unroll
        for (int i = 0; i < MI; i++) {
            const int row = row0 + (i * WARPS_M + warp) * 8 + gq;
            if (row < a.npad) {
                const double zr = a.z[row];
#pragma unroll
                for (int j = 0; j < NI; j++) {
                    const double v0 = acc[i][j][0], v1 = acc[i][j][1];
                    sq[j][0] += v0 * v0;
                    sq[j][1] += v1 * v1;
                    dt[j][0] += v0 * zr;
                    dt[j][1] += v1 * zr;
                    if (a.Vc != nullptr && row < N) {
                        const int64_t g = g0 + j * 8 + tq * 2;
                        double* dst = a.Vc + (int64_t)row * a.ldv + g;
                        if (g + 1 < a.G && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                            *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                        } else {
                            if (g < a.G) dst[0] = v0;
                            if (g + 1 < a.G) dst[1] = v1;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NI; j++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    sq[j][c] += __shfl_xor_sync(0xffffffffu, sq[j][c], o);
                    dt[j][c] += __shfl_xor_sync(0xffffffffu, dt[j][c], o);
                }
            }
        if (gq == 0) {
#pragma unroll
            for (int j = 0; j < NI; j++)
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    double* slot = colacc + (warp * BN + j * 8 + tq * 2 + c) * 2;
                    slot[0] += sq[j][c];
                    slot[1] += dt[j][c];
                }
        }
    }
    __syncthreads();
    for (int e = tid; e < BN; e += THREADS) {
        const int64_t g = g0 + e;
        if (g < a.G) {
            double s = 0.0, d = 0.0;
#pragma unroll
            for (int w = 0; w < WARPS_M; w++) {
                s += colacc[(w * BN + e) * 2 + 0];
                d += colacc[(w * BN + e) * 2 + 1];
            }
            a.var[g] = p.k0 - s;
            a.mu[g] = p.mean_H + d;
        }
    }
}

__global__ void posterior_prior_kernel(int64_t G, double mean, double k0, double* __restrict__ mu, double* __restrict__ var) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G) {
        mu[g] = mean;
        var[g] = k0;
    }
}

template <int WARPS_M, int MI, int NI, int BK>
int launch_posterior(const PostArgs& a, cudaStream_t st) {
    using Cfg = PostCfg<WARPS_M, MI, NI, BK>;
    auto kern = posterior_kernel<WARPS_M, MI, NI, BK>;
    MFGP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes()));
    const unsigned grid = (unsigned)((a.G + Cfg::BN - 1) / Cfg::BN);
    kern<<<grid, Cfg::THREADS, Cfg::smem_bytes(), st>>>(a);
    MFGP_LAUNCH_CHECK();
    return MFGP_OK;
}

}  // namespace mfgp

using namespace mfgp;

extern "C" int mfgp_posterior(const double* Xs, int64_t G, const double* Tt, int64_t NL, int64_t NH, const double* W,
                              int64_t npad, int64_t ldw, const double* z, const mfgp_params* p_host, double* mu,
                              double* var, double* Vc, int64_t ldv, void* stream) {
    if (!p_host || !mu || !var || G < 0 || NL < 0 || NH < 0) return MFGP_ERR_INVALID;
    if (G == 0) return MFGP_OK;
    if (!Xs) return MFGP_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const DevParams dp = make_dev_params(*p_host);
    const int64_t N = NL + NH;
    if (N == 0) {   // empty model: constant mean and prior variance (gaussian_process.py:139-146 with no data)
        posterior_prior_kernel<<<(unsigned)((G + 255) / 256), 256, 0, st>>>(G, dp.mean_H, dp.k0, mu, var);
        MFGP_LAUNCH_CHECK();
        return MFGP_OK;
    }
    if (!Tt || !W || !z || npad < N || npad % MFGP_TILE || ldw < npad) return MFGP_ERR_INVALID;
    if (!p_host->multi && NL != 0) return MFGP_ERR_INVALID;
    if (Vc && ldv < G) return MFGP_ERR_INVALID;
    PostArgs a;
    a.Xs = Xs; a.G = G; a.Tt = Tt; a.NL = (int)NL; a.NH = (int)NH; a.W = W; a.npad = (int)npad; a.ldw = ldw; a.z = z;
    a.mu = mu; a.var = var; a.Vc = Vc; a.ldv = ldv; a.p = dp;
    if (npad <= 128) return launch_posterior<4, 4, 4, 16>(a, st);     // BM = 128: small models (c1/c2 early iterations)
    if (npad <= 256) return launch_posterior<8, 4, 4, 16>(a, st);     // BM = 256
    return launch_posterior<8, 8, 4, 16>(a, st);                       // BM = 512, BN = 32
}
