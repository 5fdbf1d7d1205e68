// Tour planner of the Choi algorithm: replaces compute_sample_tsp (reference simulator.py:415-454), which orders every
// agent's sample points with mlrose's genetic algorithm (TSPOpt + genetic_alg(mutation_prob=0.2, max_attempts=100,
// random_state=2), :435-438).  mlrose is an unpinned third-party host routine whose tour depends on its RNG consumption;
// the replacement minimises the same objective (length of the closed tour) deterministically:
//   nearest-neighbour construction from point 0 of the cluster (ties -> lowest index), then best-improvement 2-opt with
//   position 0 fixed (all segment reversals tour[i..j], 1 <= i < j <= n-1; ties -> lowest i, then j) until no reversal
//   gains more than TSP_IMPROVE_TOL.
// One CTA per cluster; all agents of a Choi period in ONE launch.  Every move evaluates the O(n^2) reversals in parallel
// (two square roots each against the cached edge lengths), reduces to the best one with its tie rule, and reverses the
// segment cooperatively.  All arithmetic is single-rounding fp64 in a fixed order (__dmul_rn / __dadd_rn / __dsqrt_rn,
// no FMA contraction), so the tour equals the literal CPU statement oracle/tsp.py bit for bit.
#include <cfloat>

#include "common.cuh"

namespace mfgp {

constexpr int TSP_THREADS = 256;
constexpr int TSP_SMEM_POINTS = 4096;          // clusters up to this size live in shared memory, larger ones in `work`
constexpr double TSP_IMPROVE_TOL = 1e-12;

__device__ __forceinline__ double tsp_dist(double ax, double ay, double bx, double by) {
    const double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

struct TspBest {
    double d; int i, j;
};
__device__ __forceinline__ TspBest tsp_better(TspBest a, TspBest b) {      // most negative delta; ties: lowest (i, j)
    if (b.i < 0) return a;
    if (a.i < 0) return b;
    if (b.d < a.d) return b;
    if (a.d < b.d) return a;
    if (b.i < a.i || (b.i == a.i && b.j < a.j)) return b;
    return a;
}

// pts[off[c] .. off[c+1]) : the points of cluster c;  order[off[c] + k] : LOCAL index of the k-th point of the tour;
// moves[c]: number of 2-opt moves applied (diagnostics).  Per cluster scratch (tour-ordered x, y, edge lengths, used
// flags): shared memory when n <= TSP_SMEM_POINTS, else work[4 * off[c] ..].
__global__ void __launch_bounds__(TSP_THREADS) tsp_tours_kernel(const double* __restrict__ pts, const int32_t* __restrict__ off,
                                                                int32_t* __restrict__ order, int32_t* __restrict__ moves,
                                                                double* __restrict__ work) {
    extern __shared__ __align__(16) double tsp_smem[];
    __shared__ double rd[TSP_THREADS / 32];
    __shared__ int ri[TSP_THREADS / 32], rj[TSP_THREADS / 32];
    __shared__ TspBest chosen;
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int o0 = off[c], n = off[c + 1] - o0;
    if (n <= 0) return;
    const double* P = pts + 2 * (int64_t)o0;
    int32_t* ord = order + o0;
    double* base = (n <= TSP_SMEM_POINTS) ? tsp_smem : work + 4 * (int64_t)o0;
    double* tx = base;                 // coordinates in TOUR order
    double* ty = base + n;
    double* e = base + 2 * n;          // e[k] = |tour[k] tour[k+1 mod n]|
    double* used = base + 3 * n;       // construction only (0 / 1)

    // ---- nearest-neighbour construction from point 0 ----
    for (int k = tid; k < n; k += TSP_THREADS) used[k] = 0.0;
    if (tid == 0) { ord[0] = 0; used[0] = 1.0; tx[0] = P[0]; ty[0] = P[1]; }
    __syncthreads();
    for (int step = 1; step < n; step++) {
        const double cx = tx[step - 1], cy = ty[step - 1];
        TspBest b{DBL_MAX, -1, 0};
        for (int k = tid; k < n; k += TSP_THREADS) {
            if (used[k] != 0.0) continue;
            const double dx = __dsub_rn(cx, P[2 * k]), dy = __dsub_rn(cy, P[2 * k + 1]);
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            if (b.i < 0 || d2 < b.d) { b.d = d2; b.i = k; }          // ascending k per thread: first minimum kept
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            TspBest y{__shfl_xor_sync(0xffffffffu, b.d, o), __shfl_xor_sync(0xffffffffu, b.i, o), 0};
            b = tsp_better(b, y);
        }
        if (lane == 0) { rd[warp] = b.d; ri[warp] = b.i; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < TSP_THREADS / 32; w++) b = tsp_better(b, TspBest{rd[w], ri[w], 0});
            ord[step] = b.i; used[b.i] = 1.0; tx[step] = P[2 * b.i]; ty[step] = P[2 * b.i + 1];
        }
        __syncthreads();
    }

    // ---- best-improvement 2-opt, position 0 fixed ----
    int nmoves = 0;
    if (n >= 4) {
        const int max_moves = 20 * n + 100;
        const long long npairs = (long long)(n - 2) * (n - 1) / 2;      // (i, j), 1 <= i < j <= n-1, i <= n-2
        for (; nmoves < max_moves; nmoves++) {
            for (int k = tid; k < n; k += TSP_THREADS) {
                const int k1 = (k + 1 == n) ? 0 : k + 1;
                e[k] = tsp_dist(tx[k], ty[k], tx[k1], ty[k1]);
            }
            __syncthreads();
            TspBest b{-TSP_IMPROVE_TOL, -1, -1};
            // row i holds n-1-i pairs; thread t takes the pairs t, t + T, ... of the row-major enumeration
            int i = 1;
            long long row0 = 0;                                          // linear index of pair (i, i+1)
            for (long long q = tid; q < npairs; q += TSP_THREADS) {
                while (q - row0 >= n - 1 - i) { row0 += n - 1 - i; i++; }
                const int j = i + 1 + (int)(q - row0);
                const int j1 = (j + 1 == n) ? 0 : j + 1;
                const double dac = tsp_dist(tx[i - 1], ty[i - 1], tx[j], ty[j]);
                const double dbd = tsp_dist(tx[i], ty[i], tx[j1], ty[j1]);
                const double delta = __dsub_rn(__dadd_rn(dac, dbd), __dadd_rn(e[i - 1], e[j]));
                if (delta < b.d) { b.d = delta; b.i = i; b.j = j; }      // ascending (i, j) per thread: first minimum kept
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                TspBest y{__shfl_xor_sync(0xffffffffu, b.d, o), __shfl_xor_sync(0xffffffffu, b.i, o),
                          __shfl_xor_sync(0xffffffffu, b.j, o)};
                b = tsp_better(b, y);
            }
            if (lane == 0) { rd[warp] = b.d; ri[warp] = b.i; rj[warp] = b.j; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < TSP_THREADS / 32; w++) b = tsp_better(b, TspBest{rd[w], ri[w], rj[w]});
                chosen = b;
            }
            __syncthreads();
            const TspBest m = chosen;
            if (m.i < 0) break;
            const int half = (m.j - m.i + 1) / 2;
            for (int k = tid; k < half; k += TSP_THREADS) {              // reverse tour[i..j]
                const int a = m.i + k, z = m.j - k;
                const double x = tx[a], y = ty[a];
                tx[a] = tx[z]; ty[a] = ty[z]; tx[z] = x; ty[z] = y;
                const int32_t t = ord[a]; ord[a] = ord[z]; ord[z] = t;
            }
            __syncthreads();
        }
    }
    if (tid == 0 && moves != nullptr) moves[c] = nmoves;
}

}  // namespace mfgp

using namespace mfgp;

extern "C" int64_t choi_tsp_workspace_bytes(int64_t n_total) { return n_total * 4 * 8 + 256; }

extern "C" int choi_tsp_tours(const double* pts, const int32_t* off, int64_t A, int64_t n_total, int64_t n_max, int32_t* order,
                              int32_t* moves, void* work, int64_t work_bytes, void* stream) {
    if (A < 0 || n_total < 0 || n_max < 0 || n_max > n_total) return MFGP_ERR_INVALID;
    if (A == 0 || n_total == 0) return MFGP_OK;
    if (!pts || !off || !order) return MFGP_ERR_INVALID;
    if (n_max > TSP_SMEM_POINTS && (!work || work_bytes < choi_tsp_workspace_bytes(n_total))) return MFGP_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t np = n_max < TSP_SMEM_POINTS ? n_max : TSP_SMEM_POINTS;
    const int smem = (int)(np * 4 * sizeof(double));
    MFGP_CUDA_CHECK(cudaFuncSetAttribute(tsp_tours_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TSP_SMEM_POINTS * 4 * 8));
    tsp_tours_kernel<<<(unsigned)A, TSP_THREADS, smem, st>>>(pts, off, order, moves, static_cast<double*>(work));
    MFGP_LAUNCH_CHECK();
    return MFGP_OK;
}
