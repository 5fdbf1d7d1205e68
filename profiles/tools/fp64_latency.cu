// Dependent-chain latencies of the FP64 operations on the critical path of the 64x64 diagonal-block factorisation
// (potrf_diag_body): cycles per operation, one warp, measured with clock64 around a 2048-long dependent chain.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_latency.cu -o fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 2048;
template <int OP>
__global__ void chain(double x0, double y0, double* out, long long* cyc) {
    __shared__ double sm[64];
    double x = x0, y = y0;
    sm[threadIdx.x & 63] = x0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = fma(x, y, y);
        if (OP == 1) x = x * y;
        if (OP == 2) x = rsqrt(x) + y;                 // CUDA double rsqrt (+ 1 add)
        if (OP == 3) x = 1.0 / x + y;                  // CUDA double divide (+ 1 add)
        if (OP == 4) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + y; }
        if (OP == 5) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + y; }
        if (OP == 6) { __syncthreads(); x += 1.0; }
        if (OP == 7) { x = sm[(int)x & 63] ; }         // dependent shared load (+ f2i)
        if (OP == 8) { sm[threadIdx.x & 63] = x; __syncthreads(); x = sm[(threadIdx.x + 1) & 63] + y; }   // publish -> barrier -> read
        if (OP == 9) {                                 // rcp.approx + 2 Newton steps (full double reciprocal)
            double q; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(x));
            double e = fma(-x, q, 1.0); q = fma(q, e, q);
            e = fma(-x, q, 1.0); q = fma(q, e, q);
            x = q + y;
        }
        if (OP == 10) x = sqrt(x) + y;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { *cyc = t1 - t0; }
    out[threadIdx.x] = x;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const char* names[] = {"dfma", "dmul", "rsqrt(double)+dadd", "1.0/x+dadd", "rsqrt.approx.f64+dadd", "rcp.approx.f64+dadd",
                           "__syncthreads+dadd", "dependent LDS (+f2i)", "STS->bar->LDS+dadd", "rcp.approx+2 Newton+dadd", "sqrt(double)+dadd"};
    for (int threads : {32, 256}) {
        for (int op = 0; op <= 10; op++) {
            long long h = 0;
            for (int rep = 0; rep < 2; rep++) {
                switch (op) {
                    case 0: chain<0><<<1, threads>>>(1.0000001, 1e-9, out, cyc); break;
                    case 1: chain<1><<<1, threads>>>(1.0000001, 0.99999999, out, cyc); break;
                    case 2: chain<2><<<1, threads>>>(1.3, 0.7, out, cyc); break;
                    case 3: chain<3><<<1, threads>>>(1.3, 0.7, out, cyc); break;
                    case 4: chain<4><<<1, threads>>>(1.3, 0.7, out, cyc); break;
                    case 5: chain<5><<<1, threads>>>(1.3, 0.7, out, cyc); break;
                    case 6: chain<6><<<1, threads>>>(1.3, 0.7, out, cyc); break;
                    case 7: chain<7><<<1, threads>>>(3.0, 0.7, out, cyc); break;
                    case 8: chain<8><<<1, threads>>>(1.3, 0.7, out, cyc); break;
                    case 9: chain<9><<<1, threads>>>(1.3, 0.7, out, cyc); break;
                    case 10: chain<10><<<1, threads>>>(1.3, 0.7, out, cyc); break;
                }
                cudaDeviceSynchronize();
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            printf("threads=%3d %-28s %7.1f clk/iter\n", threads, names[op], (double)h / N);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
