// FP64 pipe microbenchmark for sm_100a: what does the posterior kernel have to work with?
//   (1) DMMA m8n8k4 / m16n8k8 / m16n8k16 issue rate (register-resident operands)
//   (2) DFMA issue rate
//   (3) both in one kernel (do the tensor and vector FP64 paths overlap or share a datapath?)
//   (4) fp64 exp() rate (the kernel-matrix builder's unit of work)
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_dmma884(double* out, int iters, double seed) {
    double acc[NACC][2];
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
#pragma unroll
    for (int i = 0; i < NACC; i++) { acc[i][0] = 0; acc[i][1] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma1688(double* out, int iters, double seed) {
    double acc[NACC][4];
    double a[4], b[2];
    for (int i = 0; i < 4; i++) a[i] = seed + i + threadIdx.x * 1e-9;
    for (int i = 0; i < 2; i++) b[i] = seed * 0.5 + i;
#pragma unroll
    for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) acc[i][j] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma1688(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma16816(double* out, int iters, double seed) {
    double acc[NACC][4];
    double a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = seed + i + threadIdx.x * 1e-9;
    for (int i = 0; i < 4; i++) b[i] = seed * 0.5 + i;
#pragma unroll
    for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) acc[i][j] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma16816(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double seed) {
    double acc[NACC];
    double a = 1.0 + seed * 1e-9, b = seed * 1e-9 + threadIdx.x * 1e-12;
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// per iteration: NM m8n8k4 DMMAs (256 MAC each, per warp) and NF DFMAs per thread (32 MAC per warp-instr)
template <int NM, int NF>
__global__ void k_mixed(double* out, int iters, double seed) {
    double acc[NM > 0 ? NM : 1][2];
    double f[NF > 0 ? NF : 1];
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
    double fa = 1.0 + seed * 1e-9, fb = seed * 1e-9 + threadIdx.x * 1e-12;
#pragma unroll
    for (int i = 0; i < NM; i++) { acc[i][0] = 0; acc[i][1] = 0; }
#pragma unroll
    for (int i = 0; i < NF; i++) f[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < (NM > NF ? NM : NF); i++) {
            if (i < NM) dmma884(acc[i][0], acc[i][1], a, b);
            if (i < NF) f[i] = fma(f[i], fa, fb);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; i++) s += acc[i][0] + acc[i][1];
#pragma unroll
    for (int i = 0; i < NF; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double* out, int iters, double seed) {
    double x[4];
    for (int i = 0; i < 4; i++) x[i] = -(seed + i * 0.37 + threadIdx.x * 1e-3);
    double s = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) { s += exp(x[i]); x[i] -= 1e-6; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_kernel(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 20000;
    for (int wps : {4, 8, 16}) {     // warps per CTA; 2 CTAs per SM
        int threads = wps * 32, blocks = sms * 2;
        double warps = (double)blocks * wps;
        {
            float ms = time_kernel([&] { k_dmma884<8><<<blocks, threads>>>(out, iters, 1.0); });
            double flops = warps * iters * 8.0 * 256 * 2;
            printf("dmma m8n8k4   warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", wps * 2, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_kernel([&] { k_dmma1688<8><<<blocks, threads>>>(out, iters, 1.0); });
            double flops = warps * iters * 8.0 * 1024 * 2;
            printf("dmma m16n8k8  warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", wps * 2, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_kernel([&] { k_dmma16816<8><<<blocks, threads>>>(out, iters, 1.0); });
            double flops = warps * iters * 8.0 * 2048 * 2;
            printf("dmma m16n8k16 warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", wps * 2, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_kernel([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0); });
            double flops = warps * iters * 8.0 * 32 * 2;
            printf("dfma          warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", wps * 2, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_kernel([&] { k_mixed<8, 0><<<blocks, threads>>>(out, iters, 1.0); });
            printf("mixed 8 dmma + 0 dfma  warps/SM=%2d  %.3f ms\n", wps * 2, ms);
            ms = time_kernel([&] { k_mixed<0, 8><<<blocks, threads>>>(out, iters, 1.0); });
            printf("mixed 0 dmma + 8 dfma  warps/SM=%2d  %.3f ms\n", wps * 2, ms);
            ms = time_kernel([&] { k_mixed<8, 8><<<blocks, threads>>>(out, iters, 1.0); });
            printf("mixed 8 dmma + 8 dfma  warps/SM=%2d  %.3f ms  (sum => shared pipe, max => overlapped)\n", wps * 2, ms);
            ms = time_kernel([&] { k_mixed<8, 2><<<blocks, threads>>>(out, iters, 1.0); });
            printf("mixed 8 dmma + 2 dfma  warps/SM=%2d  %.3f ms\n", wps * 2, ms);
            ms = time_kernel([&] { k_mixed<8, 4><<<blocks, threads>>>(out, iters, 1.0); });
            printf("mixed 8 dmma + 4 dfma  warps/SM=%2d  %.3f ms\n", wps * 2, ms);
        }
        {
            float ms = time_kernel([&] { k_exp<<<blocks, threads>>>(out, iters / 10, 1.0); });
            double n = warps * 32 * (iters / 10) * 4.0;
            printf("exp fp64      warps/SM=%2d  %.3f ms  %.2f Gexp/s\n", wps * 2, ms, n / ms * 1e-6);
        }
    }
    CK(cudaFree(out));
    return 0;
}
