// Measured FP64 FMA throughput of the device (the roof of the two WENO tendency kernels; BASELINE.md section 2 asks
// the builder to measure it).  8 independent DFMA chains per thread, enough resident warps to fill the FP64 pipe.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b, int iters) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 16
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_dfma<<<blocks, threads>>>(out, 0.999999, 1e-9, iters);   // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k_dfma<<<blocks, threads>>>(out, 0.999999, 1e-9, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double flop = 2.0 * 8 * (double)iters * blocks * threads;
    printf("{\"device\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.2f, \"ms\": %.3f, \"cuda_error\": \"%s\"}\n", p.name,
           p.multiProcessorCount, flop / (best * 1e-3) / 1e12, best, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
