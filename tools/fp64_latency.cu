// Dependent-chain latencies of the FP64 operations that bound the Thomas recurrence of the implicit vertical solve
// (one warp, clock64 around a chain of N dependent operations): DFMA, reciprocal (MUFU.RCP64H + Newton), division,
// and the per-level chain of the solver itself (fma -> rcp -> mul).  Measurement tool, not part of the product.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void chain(double* out, long long* cycles, double seed, int n) {
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001, t = 0.3;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        if (MODE == 0) x = fma(x, y, 1e-9);                       // dependent DFMA
        if (MODE == 1) x = 1.0 / x + 0.5;                          // reciprocal + add
        if (MODE == 2) x = y / x + 0.5;                            // division + add
        if (MODE == 3) { const double b = fma(-0.25, t, 1.5 + x * 1e-30); const double r = 1.0 / b; t = 0.4 * r; x = r; }   // solver chain
        if (MODE == 4) x = __drcp_rn(x) + 0.5;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + t;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * sizeof(double)); cudaMalloc(&cyc, sizeof(long long));
    const int n = 4096;
    const char* names[5] = {"dfma", "rcp+add", "div+add", "thomas chain (fma, rcp, mul)", "__drcp_rn+add"};
    printf("{");
    for (int m = 0; m < 5; m++) {
        long long c = 0;
        for (int rep = 0; rep < 2; rep++) {
            switch (m) {
                case 0: chain<0><<<1, 32>>>(out, cyc, 1.0, n); break;
                case 1: chain<1><<<1, 32>>>(out, cyc, 1.3, n); break;
                case 2: chain<2><<<1, 32>>>(out, cyc, 1.3, n); break;
                case 3: chain<3><<<1, 32>>>(out, cyc, 1.3, n); break;
                case 4: chain<4><<<1, 32>>>(out, cyc, 1.3, n); break;
            }
            cudaDeviceSynchronize();
            cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        }
        printf("%s\"%s\": %.1f", m ? ", " : "", names[m], (double)c / n);
    }
    printf("}\n");
    return 0;
}
