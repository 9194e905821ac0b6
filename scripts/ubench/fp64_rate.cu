// Micro-benchmark: FP64 issue rate per SM sub-partition on sm_100a (DFMA with 3 distinct sources, DFMA with a shared
// source, DMUL, DADD) at 1..8 warps per scheduler with 8 independent chains per thread. Prints cycles per warp-instruction per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters, double x, double y) {
    double a[8], b[8], c[8];
    for (int i = 0; i < 8; ++i) { a[i] = x + i + threadIdx.x; b[i] = y + 0.5 * i; c[i] = x * y + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fma(a[i], b[i], c[i]);          // 3 distinct sources
            else if (MODE == 1) a[i] = fma(a[i], a[i], a[i]);     // one source register pair
            else if (MODE == 2) a[i] = a[i] * b[i];
            else if (MODE == 3) a[i] = a[i] + b[i];
            else { a[i] = fma(a[i], b[i], c[i]); b[i] = b[i] * c[i]; }   // mix
        }
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int MODE> void run(const char* name, double* d) {
    for (int warps = 4; warps <= 32; warps *= 2) {       // warps per SM (4 schedulers)
        int iters = 20000;
        k<MODE><<<148, warps * 32>>>(d, iters, 1.0000001, 0.9999999);
        cudaDeviceSynchronize();
        double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        double n_inst = (double)iters * 8 * (MODE == 4 ? 2 : 1) * (warps / 4.0);   // warp-instructions per SMSP
        printf("%-28s warps/SMSP=%d  cycles per warp-instr per SMSP = %.3f\n", name, warps / 4, cyc / n_inst);
    }
}
int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * 8);
    run<0>("DFMA 3 distinct sources", d);
    run<1>("DFMA same source", d);
    run<2>("DMUL", d);
    run<3>("DADD", d);
    run<4>("DFMA+DMUL mix", d);
    return 0;
}
