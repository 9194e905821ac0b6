// Micro-benchmark: FP32 issue rates per SM sub-partition on sm_100a next to DFMA — FFMA with three register sources, packed FFMA2
// (fma.rn.f32x2, two results per lane and instruction), and the F2F conversions between FP64 and FP32 — at 1..8 warps per scheduler
// with 8 independent chains per thread. Prints cycles per warp-instruction per SMSP. Decides whether a Float32 build of the stage
// kernel can beat the FP64 one without packed math.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int MODE>
__global__ void k(double* out, int iters, double x, double y) {
    long long t0, t1;
    double s = 0;
    if (MODE == 0) {            // DFMA
        double a[8], b[8], c[8];
        for (int i = 0; i < 8; ++i) { a[i] = x + i + threadIdx.x; b[i] = y + 0.5 * i; c[i] = x * y + i; }
        t0 = clock64();
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b[i], c[i]);
        t1 = clock64();
        for (int i = 0; i < 8; ++i) s += a[i];
    } else if (MODE == 1) {     // FFMA, three distinct register sources
        float a[8], b[8], c[8];
        for (int i = 0; i < 8; ++i) { a[i] = (float)x + i + threadIdx.x; b[i] = (float)y + 0.5f * i; c[i] = (float)(x * y) + i; }
        t0 = clock64();
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b[i], c[i]);
        t1 = clock64();
        for (int i = 0; i < 8; ++i) s += a[i];
    } else if (MODE == 2) {     // FFMA2 (packed pair)
        unsigned long long a[8], b[8], c[8];
        for (int i = 0; i < 8; ++i) {
            float2 fa = make_float2((float)x + i + threadIdx.x, (float)x - i), fb = make_float2((float)y + 0.5f * i, (float)y), fc = make_float2((float)(x * y) + i, 1.0f);
            a[i] = *reinterpret_cast<unsigned long long*>(&fa); b[i] = *reinterpret_cast<unsigned long long*>(&fb); c[i] = *reinterpret_cast<unsigned long long*>(&fc);
        }
        t0 = clock64();
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = ffma2(a[i], b[i], c[i]);
        t1 = clock64();
        for (int i = 0; i < 8; ++i) { float2 f = *reinterpret_cast<float2*>(&a[i]); s += f.x + f.y; }
    } else {                    // F2F: double -> float -> double round trip (two conversions per iteration and chain)
        double a[8];
        for (int i = 0; i < 8; ++i) a[i] = x + i + threadIdx.x;
        t0 = clock64();
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) { float f = __double2float_rn(a[i]); asm volatile("" : "+f"(f)); a[i] = (double)f; asm volatile("" : "+d"(a[i])); }
        t1 = clock64();
        for (int i = 0; i < 8; ++i) s += a[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int MODE> void run(const char* name, double* d, int per_iter) {
    for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000;
        k<MODE><<<148, warps * 32>>>(d, iters, 1.0000001, 0.9999999);
        cudaDeviceSynchronize();
        double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        double n_inst = (double)iters * 8 * per_iter * (warps / 4.0);
        printf("%-40s warps/SMSP=%d  cycles per warp-instr per SMSP = %.3f\n", name, warps / 4, cyc / n_inst);
    }
}
int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * 8);
    run<0>("DFMA", d, 1);
    run<1>("FFMA (3 register sources)", d, 1);
    run<2>("FFMA2 fma.rn.f32x2 (2 results per lane)", d, 1);
    run<3>("F2F f64->f32 + f32->f64 (per conversion)", d, 2);
    return 0;
}
