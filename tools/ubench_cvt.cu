// Do 64-bit conversions (I2F.F64) share the FP64 pipe with DFMA on this GPU?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_cvt ubench_cvt.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(double *out, int iters, int seed) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    double a0 = tid, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7, m = 1.0000001, c = 1e-9;
    int i0 = tid + seed, i1 = tid * 3, i2 = tid * 5, i3 = tid * 7;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0 || MODE == 2) {   // 8 DFMA
                a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
                a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
            }
            if (MODE == 1 || MODE == 2) {   // 4 I2F.F64 (+ cheap integer updates so they cannot be hoisted)
                double d0, d1, d2, d3;
                asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d0) : "r"(i0));
                asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d1) : "r"(i1));
                asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d2) : "r"(i2));
                asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d3) : "r"(i3));
                s0 = __longlong_as_double(__double_as_longlong(s0) ^ __double_as_longlong(d0));
                s1 = __longlong_as_double(__double_as_longlong(s1) ^ __double_as_longlong(d1));
                s2 = __longlong_as_double(__double_as_longlong(s2) ^ __double_as_longlong(d2));
                s3 = __longlong_as_double(__double_as_longlong(s3) ^ __double_as_longlong(d3));
                i0 += 3; i1 += 5; i2 += 7; i3 += 9;
            }
            if (MODE == 3 || MODE == 4) {   // 4 magic conversions: integer add + DADD
                double d0 = __hiloint2double(0x43300000, i0 ^ 0x80000000) - 4503601774854144.0;
                double d1 = __hiloint2double(0x43300000, i1 ^ 0x80000000) - 4503601774854144.0;
                double d2 = __hiloint2double(0x43300000, i2 ^ 0x80000000) - 4503601774854144.0;
                double d3 = __hiloint2double(0x43300000, i3 ^ 0x80000000) - 4503601774854144.0;
                s0 = __longlong_as_double(__double_as_longlong(s0) ^ __double_as_longlong(d0));
                s1 = __longlong_as_double(__double_as_longlong(s1) ^ __double_as_longlong(d1));
                s2 = __longlong_as_double(__double_as_longlong(s2) ^ __double_as_longlong(d2));
                s3 = __longlong_as_double(__double_as_longlong(s3) ^ __double_as_longlong(d3));
                i0 += 3; i1 += 5; i2 += 7; i3 += 9;
                if (MODE == 4) {
                    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
                    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
                }
            }
        }
    }
    out[tid] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + s0 + s1 + s2 + s3;
}
template <int MODE> float run(const char *name) {
    double *out; cudaMalloc(&out, 148 * 4 * 512 * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148 * 4, 512>>>(out, 16, 1);
    cudaEventRecord(a);
    k<MODE><<<148 * 4, 512>>>(out, 2048, 1);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-44s %8.3f ms\n", name, ms);
    cudaFree(out);
    return ms;
}
int main() {
    float d = run<0>("8 DFMA per step");
    float c = run<1>("4 I2F.F64 per step");
    float b = run<2>("8 DFMA + 4 I2F.F64 per step");
    float m = run<3>("4 magic conversions (IADD/LOP + DADD) per step");
    float bm = run<4>("8 DFMA + 4 magic conversions per step");
    printf("DFMA+I2F = %.2f x (DFMA alone + I2F alone)   [1.0 => same pipe, max/sum => independent pipes]\n", b / (d + c));
    printf("per I2F.F64 cost in DFMA-equivalents: %.2f; per magic conversion: %.2f\n", (c / 4) / (d / 8), (m / 4) / (d / 8));
    printf("DFMA+magic = %.3f ms vs DFMA+I2F = %.3f ms\n", bm, b);
    return 0;
}
