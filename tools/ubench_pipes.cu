// Throughput microbenchmark for the integer / fp64 instructions the encode kernel leans on.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int OP> __global__ void k(int64_t *out, int iters, int a0, int b0) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    int a = a0 + tid, b = b0 ^ tid;
    int64_t acc0 = tid, acc1 = tid + 1, acc2 = tid + 2, acc3 = tid + 3, acc4 = 5, acc5 = 6, acc6 = 7, acc7 = 8;
    int i0 = tid, i1 = tid + 1, i2 = tid + 2, i3 = tid + 3, i4 = 4, i5 = 5, i6 = 6, i7 = 7;
    double d0 = tid, d1 = 1, d2 = 2, d3 = 3, d4 = 4, d5 = 5, d6 = 6, d7 = 7, da = a, db = b * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (OP == 0) {  // IMAD.WIDE: 64-bit accumulate of 32x32 product
#define MW(acc, x, y) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x), "r"(y))
                MW(acc0, a, (int)acc0); MW(acc1, b, (int)acc1); MW(acc2, a, (int)acc2); MW(acc3, b, (int)acc3); MW(acc4, a, (int)acc4); MW(acc5, b, (int)acc5); MW(acc6, a, (int)acc6); MW(acc7, b, (int)acc7);
            } else if (OP == 1) {  // IMAD 32-bit
                i0 = i0 * a + b; i1 = i1 * b + a; i2 = i2 * a + b; i3 = i3 * b + a; i4 = i4 * a + b; i5 = i5 * b + a; i6 = i6 * a + b; i7 = i7 * b + a;
            } else if (OP == 2) {  // DFMA
                d0 = fma(d0, da, db); d1 = fma(d1, da, db); d2 = fma(d2, da, db); d3 = fma(d3, da, db);
                d4 = fma(d4, da, db); d5 = fma(d5, da, db); d6 = fma(d6, da, db); d7 = fma(d7, da, db);
            } else if (OP == 3) {  // IADD3 / LOP3 mix (alu pipe)
                i0 = (i0 + a) ^ b; i1 = (i1 + b) ^ a; i2 = (i2 + a) ^ b; i3 = (i3 + b) ^ a; i4 = (i4 + a) ^ b; i5 = (i5 + b) ^ a; i6 = (i6 + a) ^ b; i7 = (i7 + b) ^ a;
            } else if (OP == 4) {  // abs + max + add (stat ops)
                i0 = max(i0, abs(i1 + a)); i1 += abs(i2 ^ b); i2 = max(i2, abs(i3 + a)); i3 += abs(i4 ^ b); i4 = max(i4, abs(i5 + a)); i5 += abs(i6 ^ b); i6 = max(i6, abs(i7 + a)); i7 += abs(i0 ^ b);
            }
        }
    }
    out[tid] = acc0 + acc1 + acc2 + acc3 + acc4 + acc5 + acc6 + acc7 + i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7 + (int64_t)(d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7);
}

template <int OP> void run(const char *name, double ops_per_iter) {
    int64_t *out; cudaMalloc(&out, 148 * 4 * 512 * 8);
    const int iters = 4096;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<148 * 4, 512>>>(out, 16, 3, 5);
    cudaEventRecord(a);
    k<OP><<<148 * 4, 512>>>(out, iters, 3, 5);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double total = (double)148 * 4 * 512 * iters * 16 * ops_per_iter;
    printf("%-28s %8.3f ms  %8.2f Tops/s  (%.1f lanes/clk/SM at 1.9 GHz)\n", name, ms, total / ms * 1e-9, total / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}
int main() {
    run<0>("IMAD.WIDE (64-bit acc)", 8);
    run<1>("IMAD (32-bit)", 8);
    run<2>("DFMA", 8);
    run<3>("IADD3+LOP3 pairs", 16);
    run<4>("abs/max/add mix", 8 * 2.5);
    return 0;
}
