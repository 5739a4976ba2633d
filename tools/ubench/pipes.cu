// Micro-benchmark: per-SM throughput of DFMA, F2F.F64.F32 (+DFMA), integer f32->f64 widening (+DFMA), FFMA.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE> __global__ void k(double *out, const float *in, int n)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    float f0 = in[threadIdx.x & 31], f1 = f0 * 1.5f, f2 = f0 * 2.5f, f3 = f0 * 3.5f;
    const double c = 1.0000001;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) { a0 = fma(a0, c, a1); a1 = fma(a1, c, a2); a2 = fma(a2, c, a3); a3 = fma(a3, c, a0); }
            if (MODE == 1) {
                a0 = fma((double)f0, c, a0); a1 = fma((double)f1, c, a1); a2 = fma((double)f2, c, a2); a3 = fma((double)f3, c, a3);
                f0 += 1.0f; f1 += 1.0f; f2 += 1.0f; f3 += 1.0f;
            }
            if (MODE == 2) {
                unsigned b0 = __float_as_uint(f0), b1 = __float_as_uint(f1), b2 = __float_as_uint(f2), b3 = __float_as_uint(f3);
                double d0 = __hiloint2double((b0 >> 3) + 0x38000000, b0 << 29), d1 = __hiloint2double((b1 >> 3) + 0x38000000, b1 << 29);
                double d2 = __hiloint2double((b2 >> 3) + 0x38000000, b2 << 29), d3 = __hiloint2double((b3 >> 3) + 0x38000000, b3 << 29);
                a0 = fma(d0, c, a0); a1 = fma(d1, c, a1); a2 = fma(d2, c, a2); a3 = fma(d3, c, a3);
                f0 += 1.0f; f1 += 1.0f; f2 += 1.0f; f3 += 1.0f;
            }
            if (MODE == 3) {
                f0 = fmaf(f0, 1.0000001f, f1); f1 = fmaf(f1, 1.0000001f, f2); f2 = fmaf(f2, 1.0000001f, f3); f3 = fmaf(f3, 1.0000001f, f0);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + f0 + f1 + f2 + f3;
}
template <int MODE> void run(const char *name, double *out, float *in)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 4, threads = 512;
    k<MODE><<<blocks, threads>>>(out, in, 16);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, in, ITERS);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * ITERS * 8 * 4;   // "groups" of one op chain element
    printf("%-28s %8.3f ms  %.3e groups/s  = %.1f per clk per SM @1.9GHz\n", name, ms, ops / (ms * 1e-3), ops / (ms * 1e-3) / 148 / 1.9e9);
}
int main()
{
    double *out; float *in; cudaMalloc(&out, 148 * 4 * 512 * 8); cudaMalloc(&in, 128); cudaMemset(in, 0, 128);
    run<0>("DFMA", out, in);
    run<1>("F2F.F64.F32 + DFMA (+FADD)", out, in);
    run<2>("int widen + DFMA (+FADD)", out, in);
    run<3>("FFMA", out, in);
    return 0;
}
