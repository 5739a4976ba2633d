// DFMA / F2F.F64.F32 latency and issue rate on one SM sub-partition: W warps per SMSP, ILP independent chains per thread.
// Prints cycles per instruction per warp (clock64 around an unrolled loop); 148 x 1 block keeps every SM equally busy.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int MODE> __global__ void k(double *out, long long *cyc, int n)
{
    double a[ILP];
    float f[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x * 1e-9 + i; f[i] = threadIdx.x * 1e-3f + i; }
    const double c = 1.0000001;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0) a[i] = fma(a[i], c, 1.0);
                else { a[i] = (double)f[i] + a[i] * 0.0; f[i] = (float)a[i] + 1.0f; }   // F2F.F64.F32 -> DFMA -> F2F.F32.F64 -> FADD chain
            }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int MODE> void run(int warps_per_smsp, double *out, long long *cyc)
{
    const int n = 512;
    k<ILP, MODE><<<148, warps_per_smsp * 4 * 32>>>(out, cyc, n);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%s warps/SMSP=%d ILP=%d : %.2f clk per instruction per warp, %.2f clk per instruction per SMSP\n", MODE ? "cvt-chain" : "DFMA", warps_per_smsp, ILP,
           (double)c / (n * 8.0 * ILP), (double)c / (n * 8.0 * ILP * warps_per_smsp));
}
int main()
{
    double *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    for (int w : {1, 2, 4}) { run<1, 0>(w, out, cyc); run<2, 0>(w, out, cyc); run<4, 0>(w, out, cyc); run<8, 0>(w, out, cyc); run<16, 0>(w, out, cyc); }
    for (int w : {1, 2, 4}) { run<1, 1>(w, out, cyc); run<4, 1>(w, out, cyc); run<8, 1>(w, out, cyc); }
    return 0;
}
