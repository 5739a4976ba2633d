// Write-only / read-only / copy bandwidth probe: what a write-dominated kernel can reach on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_write(double *p, size_t n, double v) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(p + i, v); }
__global__ void k_write_plain(double *p, size_t n, double v) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v; }
__global__ void k_write2(double2 *p, size_t n, double v) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(p + i, make_double2(v, v)); }
__global__ void k_read(const double *p, size_t n, double *out) { double s = 0; for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += __ldcs(p + i); if (s == 1.2345) *out = s; }
__global__ void k_copy(const double *a, double *b, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(b + i, __ldcs(a + i)); }
// strided-plane writes like the gas-optics outputs: rows of 16 doubles (128 B) in scattered order
__global__ void k_write_rows(double *p, size_t nrows, size_t pitch_elems) {
    const int lane16 = threadIdx.x & 15; size_t row = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 4;
    for (; row < nrows; row += ((size_t)gridDim.x * blockDim.x) >> 4) {
        size_t r = (row * 2654435761ull) % nrows;
        __stcs(p + r * pitch_elems + lane16, 1.0);
    }
}
template <typename F> float timeit(F f) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaEventRecord(a); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5; }
int main() {
    size_t n = (size_t)1 << 30;   // 8 GiB of doubles
    double *a, *b, *o; cudaMalloc(&a, n * 8); cudaMalloc(&b, n * 8); cudaMalloc(&o, 8);
    cudaMemset(a, 0, n * 8); cudaMemset(b, 0, n * 8);
    int g = 148 * 16;
    printf("write st.cs   : %.0f GB/s\n", n * 8 / (timeit([&] { k_write<<<g, 512>>>(a, n, 1.0); }) * 1e6));
    printf("write plain   : %.0f GB/s\n", n * 8 / (timeit([&] { k_write_plain<<<g, 512>>>(a, n, 1.0); }) * 1e6));
    printf("write st.cs v2: %.0f GB/s\n", n * 8 / (timeit([&] { k_write2<<<g, 512>>>((double2 *)a, n / 2, 1.0); }) * 1e6));
    printf("cudaMemset    : %.0f GB/s\n", n * 8 / (timeit([&] { cudaMemsetAsync(a, 0, n * 8); }) * 1e6));
    printf("read ld.cs    : %.0f GB/s\n", n * 8 / (timeit([&] { k_read<<<g, 512>>>(a, n, o); }) * 1e6));
    printf("copy (r+w)    : %.0f GB/s\n", 2 * n * 8 / (timeit([&] { k_copy<<<g, 512>>>(a, b, n); }) * 1e6));
    printf("cudaMemcpy D2D: %.0f GB/s (r+w)\n", 2 * n * 8 / (timeit([&] { cudaMemcpyAsync(b, a, n * 8, cudaMemcpyDeviceToDevice); }) * 1e6));
    size_t nrows = n / 16;
    printf("write 128B rows scattered: %.0f GB/s\n", n * 8 / (timeit([&] { k_write_rows<<<g, 512>>>(a, nrows, 16); }) * 1e6));
    return 0;
}
