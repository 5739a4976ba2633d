// Store-pattern probe: write 4 arrays of (ncol, nlay, ngpt) doubles exactly the way the gas-optics
// kernel's writer warps do (block = TC columns, rows of TC*8 bytes at pitch ncol*8), for several TC.
#include <cstdio>
#include <cuda_runtime.h>
template <int TC, bool CS>
__global__ void k(double *o0, double *o1, double *o2, double *o3, int ncol, int nlay, int ngpt)
{
    const int c = threadIdx.x % TC, gs = threadIdx.x / TC, GSTEP = blockDim.x / TC;
    const long long col = (long long)blockIdx.x * TC + c;
    if (col >= ncol) return;
    for (int lay = 0; lay < nlay; ++lay)
        for (int g = gs; g < ngpt; g += GSTEP) {
            const long long off = col + (long long)ncol * (lay + (long long)nlay * g);
            if (CS) { __stcs(o0 + off, 1.0); __stcs(o1 + off, 2.0); __stcs(o2 + off, 3.0); __stcs(o3 + off, 4.0); }
            else { o0[off] = 1.0; o1[off] = 2.0; o2[off] = 3.0; o3[off] = 4.0; }
        }
}
// g-outer variant: each block walks g slowly -> consecutive layers of a (block, g) are 512 KB apart as well
template <typename F> float timeit(F f) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaEventRecord(a); for (int i = 0; i < 3; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / 3; }
int main()
{
    const int ncol = 65536, nlay = 137, ngpt = 32;
    const size_t n = (size_t)ncol * nlay * ngpt;
    double *o[4]; for (auto &p : o) { cudaMalloc(&p, n * 8); cudaMemset(p, 0, n * 8); }
    const double gb = 4.0 * n * 8 / 1e9;
#define RUN(TC, CS, THREADS) printf("TC=%3d threads=%4d %s : %.0f GB/s\n", TC, THREADS, CS ? "st.cs" : "plain", gb / (timeit([&] { k<TC, CS><<<(ncol + TC - 1) / TC, THREADS>>>(o[0], o[1], o[2], o[3], ncol, nlay, ngpt); }) * 1e-3));
    RUN(16, true, 512) RUN(16, false, 512) RUN(16, true, 64) RUN(32, true, 512) RUN(32, false, 512) RUN(64, true, 512) RUN(128, true, 512) RUN(256, true, 512) RUN(512, true, 512)
    return 0;
}
