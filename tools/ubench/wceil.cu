// Write ceiling of the gas-optics output pattern: persistent blocks (one per SM) walk tiles of TC columns and store
// every (array, layer, g) row of the tile -- TC*8 bytes at pitch ncol*8 / plane ncol*nlay*8 -- with nothing else to do.
// What a kernel with this store pattern can reach at best on this GPU (compare: MEASURED_PEAKS.json copy bandwidth).
#include <cstdio>
#include <cuda_runtime.h>
template <int TC, bool GFAST>
__global__ void k(double *o0, double *o1, double *o2, double *o3, int ncol, int nlay, int ngpt, int narr)
{
    double *o[4] = {o0, o1, o2, o3};
    const int lanes = TC < 32 ? TC : 32;                       // lanes per row
    const int rpw = 32 / lanes;                                // rows per warp-instruction
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int c = lane % lanes, rsub = lane / lanes;
    const int ntile = (ncol + TC - 1) / TC;
    for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
        const long long col0 = (long long)t * TC;
        const int nrow = nlay * ngpt;
        for (int r = warp * rpw + rsub; r < nrow; r += nw * rpw) {
            const int lay = GFAST ? r / ngpt : r % nlay, g = GFAST ? r % ngpt : r / nlay;
            const long long off = col0 + (long long)ncol * (lay + (long long)nlay * g);
            for (int cc = c; cc < TC && col0 + cc < ncol; cc += lanes)
                for (int a = 0; a < narr; ++a) __stcs(o[a] + off + cc, 1.0 + a);
        }
    }
}
template <typename F> float timeit(F f) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaEventRecord(a); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5; }
int main()
{
    const int ncol = 65536, nlay = 137, ngpt = 32;
    const size_t n = (size_t)ncol * nlay * ngpt;
    double *o[4]; for (auto &p : o) { cudaMalloc(&p, n * 8); cudaMemset(p, 0, n * 8); }
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
#define RUN(TC, GF, THREADS, NARR, BPS) printf("TC=%3d %s threads=%4d arrays=%d blocks/SM=%d : %.0f GB/s\n", TC, GF ? "g-fast  " : "lay-fast", THREADS, NARR, BPS, \
        NARR * n * 8.0 / 1e9 / (timeit([&] { k<TC, GF><<<sms * BPS, THREADS>>>(o[0], o[1], o[2], o[3], ncol, nlay, ngpt, NARR); }) * 1e-3));
    RUN(16, true, 1024, 4, 1) RUN(32, true, 1024, 4, 1) RUN(32, true, 512, 4, 1) RUN(32, true, 256, 4, 1) RUN(32, true, 128, 4, 1)
    RUN(32, false, 1024, 4, 1) RUN(64, true, 1024, 4, 1) RUN(128, true, 1024, 4, 1) RUN(32, true, 1024, 4, 2) RUN(32, true, 1024, 3, 1) RUN(32, true, 1024, 1, 1)
    RUN(16, true, 1024, 3, 1) RUN(8, true, 1024, 4, 1)
    return 0;
}
