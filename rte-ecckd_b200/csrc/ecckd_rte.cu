// ecckd_rte.cu -- row N1 of SURVEY.md 8f: minimal broadband flux solvers ON THE DEVICE, consuming what the gas optics
// wrote, so that a host model (or bench.py's second e2e figure) moves (ncol, nlay+1) fluxes over PCIe instead of the
// (ncol, nlay, ngpt) optical properties and sources (230 GB per 1M columns).
//
// They stand where the reference drivers call rte-rrtmgp (which is neither vendored nor pinned by the reference):
//   rte_lw  example/rfmip-rad-irf/ecckd_rfmip_lw.F90:130-135   no-scattering, one Gauss angle, linear-in-tau source
//   rte_sw  example/rfmip-rad-irf/ecckd_rfmip_sw.F90:148-154   two-stream + adding with a direct beam
// and follow the same published formulation as the numpy solvers in tests/rte_solvers.py, against which they are
// checked (the 0.01 W m-2 criterion of BASELINE.json).  Not part of the ty_gas_optics boundary.
//
// Layout: thread = (column, g-point).  A block is 32 columns x up to 32 g-points (one warp per g-point, lanes =
// columns: every load is a 256-byte row of the gas-optics output).  The recurrences run along the layers inside a
// thread; at every level the g-points are summed through a double-buffered shared-memory tile by the warp whose turn
// it is (one __syncthreads per level).  HBM-bound: LW reads tau and lay_source twice (down and up sweep), SW needs
// three sweeps and three scratch arrays (direct beam, albedo, source of the adding method).
#include "ecckd_kernels.h"
#include <cuda_runtime.h>

namespace ecckd {

namespace {

constexpr int kRC = 32;        // columns per block
constexpr int kRG = 32;        // g-points (warps) per block and pass

struct LevelSum {
    double (*buf)[kRG][kRC];   // [2][g][col]
    int nw, warp, lane;
    // adds the block's g-points at one level; warp (turn % nw) writes the sum (or accumulates it for later g-passes)
    __device__ __forceinline__ void operator()(int turn, double v, double *dst, long long idx, bool col_ok, bool accumulate, double scale) const
    {
        buf[turn & 1][warp][lane] = v;
        __syncthreads();
        if (warp == turn % nw && col_ok) {
            double s = 0.0;
            for (int g = 0; g < nw; ++g) s += buf[turn & 1][g][lane];
            s *= scale;
            dst[idx] = accumulate ? dst[idx] + s : s;
        }
    }
};

// ---- longwave ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRC *kRG)
rte_lw_kernel(int ncol, int nlay, int ngpt, const double *__restrict__ tau, const double *__restrict__ lay, const double *__restrict__ inc,
              const double *__restrict__ dec, const double *__restrict__ sfc, double emis, double D, double *flux_up, double *flux_dn)
{
    __shared__ double buf[2][kRG][kRC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long col = (long long)blockIdx.x * kRC + lane;
    const bool col_ok = col < ncol;
    const long long c = col_ok ? col : ncol - 1;
    const long long plane = (long long)ncol * nlay;
    const double pi = 3.14159265358979323846;
    const LevelSum sum{buf, nw, warp, lane};
    int turn = 0;
    for (int g0 = 0; g0 < ngpt; g0 += nw) {          // g-passes (one for ngpt <= 32)
        const int g = g0 + warp;
        const bool g_ok = g < ngpt;
        const long long base = c + plane * (g_ok ? g : ngpt - 1);
        const bool acc = g0 > 0;
        // down sweep
        double rad = 0.0;
        sum(turn++, 0.0, flux_dn, c, col_ok, acc, pi);
        // (the inputs of the next layer are fetched before the level barrier of this one: the barrier would otherwise
        //  expose a DRAM round trip per level)
        double tn = tau[base], ln = lay[base], sn = inc[base];
        for (int j = 0; j < nlay; ++j) {
            const double tj = tn, lj = ln, ij = sn;
            if (j + 1 < nlay) { const long long o = base + (long long)ncol * (j + 1); tn = tau[o]; ln = lay[o]; sn = inc[o]; }
            const double x = tj * D, t = exp(-x);
            const double fact = x > 1e-8 ? (1.0 - t) / x - t : 0.5 * x;
            rad = t * rad + ((1.0 - t) * ij + 2.0 * fact * (lj - ij));
            sum(turn++, g_ok ? rad : 0.0, flux_dn, c + (long long)ncol * (j + 1), col_ok, acc, pi);
        }
        // up sweep
        rad = emis * sfc[c + (long long)ncol * (g_ok ? g : ngpt - 1)] + (1.0 - emis) * rad;
        sum(turn++, g_ok ? rad : 0.0, flux_up, c + (long long)ncol * nlay, col_ok, acc, pi);
        { const long long o = base + (long long)ncol * (nlay - 1); tn = tau[o]; ln = lay[o]; sn = dec[o]; }
        for (int j = nlay - 1; j >= 0; --j) {
            const double tj = tn, lj = ln, dj = sn;
            if (j > 0) { const long long o = base + (long long)ncol * (j - 1); tn = tau[o]; ln = lay[o]; sn = dec[o]; }
            const double x = tj * D, t = exp(-x);
            const double fact = x > 1e-8 ? (1.0 - t) / x - t : 0.5 * x;
            rad = t * rad + ((1.0 - t) * dj + 2.0 * fact * (lj - dj));
            sum(turn++, g_ok ? rad : 0.0, flux_up, c + (long long)ncol * j, col_ok, acc, pi);
        }
    }
}

// ---- shortwave ---------------------------------------------------------------------------------------------------
struct TwoStream { double Rdif, Tdif, Rdir, Tdir, Tnoscat; };

__device__ __forceinline__ TwoStream two_stream(double tau, double ssa, double g, double mu0)
{
    TwoStream r;
    const double gamma1 = (8.0 - ssa * (5.0 + 3.0 * g)) * 0.25;
    const double gamma2 = 3.0 * (ssa * (1.0 - g)) * 0.25;
    const double gamma3 = (2.0 - 3.0 * mu0 * g) * 0.25;
    const double gamma4 = 1.0 - gamma3;
    const double k = sqrt(fmax((gamma1 - gamma2) * (gamma1 + gamma2), 1e-12));
    const double e = exp(-tau * k), e2 = e * e;
    const double rt = 1.0 / (k + gamma1 + (k - gamma1) * e2);
    r.Rdif = rt * gamma2 * (1.0 - e2);
    r.Tdif = rt * 2.0 * k * e;
    r.Tnoscat = exp(-tau / mu0);
    const double k_mu = k * mu0;
    const double d0 = 1.0 - k_mu * k_mu;
    const double denom = fabs(d0) > 1e-8 ? d0 : 1e-8;
    const double rt2 = ssa * rt / denom;
    const double a13 = gamma1 * gamma3 + gamma2 * gamma4, a14 = gamma1 * gamma4 + gamma2 * gamma3;
    r.Rdir = rt2 * ((1.0 - k_mu) * (a13 + k * gamma3) - (1.0 + k_mu) * (a13 - k * gamma3) * e2 -
                    2.0 * (k * gamma3 - a13 * k_mu) * e * r.Tnoscat);
    r.Tdir = -rt2 * ((1.0 + k_mu) * (a14 + k * gamma4) * r.Tnoscat - (1.0 - k_mu) * (a14 - k * gamma4) * e2 * r.Tnoscat -
                     2.0 * (k * gamma4 + a14 * k_mu) * e);
    return r;
}

// scratch arrays are (ncol, nlay+1, ngpt) like the level quantities they hold
__global__ void __launch_bounds__(kRC *kRG)
rte_sw_kernel(int ncol, int nlay, int ngpt, const double *__restrict__ tau, const double *__restrict__ ssa, const double *__restrict__ asy,
              const double *__restrict__ toa, double mu0, double alb_sfc, double *s_dir, double *s_alb, double *s_src,
              double *flux_up, double *flux_dn, double *flux_dir)
{
    __shared__ double buf[2][kRG][kRC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long col = (long long)blockIdx.x * kRC + lane;
    const bool col_ok = col < ncol;
    const long long c = col_ok ? col : ncol - 1;
    const long long plane = (long long)ncol * nlay, lplane = (long long)ncol * (nlay + 1);
    const LevelSum sum{buf, nw, warp, lane};
    int turn = 0;
    for (int g0 = 0; g0 < ngpt; g0 += nw) {
        const int g = g0 + warp;
        const bool g_ok = g < ngpt;
        const int gg = g_ok ? g : ngpt - 1;
        const long long base = c + plane * gg, lbase = c + lplane * gg;
        const bool acc = g0 > 0;
        // A: direct beam, top down
        double dir = toa[c + (long long)ncol * gg] * mu0;
        s_dir[lbase] = dir;
        sum(turn++, g_ok ? dir : 0.0, flux_dir, c, col_ok, acc, 1.0);
        double tn = tau[base];
        for (int j = 0; j < nlay; ++j) {
            const double tj = tn;
            if (j + 1 < nlay) tn = tau[base + (long long)ncol * (j + 1)];
            dir *= exp(-tj / mu0);
            s_dir[lbase + (long long)ncol * (j + 1)] = dir;
            sum(turn++, g_ok ? dir : 0.0, flux_dir, c + (long long)ncol * (j + 1), col_ok, acc, 1.0);
        }
        // B: adding, surface upwards
        double albedo = alb_sfc, src = alb_sfc * dir;
        s_alb[lbase + (long long)ncol * nlay] = albedo;
        s_src[lbase + (long long)ncol * nlay] = src;
        double sn, an, dn_;
        { const long long o = base + (long long)ncol * (nlay - 1); tn = tau[o]; sn = ssa[o]; an = asy[o]; dn_ = s_dir[lbase + (long long)ncol * (nlay - 1)]; }
        for (int j = nlay - 1; j >= 0; --j) {
            const double tj = tn, sj = sn, aj = an, dj = dn_;
            if (j > 0) { const long long o = base + (long long)ncol * (j - 1); tn = tau[o]; sn = ssa[o]; an = asy[o]; dn_ = s_dir[lbase + (long long)ncol * (j - 1)]; }
            const TwoStream ts = two_stream(tj, sj, aj, mu0);
            const double d = 1.0 / (1.0 - ts.Rdif * albedo);
            const double src_up = ts.Rdir * dj, src_dn = ts.Tdir * dj;
            src = src_up + ts.Tdif * d * (src + albedo * src_dn);
            albedo = ts.Rdif + ts.Tdif * ts.Tdif * albedo * d;
            s_alb[lbase + (long long)ncol * j] = albedo;
            s_src[lbase + (long long)ncol * j] = src;
        }
        // C: fluxes, top down
        double dn = 0.0;
        sum(turn++, g_ok ? src : 0.0, flux_up, c, col_ok, acc, 1.0);                                   // up(0) = src(0)
        sum(turn++, g_ok ? s_dir[lbase] : 0.0, flux_dn, c, col_ok, acc, 1.0);                          // dn(0) + dir(0)
        double a1n, s1n, d1n, djn = s_dir[lbase];
        { tn = tau[base]; sn = ssa[base]; an = asy[base]; const long long l1 = lbase + (long long)ncol; a1n = s_alb[l1]; s1n = s_src[l1]; d1n = s_dir[l1]; }
        for (int j = 0; j < nlay; ++j) {
            const double tj = tn, sj = sn, aj = an, a1 = a1n, s1 = s1n, d1 = d1n, dj = djn;
            if (j + 1 < nlay) {
                const long long o = base + (long long)ncol * (j + 1), l1 = lbase + (long long)ncol * (j + 2);
                tn = tau[o]; sn = ssa[o]; an = asy[o]; a1n = s_alb[l1]; s1n = s_src[l1]; d1n = s_dir[l1];
            }
            djn = d1;
            const TwoStream ts = two_stream(tj, sj, aj, mu0);
            const double d = 1.0 / (1.0 - ts.Rdif * a1);
            dn = (ts.Tdif * dn + ts.Rdif * s1 + ts.Tdir * dj) * d;
            const double up = dn * a1 + s1;
            sum(turn++, g_ok ? up : 0.0, flux_up, c + (long long)ncol * (j + 1), col_ok, acc, 1.0);
            sum(turn++, g_ok ? dn + d1 : 0.0, flux_dn, c + (long long)ncol * (j + 1), col_ok, acc, 1.0);
        }
    }
}

}  // namespace

cudaError_t launch_rte_lw(int ncol, int nlay, int ngpt, const double *tau, const double *lay, const double *inc, const double *dec,
                          const double *sfc, double emis, double D, double *flux_up, double *flux_dn, cudaStream_t s)
{
    const int nw = ngpt < kRG ? ngpt : kRG;
    rte_lw_kernel<<<(unsigned)((ncol + kRC - 1) / kRC), kRC * nw, 0, s>>>(ncol, nlay, ngpt, tau, lay, inc, dec, sfc, emis, D, flux_up, flux_dn);
    return cudaGetLastError();
}

size_t rte_sw_scratch_bytes(int ncol, int nlay, int ngpt) { return (size_t)3 * ncol * (nlay + 1) * ngpt * sizeof(double); }

cudaError_t launch_rte_sw(int ncol, int nlay, int ngpt, const double *tau, const double *ssa, const double *g, const double *toa,
                          double mu0, double alb_sfc, void *scratch, double *flux_up, double *flux_dn, double *flux_dir, cudaStream_t s)
{
    const int nw = ngpt < kRG ? ngpt : kRG;
    const size_t n = (size_t)ncol * (nlay + 1) * ngpt;
    double *sc = static_cast<double *>(scratch);
    rte_sw_kernel<<<(unsigned)((ncol + kRC - 1) / kRC), kRC * nw, 0, s>>>(ncol, nlay, ngpt, tau, ssa, g, toa, mu0, alb_sfc, sc, sc + n, sc + 2 * n,
                                                                         flux_up, flux_dn, flux_dir);
    return cudaGetLastError();
}

}  // namespace ecckd
