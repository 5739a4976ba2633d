// ecckd_kernels.h -- host-visible launch interface of ecckd_kernels.cu
#pragma once
#include <cuda_runtime.h>

namespace ecckd {

struct DevModel;
struct CallGases;
struct GasRef;

// Arguments of one gas-optics launch.  All pointers are device memory, Fortran
// layout (column fastest).  LW: out = {tau, lay_source, lev_source_inc,
// lev_source_dec}, out2d = sfc_source.  SW: out = {tau, ssa, g, -}, out2d = toa_src.
// NULL outputs are skipped (reference error paths :414-417, :461-463).
template <typename Real> struct KernelArgs {
    int ncol, nlay;
    const Real *plev, *tlay, *tlev, *tsfc;
    Real *out[4];
    Real *out2d;
    int dbg;   // experiment switches (ECCKD_DBG env, tuning builds only); 0 in production
};

// Bytes of stream-ordered scratch launch_walk needs for the per-call mixed table (0: none).
size_t walk_scratch_bytes(const DevModel &, const CallGases &, size_t real_size);
// Launches the table-mixing kernel (when the call has constant-factor gases) and the walking kernel.
// `which`: 0 = quad kernel when the call is eligible, else the one-column-per-warp walk kernel; 1 = quad kernel only
// (cudaErrorInvalidValue when not eligible); 2 = one-column-per-warp walk kernel only.
template <typename Real, bool SW>
cudaError_t launch_walk(const DevModel &, const CallGases &, const KernelArgs<Real> &, void *scratch, cudaStream_t, int which = 0);
template <typename Real, bool SW>
cudaError_t launch_generic(const DevModel &, const CallGases &, const KernelArgs<Real> &, cudaStream_t);
template <typename Real>
cudaError_t launch_indices(const DevModel &, const GasRef *lut, int ncol, int nlay, const Real *plev, const Real *tlay,
                           int *ip0, double *pw1, int *it0, double *tw1, int *iv0, double *vw1, cudaStream_t);
// N1: minimal broadband flux solvers on the device (ecckd_rte.cu); arrays as the gas optics wrote them
cudaError_t launch_rte_lw(int ncol, int nlay, int ngpt, const double *tau, const double *lay, const double *inc, const double *dec,
                          const double *sfc, double emis, double D, double *flux_up, double *flux_dn, cudaStream_t);
size_t rte_sw_scratch_bytes(int ncol, int nlay, int ngpt);
cudaError_t launch_rte_sw(int ncol, int nlay, int ngpt, const double *tau, const double *ssa, const double *g, const double *toa,
                          double mu0, double alb_sfc, void *scratch, double *flux_up, double *flux_dn, double *flux_dir, cudaStream_t);
cudaError_t launch_div_const_check(long long n, const double *x, double c, double rc, unsigned long long *mismatches, cudaStream_t);
cudaError_t launch_planck_indices(const DevModel &, long long n, const double *T, int *it0, double *w1, cudaStream_t);
// Geometry of a walk-kernel launch over `ntile` 16-column tiles on a device with `sms` SMs: tiles per chunk (1 or 2)
// and blocks per g-group; block b walks chunks b, b + blocks, ...  Pure host arithmetic.
void walk_plan(int ntile, bool shortwave, int sms, int &tiles_per_chunk, int &blocks);
// quad kernel (ecckd_quad.cu): the product path for calls whose outputs admit TMA tensor maps
constexpr int kQuadTileCols = 32;   // columns per tile
constexpr int kQuadMaxVar = 3;      // bilinear gases with a varying concentration it keeps in registers
constexpr int kQuadMaxExtra = 4;    // further varying gases it reads uncached (3 cached + 4: nvar <= 7)
template <typename Real> bool quad_eligible(const CallGases &, const KernelArgs<Real> &);
template <typename Real, bool SW>
cudaError_t launch_quad(const DevModel &, const CallGases &, const KernelArgs<Real> &, const Real *mix, int sms, cudaStream_t);
constexpr int kWalkTileCols = 16;   // columns per tile of the fp64 / fp32 walk kernels with at most one varying gas

}  // namespace ecckd
