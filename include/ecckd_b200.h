/*
 * ecckd_b200.h -- C ABI of the B200-native ECCKD gas-optics path.
 *
 * This is the drop-in boundary: the entry points below are what the reference's
 * Fortran type-bound procedures for this path bind to once the Fortran loops are
 * replaced by the ISO_C_BINDING shim in rte-ecckd_b200/fortran/ (see
 * INTEGRATION.md).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Reference interface replaced (all in /root/reference/src/gas_optics_ecckd.f90):
 *   ecckd_model_create        <- the state load_and_init leaves in ty_gas_optics_ecckd
 *                                (:23-36; example/rfmip-rad-irf/mo_load_coefficients.F90:19-146)
 *   ecckd_model_load          <- load_and_init itself (netCDF-3 classic ck-definition file)
 *   ecckd_gas_optics_lw_*     <- gas_optics_int  (:381-426)  tau + Planck lay/lev/sfc sources
 *   ecckd_gas_optics_sw_*     <- gas_optics_ext  (:431-473)  tau/ssa/g + Rayleigh + TOA source
 *   ecckd_get_*               <- get_ngas, get_gases, get_press_min/max, get_temp_min/max,
 *                                source_is_internal/external (:477-553)
 *   ecckd_interpolation_indices_* <- (debug) ip0/it0/iv0 and weights of calculate_optical_depth (:120-163)
 *
 * Array layout is RTE's, unchanged: Fortran column-major with the column index
 * fastest -- plev(ncol,nlay+1), tlay(ncol,nlay), tau(ncol,nlay,ngpt),
 * sfc_source(ncol,ngpt): element (col,lay,g) lives at col + ncol*(lay + nlay*g).
 * Layers must be ordered top-of-atmosphere first, as the reference assumes
 * (:143; surface-first input clamps every gas optical depth to zero).
 *
 * There is no CPU fallback: every compute entry point runs sm_100a CUDA kernels
 * and fails with ECCKD_ERR_CUDA when no device/driver is usable.
 */
#ifndef ECCKD_B200_H
#define ECCKD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECCKD_MAX_GASES 16 /* absorption(16), gas(16): gas_optics_ecckd.f90:24-25 */
#define ECCKD_NAME_LEN 32  /* character(len=32) gas names */
#define ECCKD_ERRLEN 128   /* character(len=128) error_msg */

/* concentration_dependence_code, gas_optics_ecckd.f90:54-57 */
#define ECCKD_CODE_NONE 0
#define ECCKD_CODE_LINEAR 1
#define ECCKD_CODE_LOOK_UP_TABLE 2
#define ECCKD_CODE_RELATIVE_LINEAR 3

/* Return codes.  ECCKD_ERR_REFERENCE marks the errors the reference itself
 * returns through error_msg; the message text is the reference's, verbatim. */
#define ECCKD_OK 0
#define ECCKD_ERR_ARG 1
#define ECCKD_ERR_CUDA 2
#define ECCKD_ERR_REFERENCE 3
#define ECCKD_ERR_IO 4

/* Kernel selection for the compute entry points (the tests exercise every one; AUTO
 * picks the fastest kernel the call qualifies for: QUAD, else COLWALK, else GENERIC). */
#define ECCKD_KERNEL_AUTO 0
#define ECCKD_KERNEL_WALK 1    /* a table-walking kernel (QUAD when the call qualifies, else COLWALK); error if neither */
#define ECCKD_KERNEL_GENERIC 2 /* thread-per-(column,layer); any table shape / gas mix */
#define ECCKD_KERNEL_QUAD 3    /* the product path: 4 columns per warp, TMA tensor stores (needs 16-byte aligned outputs
                                  and a column count that is a multiple of 16 bytes: even for fp64, of 4 for fp32) */
#define ECCKD_KERNEL_COLWALK 4 /* one column per warp; writer-warp stores for odd column counts / unaligned outputs */

typedef struct ecckd_model ecckd_model; /* opaque: device-resident packed tables */

/* One AbsorptionTable (gas_optics_ecckd.f90:13-19) plus its name in this%gas. */
typedef struct {
    char name[ECCKD_NAME_LEN];        /* blank- or NUL-padded */
    int32_t composite_only;
    int32_t concentration_dependence_code;
    int32_t n_mole_fraction;          /* size(mole_fraction); 0 when not allocated */
    int32_t reserved;
    double reference_mole_fraction;
    const double *mole_fraction;      /* (n_mole_fraction) or NULL */
    const double *coefficient;        /* Fortran coefficient(ngpt,npress,ntemp,nv), nv = max(n_mole_fraction,1) */
} ecckd_gas_table;

/* The public components of ty_gas_optics_ecckd (gas_optics_ecckd.f90:24-36) as
 * host arrays in Fortran layout, i.e. exactly what c_loc() of them yields. */
typedef struct {
    int32_t shortwave;                /* this%shortwave */
    int32_t ngpt;                     /* size(gpoint_fraction, 2) */
    int32_t npress;                   /* size(log_pressure) */
    int32_t ntemp;                    /* size(temperature, 2) */
    int32_t nplanck;                  /* size(temperature_planck); 0 for SW */
    int32_t num_gases;                /* this%num_gases */
    const double *log_pressure;       /* (npress) */
    const double *temperature;        /* (npress, ntemp) */
    const double *temperature_planck; /* (nplanck) or NULL */
    const double *planck_function;    /* (ngpt, nplanck) or NULL */
    const double *rayleigh_molar_scattering_coeff; /* (ngpt) or NULL */
    const double *solar_irradiance;   /* (ngpt) or NULL */
    const ecckd_gas_table *gas;       /* (num_gases) */
    /* what ty_gas_optics%init receives (mo_load_coefficients.F90:55-74); may be NULL/0 */
    int32_t nband;
    int32_t reserved;
    const double *band_lims_wvn;      /* (2, nband) */
    const int32_t *band2gpt;          /* (2, nband), 1-based inclusive */
} ecckd_host_tables;

/* One gas of the caller's ty_gas_concs: value(col,lay) = ptr[col*col_stride + lay*lay_stride]
 * (strides in elements of the call's real kind).  Scalar: 0,0.  Profile (nlay): 0,1.
 * Field (ncol,nlay): 1,ncol.  This is get_vmr's broadcasting (call site :351)
 * without materialising layer_vmr. */
typedef struct {
    const void *ptr;
    int64_t col_stride;
    int64_t lay_stride;
} ecckd_vmr_view;

/* ---- model lifetime ------------------------------------------------------ */

/* Pack the tables once and replicate them on `ndev` CUDA devices (devices ==
 * NULL: devices 0..ndev-1).  Coefficients are stored on the device as fp32 when
 * every value is exactly representable (always true for tables read from the
 * f4 ck-definition files) in addition to the promoted fp64 copy. */
int ecckd_model_create(const ecckd_host_tables *tables, int ndev, const int *devices,
                       ecckd_model **model, char err[ECCKD_ERRLEN]);

/* load_and_init (mo_load_coefficients.F90:19-146) from a netCDF-3 classic file,
 * then ecckd_model_create.  For hosts without the Fortran loader. */
int ecckd_model_load(const char *ckd_file, int ndev, const int *devices,
                     ecckd_model **model, char err[ECCKD_ERRLEN]);

void ecckd_model_destroy(ecckd_model *model);

/* The loader alone (no device needed): load_and_init into host tables owned by the
 * library -- inspect them, hand them to ecckd_model_create, then ecckd_tables_free. */
int ecckd_tables_load(const char *ckd_file, const ecckd_host_tables **tables, char err[ECCKD_ERRLEN]);
void ecckd_tables_free(const ecckd_host_tables *tables);

/* ---- getters (gas_optics_ecckd.f90:477-553 + inherited get_ngpt/get_nband) - */
int ecckd_get_ngas(const ecckd_model *model);
int ecckd_get_gas_name(const ecckd_model *model, int i, char name[ECCKD_NAME_LEN]); /* 0-based */
int ecckd_get_ngpt(const ecckd_model *model);
int ecckd_get_nband(const ecckd_model *model);
int ecckd_get_band2gpt(const ecckd_model *model, int32_t *band2gpt /* (2,nband) */);
int ecckd_get_band_lims_wvn(const ecckd_model *model, double *band_lims /* (2,nband) */);
/* Loader hardening: the algorithm takes every grid spacing from the first two entries (gas_optics_ecckd.f90:104-106,
 * 154-155,271) and so does this library; grids that are NOT uniform within 1e-3 are named in this string ("" = all
 * uniform).  With ECCKD_STRICT_GRIDS=1 in the environment ecckd_model_create refuses such tables instead. */
const char *ecckd_get_table_warnings(const ecckd_model *model);
int ecckd_get_ndev(const ecckd_model *model);
int ecckd_get_device(const ecckd_model *model, int dev_index);
int ecckd_source_is_internal(const ecckd_model *model);
int ecckd_source_is_external(const ecckd_model *model);
double ecckd_get_press_min(const ecckd_model *model);
double ecckd_get_press_max(const ecckd_model *model);
double ecckd_get_temp_min(const ecckd_model *model);
double ecckd_get_temp_max(const ecckd_model *model);
double ecckd_get_total_solar_irradiance(const ecckd_model *model);

/* ---- compute: device-resident arrays, asynchronous on `stream` ------------
 * dev_index selects one of the model's devices; all pointers (including every
 * vmr[].ptr) must be device memory on that device.  `stream` is a cudaStream_t
 * (NULL = the legacy default stream).  `play` and `col_dry` of the reference
 * signature are never read by it (:386,394) and are not part of this ABI.
 *
 * Gas handling follows gas_optical_depth (:323-376): the caller's names are
 * visited in order, names the model does not know are skipped silently, and the
 * composite table is evaluated once, on the first of its alias gases.
 *
 * LW: tlev == NULL reproduces the reference: tau, lay_source and sfc_source are
 * written, then ECCKD_ERR_REFERENCE "tlev is required for ecckd" (:414-417).
 * SW: ssa == NULL or g == NULL models a non-two-stream optical_props: tau (gas +
 * Rayleigh) is written, then "shortwave must use ty_optical_props_2str" (:461-463).
 */
int ecckd_gas_optics_lw_f64(ecckd_model *model, int dev_index, int ncol, int nlay,
                            const double *plev, const double *tlay, const double *tlev, const double *tsfc,
                            int ngas, const char *const *gas_names, const ecckd_vmr_view *vmr,
                            double *tau, double *lay_source, double *lev_source_inc, double *lev_source_dec,
                            double *sfc_source, int kernel, void *stream, char err[ECCKD_ERRLEN]);
int ecckd_gas_optics_lw_f32(ecckd_model *model, int dev_index, int ncol, int nlay,
                            const float *plev, const float *tlay, const float *tlev, const float *tsfc,
                            int ngas, const char *const *gas_names, const ecckd_vmr_view *vmr,
                            float *tau, float *lay_source, float *lev_source_inc, float *lev_source_dec,
                            float *sfc_source, int kernel, void *stream, char err[ECCKD_ERRLEN]);
int ecckd_gas_optics_sw_f64(ecckd_model *model, int dev_index, int ncol, int nlay,
                            const double *plev, const double *tlay,
                            int ngas, const char *const *gas_names, const ecckd_vmr_view *vmr,
                            double *tau, double *ssa, double *g, double *toa_src,
                            int kernel, void *stream, char err[ECCKD_ERRLEN]);
int ecckd_gas_optics_sw_f32(ecckd_model *model, int dev_index, int ncol, int nlay,
                            const float *plev, const float *tlay,
                            int ngas, const char *const *gas_names, const ecckd_vmr_view *vmr,
                            float *tau, float *ssa, float *g, float *toa_src,
                            int kernel, void *stream, char err[ECCKD_ERRLEN]);

/* ---- compute: host-resident arrays (what the plain Fortran shim passes) -----
 * Synchronous.  Columns are split into contiguous shards over the model's
 * devices (no collective: columns are independent) and, within a device, into
 * blocks of `block_cols` columns (0 = library default) that are pipelined
 * H2D -> kernel -> D2H over three stream slots per device.  The copies run
 * straight between the caller's arrays and device buffers.  With page-locked
 * caller arrays (ecckd_host_register below, cudaHostRegister, or a pinned
 * allocator) the copies are asynchronous and the slots overlap -- PCIe-bound,
 * ~53 GB/s D2H measured; pageable arrays go
 * through the driver's staged copies.  Device block buffers have their own
 * (even) column pitch, so any `ncol` takes the TMA-store kernel.
 * Same semantics and error behaviour as the device entry points. */
int ecckd_gas_optics_lw_host_f64(ecckd_model *model, int ncol, int nlay,
                                 const double *plev, const double *tlay, const double *tlev, const double *tsfc,
                                 int ngas, const char *const *gas_names, const ecckd_vmr_view *vmr,
                                 double *tau, double *lay_source, double *lev_source_inc, double *lev_source_dec,
                                 double *sfc_source, int block_cols, char err[ECCKD_ERRLEN]);
int ecckd_gas_optics_lw_host_f32(ecckd_model *model, int ncol, int nlay,
                                 const float *plev, const float *tlay, const float *tlev, const float *tsfc,
                                 int ngas, const char *const *gas_names, const ecckd_vmr_view *vmr,
                                 float *tau, float *lay_source, float *lev_source_inc, float *lev_source_dec,
                                 float *sfc_source, int block_cols, char err[ECCKD_ERRLEN]);
int ecckd_gas_optics_sw_host_f64(ecckd_model *model, int ncol, int nlay,
                                 const double *plev, const double *tlay,
                                 int ngas, const char *const *gas_names, const ecckd_vmr_view *vmr,
                                 double *tau, double *ssa, double *g, double *toa_src,
                                 int block_cols, char err[ECCKD_ERRLEN]);
int ecckd_gas_optics_sw_host_f32(ecckd_model *model, int ncol, int nlay,
                                 const float *plev, const float *tlay,
                                 int ngas, const char *const *gas_names, const ecckd_vmr_view *vmr,
                                 float *tau, float *ssa, float *g, float *toa_src,
                                 int block_cols, char err[ECCKD_ERRLEN]);

/* ---- next row N1: minimal flux solvers on the device ------------------------
 * Not part of the ty_gas_optics boundary: they stand where the reference drivers call rte-rrtmgp's rte_lw / rte_sw
 * (ecckd_rfmip_lw.F90:130-135, ecckd_rfmip_sw.F90:148-154) so that a device-resident pipeline -- gas optics, then
 * fluxes -- sends (ncol, nlay+1) fluxes over PCIe instead of the (ncol, nlay, ngpt) optical properties.  Same
 * published formulation as tests/rte_solvers.py (LW: no scattering, one quadrature angle `diffusivity` (1.66), source
 * linear in tau, surface emissivity; SW: two-stream + adding with a direct beam at cosine `mu0`, Lambertian surface),
 * summed over the g-points.  Device arrays in the gas-optics layout; fluxes are (ncol, nlay+1), top first, W m-2. */
int ecckd_rte_lw_f64(ecckd_model *model, int dev_index, int ncol, int nlay, int ngpt,
                     const double *tau, const double *lay_source, const double *lev_source_inc, const double *lev_source_dec,
                     const double *sfc_source, double sfc_emis, double diffusivity,
                     double *flux_up, double *flux_dn, void *stream, char err[ECCKD_ERRLEN]);
int ecckd_rte_sw_f64(ecckd_model *model, int dev_index, int ncol, int nlay, int ngpt,
                     const double *tau, const double *ssa, const double *g, const double *toa_src, double mu0, double sfc_albedo,
                     double *flux_up, double *flux_dn, double *flux_dn_dir, void *stream, char err[ECCKD_ERRLEN]);

/* Page-lock / release a caller array so that the host entry points above copy asynchronously.  Registration costs
 * about as much as one pass over the memory: do it once for arrays that are passed call after call (the Fortran shim
 * exposes it as a plain subroutine).  Already page-locked memory is accepted (ECCKD_OK). */
int ecckd_host_register(void *ptr, size_t bytes, char err[ECCKD_ERRLEN]);
int ecckd_host_unregister(void *ptr);

/* ---- debug: the interpolation indices the kernels use -----------------------
 * Device arrays, (ncol,nlay) each; indices are the reference's 1-based ip0, it0,
 * iv0 (:126,138,161) and the upper weights pressure/temperature/vmr_weight1.
 * `lut_vmr` is the vmr view of the look-up-table gas named `lut_gas` (NULL name:
 * iv0 = 1, vw1 = 0).  Computed by the same device function the compute kernels
 * call.  Planck: it0 as at :278 (0 marks the T < T_planck(1) branch, :283-284). */
int ecckd_interpolation_indices_f64(ecckd_model *model, int dev_index, int ncol, int nlay,
                                    const double *plev, const double *tlay,
                                    const char *lut_gas, const ecckd_vmr_view *lut_vmr,
                                    int32_t *ip0, double *pw1, int32_t *it0, double *tw1,
                                    int32_t *iv0, double *vw1, void *stream, char err[ECCKD_ERRLEN]);
int ecckd_planck_indices_f64(ecckd_model *model, int dev_index, int64_t n, const double *temperature,
                             int32_t *it0, double *w1, void *stream, char err[ECCKD_ERRLEN]);

/* debug: the kernels divide by per-model constants (grid spacings, pi) with a reciprocal multiplication plus an exact FMA
 * residual correction (correctly rounded: bit for bit the IEEE quotient the reference computes).  This runs that
 * sequence next to the device's IEEE division on `n` dividends in device memory and counts the quotients that differ. */
int ecckd_debug_div_const(int64_t n, const double *x_dev, double divisor, uint64_t *mismatches, char err[ECCKD_ERRLEN]);

/* debug: launch geometry of the column-walking kernel for a call of `ncol` columns on a device with `sm_count` SMs --
 * 16-column tiles are grouped into chunks of `tiles_per_chunk` (1 or 2) and block b of `blocks` walks chunks
 * b, b + blocks, ...  Host arithmetic only (no device needed); the CPU tests check that every tile is walked once. */
void ecckd_debug_walk_plan(int ncol, int shortwave, int sm_count, int *tiles_per_chunk, int *blocks);

/* Library identification / build info ("sm_100a", table precision, ...). */
const char *ecckd_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ECCKD_B200_H */
