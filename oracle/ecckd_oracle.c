/*
 * ecckd_oracle.c -- CPU restatement of the reference ECCKD gas optics (plain C).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rte-ecckd_b200/ may include, link or
 * call this file; it exists so tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs have something to check and time the
 * CUDA path against.
 *
 * PARITY UNPINNED: the reference (/root/reference/src/gas_optics_ecckd.f90) ships
 * no golden vectors or known-answer tests, and cannot be compiled in this image
 * (no Fortran compiler, no netCDF-Fortran, no rte-rrtmgp).  This file follows the
 * Fortran statement by statement -- same loop order (layer outer, column inner),
 * same per-gas (ncol,nlay,ngpt) temporary that is reallocated on every call, same
 * separate accumulate pass, same default-real literals -- and is cross-checked
 * against an independently written numpy twin (oracle/ecckd_oracle_np.py).
 *
 * All arrays are Fortran (column-major) layout, exactly as the reference sees
 * them: plev(ncol,nlay+1), tlay(ncol,nlay), tau(ncol,nlay,ngpt),
 * coefficient(ngpt,npress,ntemp,nv), temperature(npress,ntemp),
 * planck_function(ngpt,nplanck).  Indices in comments are the Fortran 1-based ones.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off for the checker so that no
 * a*b+c is fused; -O3 -march=native + pthreads for the timed baseline).
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* Fortran MAX/MIN as gfortran expands them for non-NaN operands. */
#define MAXD(a, b) ((a) > (b) ? (a) : (b))
#define MIND(a, b) ((a) < (b) ? (a) : (b))

#define ORACLE_MAX_GASES 16 /* gas_optics_ecckd.f90:24-25 */

/* concentration_dependence_code values, gas_optics_ecckd.f90:54-57 */
enum { none_ = 0, linear = 1, look_up_table = 2, relative_linear = 3 };

/* Mirror of type ty_gas_optics_ecckd + AbsorptionTable, gas_optics_ecckd.f90:13-36. */
typedef struct {
    int shortwave;
    int ngpt;      /* size(gpoint_fraction, 2) */
    int npress;    /* size(log_pressure) */
    int ntemp;     /* size(temperature, 2) */
    int nplanck;   /* size(temperature_planck) */
    int num_gases;
    const double *log_pressure;
    const double *temperature;
    const double *planck_function;
    const double *temperature_planck;
    const double *rayleigh_molar_scattering_coeff;
    const double *solar_irradiance;
    char gas[ORACLE_MAX_GASES][32];
    int composite_only[ORACLE_MAX_GASES];
    int concentration_dependence_code[ORACLE_MAX_GASES];
    int num_mole_fraction[ORACLE_MAX_GASES];
    const double *mole_fraction[ORACLE_MAX_GASES];
    double reference_mole_fraction[ORACLE_MAX_GASES];
    const double *coefficient[ORACLE_MAX_GASES];
} oracle_model;

/* Stand-in for ty_gas_concs (rte-rrtmgp, not in tree): kind 0 = scalar,
 * 1 = (nlay) profile, 2 = (ncol,nlay) field. */
typedef struct {
    int ngas;
    const char *const *names;
    const double *const *values;
    const int *kinds;
} oracle_gas_concs;

/* gas_optics_ecckd.f90:51-53: default-real literals assigned to real(wp). */
static const double gravity = (double)9.80665f;
static const double dry_air_molar_mass = (double)28.970f;
static const double pi = (double)3.14159265359f;
static const double milli = (double)0.001f; /* the literal 0.001 at :107 and :314 */

/* ty_gas_concs%get_vmr broadcasting; call site gas_optics_ecckd.f90:351. */
static void get_vmr(const oracle_gas_concs *c, int j, int ncol, int nlay, double *layer_vmr)
{
    const double *v = c->values[j];
    for (int l = 0; l < nlay; ++l)
        for (int i = 0; i < ncol; ++i) {
            double x;
            if (c->kinds[j] == 0) x = v[0];
            else if (c->kinds[j] == 1) x = v[l];
            else x = v[i + (size_t)ncol * l];
            layer_vmr[i + (size_t)ncol * l] = x;
        }
}

/* gas_optics_ecckd.f90:64-241, logarithmic_interpolation = .false. (:368-369; the
 * log branches are dead code because tables hold exact zeros).  When idx_* are
 * non-NULL the interpolation indices/weights are also reported (test hook; the
 * arithmetic is the same statements). */
static void calculate_optical_depth(const oracle_model *this, int gas, int num_columns, int num_layers,
                                    const double *level_pressure, const double *layer_temperature,
                                    const double *layer_vmr, double **optical_depth_p,
                                    int *idx_ip0, double *idx_pw1, int *idx_it0, double *idx_tw1,
                                    int *idx_iv0, double *idx_vw1)
{
    const int num_gpoints = this->ngpt;
    const int np = this->npress, nt = this->ntemp;
    const double *coef = this->coefficient[gas];
    const int code = this->concentration_dependence_code[gas];
    const size_t s_p = (size_t)num_gpoints, s_t = s_p * np, s_v = s_t * nt;
#define COEF(ip, it, iv) (coef + s_p * ((ip) - 1) + s_t * ((it) - 1) + s_v * ((iv) - 1))
#define TEMP(ip, it) (this->temperature[((ip) - 1) + (size_t)np * ((it) - 1)])
#define LP(i, j) (level_pressure[((i) - 1) + (size_t)num_columns * ((j) - 1)])
#define L2(a, i, j) ((a)[((i) - 1) + (size_t)num_columns * ((j) - 1)])

    const double log_p_0 = this->log_pressure[0];                                  /* :104 */
    const double d_log_p = this->log_pressure[1] - this->log_pressure[0];          /* :105 */
    const double dt = TEMP(1, 2) - TEMP(1, 1);                                     /* :106 */
    const double global_weight = 1. / (gravity * milli * dry_air_molar_mass);      /* :107 */

    if (*optical_depth_p) free(*optical_depth_p);                                  /* :112-115 */
    double *optical_depth = (double *)malloc(sizeof(double) * (size_t)num_columns * num_layers * num_gpoints);
    *optical_depth_p = optical_depth;
    const size_t s_g = (size_t)num_columns * num_layers;

    for (int j = 1; j <= num_layers; ++j) {                                        /* :117 */
        for (int i = 1; i <= num_columns; ++i) {                                   /* :118 */
            double log_pressure = log(0.5 * (LP(i, j + 1) + LP(i, j)));            /* :120 */
            double pressure_index = (log_pressure - log_p_0) / d_log_p;            /* :121 */
            pressure_index = 1. + MAXD(0., MIND(pressure_index, (double)np - 1.0001)); /* :122-125 */
            int ip0 = (int)pressure_index;                                         /* :126 */
            double pressure_weight1 = pressure_index - ip0;
            double pressure_weight0 = 1. - pressure_weight1;

            double t0 = pressure_weight0 * TEMP(ip0, 1) + pressure_weight1 * TEMP(ip0 + 1, 1); /* :131-132 */
            double temperature_index = (L2(layer_temperature, i, j) - t0) / dt;    /* :133 */
            temperature_index = 1. + MAXD(0., MIND(temperature_index, (double)nt - 1.0001));
            int it0 = (int)temperature_index;                                      /* :138 */
            double temperature_weight1 = temperature_index - it0;
            double temperature_weight0 = 1. - temperature_weight1;

            double simple_weight = global_weight * (LP(i, j + 1) - LP(i, j));      /* :143 */
            double weight = 0.;
            if (code == relative_linear)
                weight = simple_weight * (L2(layer_vmr, i, j) - this->reference_mole_fraction[gas]);
            else
                weight = simple_weight * L2(layer_vmr, i, j);

            double *od = optical_depth + (i - 1) + (size_t)num_columns * (j - 1);
            int iv0 = 1;
            double vmr_weight1 = 0.;
            if (code == look_up_table) {
                const double *mf = this->mole_fraction[gas];
                double log_vmr = log(MAXD(L2(layer_vmr, i, j), mf[0]));            /* :153 */
                double d_log_vmr = log(mf[1] / mf[0]);                             /* :154-155 */
                double vmr_index = (log_vmr - log(mf[0])) / d_log_vmr;             /* :156 */
                vmr_index = 1. + MAXD(0., MIND(vmr_index, (double)this->num_mole_fraction[gas] - 1.001));
                iv0 = (int)vmr_index;                                              /* :161 */
                vmr_weight1 = vmr_index - iv0;
                double vmr_weight0 = 1. - vmr_weight1;
                const double *k000 = COEF(ip0, it0, iv0), *k100 = COEF(ip0 + 1, it0, iv0);
                const double *k010 = COEF(ip0, it0 + 1, iv0), *k110 = COEF(ip0 + 1, it0 + 1, iv0);
                const double *k001 = COEF(ip0, it0, iv0 + 1), *k101 = COEF(ip0 + 1, it0, iv0 + 1);
                const double *k011 = COEF(ip0, it0 + 1, iv0 + 1), *k111 = COEF(ip0 + 1, it0 + 1, iv0 + 1);
                for (int k = 0; k < num_gpoints; ++k)                              /* :167-178 */
                    od[s_g * k] = weight * (vmr_weight0 * (temperature_weight0 * (
                        pressure_weight0 * k000[k] + pressure_weight1 * k100[k]) +
                        temperature_weight1 * (
                        pressure_weight0 * k010[k] + pressure_weight1 * k110[k])) +
                        vmr_weight1 * (temperature_weight0 * (
                        pressure_weight0 * k001[k] + pressure_weight1 * k101[k]) +
                        temperature_weight1 * (
                        pressure_weight0 * k011[k] + pressure_weight1 * k111[k])));
            } else {
                const double *k00 = COEF(ip0, it0, 1), *k10 = COEF(ip0 + 1, it0, 1);
                const double *k01 = COEF(ip0, it0 + 1, 1), *k11 = COEF(ip0 + 1, it0 + 1, 1);
                /* :198-203 (linear / relative_linear) use weight; :216-221 (none_) use simple_weight */
                const double w = (code == linear || code == relative_linear) ? weight : simple_weight;
                for (int k = 0; k < num_gpoints; ++k)
                    od[s_g * k] = w * (temperature_weight0 * (
                        pressure_weight0 * k00[k] + pressure_weight1 * k10[k]) +
                        temperature_weight1 * (
                        pressure_weight0 * k01[k] + pressure_weight1 * k11[k]));
            }
            for (int k = 0; k < num_gpoints; ++k)                                  /* :233-238 */
                if (od[s_g * k] < 0.) od[s_g * k] = 0.;

            if (idx_ip0) {
                size_t c = (size_t)(i - 1) + (size_t)num_columns * (j - 1);
                idx_ip0[c] = ip0; idx_pw1[c] = pressure_weight1;
                idx_it0[c] = it0; idx_tw1[c] = temperature_weight1;
                idx_iv0[c] = iv0; idx_vw1[c] = vmr_weight1;
            }
        }
    }
#undef COEF
#undef TEMP
#undef LP
#undef L2
}

/* gas_optics_ecckd.f90:245-289 */
static void calculate_planck_function(const oracle_model *this, int num_columns, int num_levels,
                                      const double *level_temperature, double **planck_p,
                                      int *idx_it0, double *idx_w1)
{
    const int num_gpoints = this->ngpt;
    if (*planck_p) free(*planck_p);                                                /* :266-269 */
    double *planck = (double *)malloc(sizeof(double) * (size_t)num_columns * num_levels * num_gpoints);
    *planck_p = planck;
    const size_t s_g = (size_t)num_columns * num_levels;
    const double dt = this->temperature_planck[1] - this->temperature_planck[0];   /* :271 */
    const double t0 = this->temperature_planck[0];                                 /* :272 */
    for (int j = 1; j <= num_levels; ++j) {
        for (int i = 1; i <= num_columns; ++i) {
            const double T = level_temperature[(i - 1) + (size_t)num_columns * (j - 1)];
            double *p = planck + (i - 1) + (size_t)num_columns * (j - 1);
            double temperature_index = (T - t0) / dt;                              /* :275 */
            int it0 = 0;
            double temperature_weight1 = 0.;
            if (temperature_index >= 0) {
                temperature_index = 1. + temperature_index;
                it0 = (int)temperature_index;
                if (it0 > this->nplanck - 1) it0 = this->nplanck - 1;              /* :278 */
                temperature_weight1 = temperature_index - it0;
                double temperature_weight0 = 1. - temperature_weight1;
                const double *p0 = this->planck_function + (size_t)num_gpoints * (it0 - 1);
                const double *p1 = this->planck_function + (size_t)num_gpoints * it0;
                for (int k = 0; k < num_gpoints; ++k)                              /* :281-282 */
                    p[s_g * k] = temperature_weight0 * p0[k] + temperature_weight1 * p1[k];
            } else {
                for (int k = 0; k < num_gpoints; ++k)                              /* :284 */
                    p[s_g * k] = (T / t0) * this->planck_function[k];
            }
            if (idx_it0) {
                idx_it0[(i - 1) + (size_t)num_columns * (j - 1)] = it0;
                idx_w1[(i - 1) + (size_t)num_columns * (j - 1)] = temperature_weight1;
            }
        }
    }
    const size_t n = s_g * num_gpoints;                                            /* :288 */
    for (size_t q = 0; q < n; ++q) planck[q] = planck[q] / pi;
}

/* gas_optics_ecckd.f90:293-319 */
static void calculate_rayleigh_optical_depth(const oracle_model *this, int num_columns, int num_layers,
                                             const double *level_pressure, double **optical_depth_p)
{
    const int num_gpoints = this->ngpt;
    if (*optical_depth_p) free(*optical_depth_p);
    double *optical_depth = (double *)malloc(sizeof(double) * (size_t)num_columns * num_layers * num_gpoints);
    *optical_depth_p = optical_depth;
    const size_t n2 = (size_t)num_columns * num_layers;
    double *moles_per_layer = (double *)malloc(sizeof(double) * n2);
    for (int j = 0; j < num_layers; ++j)                                           /* :313-314 */
        for (int i = 0; i < num_columns; ++i)
            moles_per_layer[i + (size_t)num_columns * j] =
                (level_pressure[i + (size_t)num_columns * (j + 1)] - level_pressure[i + (size_t)num_columns * j]) *
                (1. / (gravity * milli * dry_air_molar_mass));
    for (int k = 0; k < num_gpoints; ++k)                                          /* :315-317 */
        for (size_t q = 0; q < n2; ++q)
            optical_depth[q + n2 * k] = moles_per_layer[q] * this->rayleigh_molar_scattering_coeff[k];
    free(moles_per_layer);
}

static void trimmed(char *dst, const char *src) /* trim() of a blank-padded character(len=32) */
{
    size_t n = strnlen(src, 31);
    while (n > 0 && src[n - 1] == ' ') --n;
    memcpy(dst, src, n);
    dst[n] = 0;
}

/* gas_optics_ecckd.f90:323-376 */
static int gas_optical_depth(const oracle_model *this, int ncol, int nlay, const double *plev, const double *tlay,
                             const oracle_gas_concs *gas_desc, double *tau, char *error_msg)
{
    const size_t n3 = (size_t)ncol * nlay * this->ngpt;
    const int n = gas_desc->ngas;                                                  /* :340 */
    double *layer_vmr = (double *)malloc(sizeof(double) * (size_t)ncol * nlay);     /* :343 */
    double *optical_depth = (double *)calloc(n3, sizeof(double));                   /* :344-345 */
    memset(tau, 0, sizeof(double) * n3);                                            /* :346 */
    int first_calc = 1;
    error_msg[0] = 0;
    for (int j = 0; j < n; ++j) {                                                  /* :348 */
        char want[32], have[32];
        trimmed(want, gas_desc->names[j]);
        int i;
        for (i = 0; i < this->num_gases; ++i) {                                    /* :349-357 */
            trimmed(have, this->gas[i]);
            if (strcmp(have, want) == 0) {
                get_vmr(gas_desc, j, ncol, nlay, layer_vmr);
                break;
            }
        }
        if (i >= this->num_gases) continue;                                        /* :358-364 */
        if (this->composite_only[i] && !first_calc) continue;                      /* :365-367 */
        calculate_optical_depth(this, i, ncol, nlay, plev, tlay, layer_vmr, &optical_depth,
                                NULL, NULL, NULL, NULL, NULL, NULL);              /* :368-369 */
        for (size_t q = 0; q < n3; ++q) tau[q] = tau[q] + optical_depth[q];         /* :370 */
        if (this->composite_only[i]) first_calc = 0;                               /* :371-373 */
    }
    free(layer_vmr);
    free(optical_depth);
    return 0;
}

/* gas_optics_ecckd.f90:381-426.  tlev may be NULL (absent optional): the error is
 * raised only after tau, lay_source and sfc_source are written (:414-417). */
int oracle_gas_optics_int(const oracle_model *this, int ncol, int nlay, const double *plev, const double *tlay,
                          const double *tsfc, const double *tlev, const oracle_gas_concs *gas_desc,
                          double *tau, double *lay_source, double *lev_source_inc, double *lev_source_dec,
                          double *sfc_source, char error_msg[128])
{
    const int ngpt = this->ngpt;
    gas_optical_depth(this, ncol, nlay, plev, tlay, gas_desc, tau, error_msg);      /* :401 */
    if (error_msg[0]) return 1;

    double *buffer = NULL;
    calculate_planck_function(this, ncol, nlay, tlay, &buffer, NULL, NULL);         /* :407 */
    memcpy(lay_source, buffer, sizeof(double) * (size_t)ncol * nlay * ngpt);
    free(buffer); buffer = NULL;
    calculate_planck_function(this, ncol, 1, tsfc, &buffer, NULL, NULL);            /* :408-412 */
    memcpy(sfc_source, buffer, sizeof(double) * (size_t)ncol * ngpt);
    free(buffer); buffer = NULL;
    if (!tlev) {                                                                    /* :414-417 */
        snprintf(error_msg, 128, "tlev is required for ecckd");
        return 1;
    }
    calculate_planck_function(this, ncol, nlay + 1, tlev, &buffer, NULL, NULL);     /* :419-422 */
    for (int k = 0; k < ngpt; ++k)
        for (int j = 0; j < nlay; ++j)
            for (int i = 0; i < ncol; ++i) {
                size_t o = i + (size_t)ncol * (j + (size_t)nlay * k);
                lev_source_inc[o] = buffer[i + (size_t)ncol * ((j + 1) + (size_t)(nlay + 1) * k)]; /* :423 */
                lev_source_dec[o] = buffer[i + (size_t)ncol * (j + (size_t)(nlay + 1) * k)];       /* :424 */
            }
    free(buffer);
    return 0;
}

/* gas_optics_ecckd.f90:431-473.  two_stream == 0 models a non-2str optical_props
 * argument: tau (+rayleigh) is written, then the error at :461-463. */
int oracle_gas_optics_ext(const oracle_model *this, int ncol, int nlay, const double *plev, const double *tlay,
                          const oracle_gas_concs *gas_desc, int two_stream,
                          double *tau, double *ssa, double *g, double *toa_src, char error_msg[128])
{
    const size_t n3 = (size_t)ncol * nlay * this->ngpt;
    gas_optical_depth(this, ncol, nlay, plev, tlay, gas_desc, tau, error_msg);      /* :449 */
    if (error_msg[0]) return 1;
    double *optical_depth = NULL;
    calculate_rayleigh_optical_depth(this, ncol, nlay, plev, &optical_depth);       /* :455 */
    for (size_t q = 0; q < n3; ++q) tau[q] = tau[q] + optical_depth[q];             /* :456 */
    if (!two_stream) {
        free(optical_depth);
        snprintf(error_msg, 128, "shortwave must use ty_optical_props_2str");
        return 1;
    }
    for (size_t q = 0; q < n3; ++q) ssa[q] = optical_depth[q] / tau[q];             /* :459 */
    for (size_t q = 0; q < n3; ++q) g[q] = 0;                                       /* :460 */
    free(optical_depth);
    for (int j = 0; j < this->ngpt; ++j)                                            /* :468-472 */
        for (int i = 0; i < ncol; ++i) toa_src[i + (size_t)ncol * j] = this->solar_irradiance[j];
    return 0;
}

/* Test hook: the interpolation indices the reference computes for model gas
 * `gas` (0-based position in this->gas) -- same statements as above. */
void oracle_indices(const oracle_model *this, int gas, int ncol, int nlay, const double *plev, const double *tlay,
                    const double *layer_vmr, int *ip0, double *pw1, int *it0, double *tw1, int *iv0, double *vw1)
{
    double *od = NULL;
    calculate_optical_depth(this, gas, ncol, nlay, plev, tlay, layer_vmr, &od, ip0, pw1, it0, tw1, iv0, vw1);
    free(od);
}

void oracle_planck_indices(const oracle_model *this, int ncol, int nlev, const double *temperature, int *it0, double *w1)
{
    double *b = NULL;
    calculate_planck_function(this, ncol, nlev, temperature, &b, it0, w1);
    free(b);
}

/* Constants the tests pin (SURVEY.md 8c "hand-derivable checkpoints"). */
void oracle_constants(double out[5])
{
    out[0] = gravity; out[1] = dry_air_molar_mass; out[2] = pi; out[3] = milli;
    out[4] = 1. / (gravity * milli * dry_air_molar_mass);
}

/* ---------------------------------------------------------------------------
 * Timed baseline: the reference driver's loop over independent column blocks
 * (ecckd_rfmip_lw.F90:107-136, ecckd_rfmip_sw.F90:112-162), with the blocks
 * handed out to POSIX threads (this image's gcc has no libgomp).  Inputs are
 * already blocked (block, nlay[+1], nblocks) like read_and_block_pt produces;
 * every thread owns one set of output arrays that each block overwrites
 * (optical_props / source are allocated once outside the loop,
 * ecckd_rfmip_lw.F90:102-103).  2-D gases are blocked too; scalars stay scalars.
 * Returns the number of blocks processed.
 * ------------------------------------------------------------------------- */
typedef struct {
    const oracle_model *model;
    int lw, block, nlay, ngas;
    long nblocks;
    const double *plev, *tlay, *tsfc, *tlev;
    const char *const *names;
    const double *const *values;
    const int *kinds;
    atomic_long *next;
    double total;
} block_job;

static void *block_worker(void *arg)
{
    block_job *jb = (block_job *)arg;
    const oracle_model *this = jb->model;
    const int block = jb->block, nlay = jb->nlay, ngas = jb->ngas;
    const size_t n3 = (size_t)block * nlay * this->ngpt, n2 = (size_t)block * this->ngpt;
    double *a = (double *)malloc(sizeof(double) * n3), *b = (double *)malloc(sizeof(double) * n3);
    double *c = (double *)malloc(sizeof(double) * n3), *d = (double *)malloc(sizeof(double) * n3);
    double *e = (double *)malloc(sizeof(double) * n2);
    const double **vals = (const double **)malloc(sizeof(double *) * (ngas > 0 ? ngas : 1));
    char err[128];
    double total = 0.;
    for (;;) {
        long blk = atomic_fetch_add(jb->next, 1);
        if (blk >= jb->nblocks) break;
        for (int q = 0; q < ngas; ++q)
            vals[q] = jb->values[q] + (jb->kinds[q] == 2 ? (size_t)block * nlay * blk : 0);
        oracle_gas_concs gc = { ngas, jb->names, vals, jb->kinds };
        const double *pl = jb->plev + (size_t)block * (nlay + 1) * blk;
        const double *tl = jb->tlay + (size_t)block * nlay * blk;
        if (jb->lw)
            oracle_gas_optics_int(this, block, nlay, pl, tl, jb->tsfc + (size_t)block * blk,
                                  jb->tlev + (size_t)block * (nlay + 1) * blk, &gc, a, b, c, d, e, err);
        else
            oracle_gas_optics_ext(this, block, nlay, pl, tl, &gc, 1, a, b, c, e, err);
        total += a[n3 - 1] + b[0];
    }
    free(a); free(b); free(c); free(d); free(e); free(vals);
    jb->total = total;
    return NULL;
}

long oracle_run_blocks(const oracle_model *this, int lw, int block, int nlay, long nblocks,
                       const double *plev, const double *tlay, const double *tsfc, const double *tlev,
                       int ngas, const char *const *names, const double *const *values, const int *kinds,
                       int nthreads, double *checksum)
{
    if (nthreads < 1) nthreads = 1;
    atomic_long next = 0;
    block_job *jobs = (block_job *)malloc(sizeof(block_job) * nthreads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 0; t < nthreads; ++t) {
        block_job jb = { this, lw, block, nlay, ngas, nblocks, plev, tlay, tsfc, tlev, names, values, kinds, &next, 0. };
        jobs[t] = jb;
        if (t > 0) pthread_create(&th[t], NULL, block_worker, &jobs[t]);
    }
    block_worker(&jobs[0]);
    double total = jobs[0].total;
    for (int t = 1; t < nthreads; ++t) { pthread_join(th[t], NULL); total += jobs[t].total; }
    free(jobs); free(th);
    if (checksum) *checksum = total;
    return nblocks;
}

int oracle_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
