/* shim_sequence.c -- replays, in plain C, exactly what the Fortran shim
 * (rte-ecckd_b200/fortran/gas_optics_ecckd_b200_shim.f90) does with the C ABI of libecckd_b200.so:
 *
 *   1. create_model():   ecckd_model_create from HOST tables with nband = 0, NULL band arrays and
 *                        blank-padded (not NUL-terminated) 32-character gas names, one device;
 *   2. gas_optics_int(): ecckd_gas_optics_lw_host_f64 on PAGEABLE (malloc) arrays, gases marshalled as the shim
 *                        does (constant fields as scalar views, others as (ncol,nlay) fields), block_cols = 0;
 *      gas_optics_ext(): ecckd_gas_optics_sw_host_f64 likewise;
 *   3. the reference's error strings: tlev absent -> "tlev is required for ecckd" after the partial write,
 *      ssa/g absent -> "shortwave must use ty_optical_props_2str".
 *
 * The tables come from the library's own loader (ecckd_tables_load) and are then re-packed the way the
 * shim packs them.  Inputs and outputs are written to a flat file; tests/test_shim_sequence.py checks them
 * against the oracle.  No Python, no torch in this translation unit: gcc + the public header only.
 *
 * usage: shim_sequence <ck-definition.nc> <out.bin> [ncol nlay]       exit code 0 = every step behaved
 *        shim_sequence --no-device <ck-definition.nc>                  expects ECCKD_ERR_CUDA (CPU-only hosts) */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ecckd_b200.h"

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "shim_sequence: " __VA_ARGS__); fprintf(stderr, "\n"); return 1; } } while (0)

static double frand(unsigned long long *s)   /* splitmix64 -> U(0,1) */
{
    unsigned long long z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

/* the shim's create_model(): copy the loader's tables into a fresh ecckd_host_tables */
static int shim_create_model(const ecckd_host_tables *src, ecckd_model **model, char err[ECCKD_ERRLEN])
{
    ecckd_host_tables t;
    ecckd_gas_table g[ECCKD_MAX_GASES];
    memset(&t, 0, sizeof t);
    t.shortwave = src->shortwave; t.ngpt = src->ngpt; t.npress = src->npress; t.ntemp = src->ntemp;
    t.nplanck = src->shortwave ? 0 : src->nplanck;
    t.num_gases = src->num_gases;
    t.log_pressure = src->log_pressure; t.temperature = src->temperature;
    t.temperature_planck = src->shortwave ? NULL : src->temperature_planck;
    t.planck_function = src->shortwave ? NULL : src->planck_function;
    t.rayleigh_molar_scattering_coeff = src->shortwave ? src->rayleigh_molar_scattering_coeff : NULL;
    t.solar_irradiance = src->shortwave ? src->solar_irradiance : NULL;
    for (int i = 0; i < src->num_gases; ++i) {
        g[i] = src->gas[i];
        size_t n = strnlen(src->gas[i].name, ECCKD_NAME_LEN);
        memset(g[i].name, ' ', ECCKD_NAME_LEN);              /* character(len=32): blank padded, no NUL */
        memcpy(g[i].name, src->gas[i].name, n);
    }
    t.gas = g;
    t.nband = 0; t.band_lims_wvn = NULL; t.band2gpt = NULL;  /* kept by ty_optical_props%init in the shim */
    return ecckd_model_create(&t, 1, NULL, model, err);
}

int main(int argc, char **argv)
{
    char err[ECCKD_ERRLEN];
    const ecckd_host_tables *tab = NULL;
    if (argc >= 3 && strcmp(argv[1], "--no-device") == 0) {
        CHECK(ecckd_tables_load(argv[2], &tab, err) == ECCKD_OK, "tables_load: %s", err);
        ecckd_model *m = NULL;
        int rc = shim_create_model(tab, &m, err);
        CHECK(rc == ECCKD_ERR_CUDA && m == NULL && strlen(err) > 0, "expected ECCKD_ERR_CUDA without a device, got %d '%s'", rc, err);
        printf("no device: model_create refused loudly: %s\n", err);
        ecckd_tables_free(tab);
        return 0;
    }
    CHECK(argc >= 3, "usage: shim_sequence <ck.nc> <out.bin> [ncol nlay]");
    const int ncol = argc > 3 ? atoi(argv[3]) : 37, nlay = argc > 4 ? atoi(argv[4]) : 60;
    CHECK(ecckd_tables_load(argv[1], &tab, err) == ECCKD_OK, "tables_load: %s", err);
    const int ngpt = tab->ngpt, sw = tab->shortwave;
    ecckd_model *model = NULL;
    CHECK(shim_create_model(tab, &model, err) == ECCKD_OK, "model_create: %s", err);
    CHECK(ecckd_get_ngpt(model) == ngpt && ecckd_get_nband(model) == 0, "model extents");

    /* pageable host arrays, Fortran layout (column fastest), TOA first */
    const size_t n2 = (size_t)ncol * nlay, n2p = (size_t)ncol * (nlay + 1), n3 = n2 * ngpt;
    double *plev = malloc(n2p * 8), *tlay = malloc(n2 * 8), *tlev = malloc(n2p * 8), *tsfc = malloc((size_t)ncol * 8);
    double *h2o = malloc(n2 * 8), *o3 = malloc(n2 * 8), *co2f = malloc(n2 * 8);
    double *out[4], *out2d = malloc((size_t)ncol * ngpt * 8);
    for (int o = 0; o < 4; ++o) out[o] = malloc(n3 * 8);
    unsigned long long seed = 20260101;
    for (int c = 0; c < ncol; ++c) {
        const double psfc = 5.0e4 + 5.5e4 * frand(&seed);
        for (int k = 0; k <= nlay; ++k) {
            const double s = (double)k / nlay;
            plev[c + (size_t)ncol * k] = 1.0 + (psfc - 1.0) * pow(s, 2.5);
        }
        for (int k = 0; k < nlay; ++k) {
            const double p = 0.5 * (plev[c + (size_t)ncol * k] + plev[c + (size_t)ncol * (k + 1)]);
            tlay[c + (size_t)ncol * k] = 200.0 + 90.0 * pow(p / 1.05e5, 0.19) + 10.0 * (frand(&seed) - 0.5);
            h2o[c + (size_t)ncol * k] = 1e-7 + 2e-2 * pow(p / psfc, 3.0) * frand(&seed);
            o3[c + (size_t)ncol * k] = 2e-8 + 8e-6 * exp(-0.5 * pow(log(p / 1000.0), 2.0));
            co2f[c + (size_t)ncol * k] = 397.5e-6;           /* a (ncol,nlay) field that happens to be constant */
        }
        for (int k = 0; k <= nlay; ++k) {
            const int a = k > 0 ? k - 1 : 0, b = k < nlay ? k : nlay - 1;
            tlev[c + (size_t)ncol * k] = 0.5 * (tlay[c + (size_t)ncol * a] + tlay[c + (size_t)ncol * b]);
        }
        tsfc[c] = tlev[c + (size_t)ncol * nlay] + 2.0 * (frand(&seed) - 0.5);
    }
    /* marshal_gases(): the RFMIP driver's order; constant fields become scalar views (min == max) */
    const char *names[9] = {"co2", "ch4", "n2o", "o2", "cfc11", "cfc12", "h2o", "o3", "no2"};
    static double ch4 = 1831e-9 * 1.2, n2o = 327e-9 * 1.1, o2 = 0.209, cfc11 = 233e-12, cfc12 = 520e-12, no2 = 0.0;
    ecckd_vmr_view v[9] = {{co2f, 0, 0}, {&ch4, 0, 0}, {&n2o, 0, 0}, {&o2, 0, 0}, {&cfc11, 0, 0}, {&cfc12, 0, 0},
                           {h2o, 1, ncol}, {o3, 1, ncol}, {&no2, 0, 0}};
    int rc;
    if (!sw) {
        rc = ecckd_gas_optics_lw_host_f64(model, ncol, nlay, plev, tlay, NULL, tsfc, 9, names, v, out[0], out[1], out[2], out[3], out2d, 0, err);
        CHECK(rc == ECCKD_ERR_REFERENCE && strcmp(err, "tlev is required for ecckd") == 0, "missing tlev: rc %d '%s'", rc, err);
        CHECK(out[0][0] > 0.0 && out[1][0] > 0.0 && out2d[0] > 0.0, "tau / lay_source / sfc_source are written before the tlev error (:414-417)");
        rc = ecckd_gas_optics_lw_host_f64(model, ncol, nlay, plev, tlay, tlev, tsfc, 9, names, v, out[0], out[1], out[2], out[3], out2d, 0, err);
        CHECK(rc == ECCKD_OK && err[0] == '\0', "lw: rc %d '%s'", rc, err);
    } else {
        rc = ecckd_gas_optics_sw_host_f64(model, ncol, nlay, plev, tlay, 9, names, v, out[0], NULL, NULL, out2d, 0, err);
        CHECK(rc == ECCKD_ERR_REFERENCE && strcmp(err, "shortwave must use ty_optical_props_2str") == 0, "1scl props: rc %d '%s'", rc, err);
        rc = ecckd_gas_optics_sw_host_f64(model, ncol, nlay, plev, tlay, 9, names, v, out[0], out[1], out[2], out2d, 0, err);
        CHECK(rc == ECCKD_OK && err[0] == '\0', "sw: rc %d '%s'", rc, err);
    }
    /* argument errors instead of crashes */
    rc = ecckd_gas_optics_lw_host_f64(model, ncol, nlay, plev, tlay, tlev, tsfc, 9, NULL, v, out[0], out[1], out[2], out[3], out2d, 0, err);
    CHECK(rc == ECCKD_ERR_ARG, "NULL names must be an argument error, got %d", rc);
    rc = ecckd_gas_optics_sw_host_f64(model, ncol, nlay, plev, tlay, 9, names, NULL, out[0], out[1], out[2], out2d, 0, err);
    CHECK(rc == ECCKD_ERR_ARG, "NULL vmr must be an argument error, got %d", rc);

    FILE *f = fopen(argv[2], "wb");
    CHECK(f != NULL, "cannot write %s", argv[2]);
    const int hdr[4] = {ncol, nlay, ngpt, sw};
    fwrite(hdr, sizeof hdr, 1, f);
    fwrite(plev, 8, n2p, f); fwrite(tlay, 8, n2, f); fwrite(tlev, 8, n2p, f); fwrite(tsfc, 8, ncol, f);
    fwrite(h2o, 8, n2, f); fwrite(o3, 8, n2, f);
    const double scal[7] = {397.5e-6, ch4, n2o, o2, cfc11, cfc12, no2};
    fwrite(scal, 8, 7, f);
    for (int o = 0; o < (sw ? 3 : 4); ++o) fwrite(out[o], 8, n3, f);
    fwrite(out2d, 8, (size_t)ncol * ngpt, f);
    fclose(f);
    ecckd_model_destroy(model);
    ecckd_tables_free(tab);
    printf("shim sequence ok: %s, %d columns x %d layers x %d g-points\n", sw ? "shortwave" : "longwave", ncol, nlay, ngpt);
    return 0;
}
