// ck_loader.h -- host-side ck-definition loader of the product:
// a netCDF-3 classic reader plus load_and_init semantics
// (/root/reference/example/rfmip-rad-irf/mo_load_coefficients.F90:19-203).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "ecckd_b200.h"

namespace ecckd {

struct NcVar {
    std::vector<std::string> dims;
    std::vector<size_t> shape;
    int nc_type = 0;             // 1 byte, 2 char, 3 short, 4 int, 5 float, 6 double
    std::vector<double> data;    // numeric values promoted to double (exact for all of these types)
};

struct NcFile {
    std::map<std::string, size_t> dims;
    std::map<std::string, std::string> text_attrs;   // global character attributes
    std::map<std::string, NcVar> vars;
    // throws std::runtime_error
    static NcFile read(const std::string &path);
    bool has(const std::string &v) const { return vars.count(v) != 0; }
};

// Everything load_and_init leaves in ty_gas_optics_ecckd, owning its storage;
// view() exposes it as the C ABI's ecckd_host_tables.
struct HostModel {
    bool shortwave = false;
    int ngpt = 0, npress = 0, ntemp = 0, nplanck = 0, nband = 0, num_composite_gases = 0;
    std::vector<double> log_pressure, temperature, temperature_planck, planck_function;
    std::vector<double> rayleigh, solar_irradiance, band_lims_wvn;
    std::vector<int32_t> band2gpt;
    struct Gas {
        std::string name;
        bool composite_only = false;
        int code = 0;
        double reference_mole_fraction = 0.0;
        std::vector<double> mole_fraction, coefficient;
    };
    std::vector<Gas> gas;

    std::vector<ecckd_gas_table> gas_view;   // filled by view()
    ecckd_host_tables view();
};

std::vector<std::string> tokenize(const std::string &buffer);   // mo_load_coefficients.F90:244-293
HostModel load_and_init(const std::string &filename);           // mo_load_coefficients.F90:19-146

}  // namespace ecckd
