// ck_loader.cpp -- netCDF-3 classic reader + ck-definition loader (host side of the product).
//
// The Fortran shim keeps the reference's own netCDF loader (north_star item 1);
// this file gives C/C++/Python hosts the same result without libnetcdf.  The file
// format is the published "NetCDF Classic Format Specification" (CDF-1 / CDF-2);
// the semantics are load_and_init's, cited line by line below.
#include "ck_loader.h"

#include <cmath>
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace ecckd {
namespace {

struct Cursor {
    const std::vector<unsigned char> &b;
    size_t pos = 0;
    explicit Cursor(const std::vector<unsigned char> &buf) : b(buf) {}
    void need(size_t n) const {
        if (pos + n > b.size()) throw std::runtime_error("netcdf3: truncated header");
    }
    uint32_t u32() {
        need(4);
        uint32_t v = (uint32_t)b[pos] << 24 | (uint32_t)b[pos + 1] << 16 | (uint32_t)b[pos + 2] << 8 | b[pos + 3];
        pos += 4;
        return v;
    }
    int32_t i32() { return (int32_t)u32(); }
    int64_t i64() {
        uint64_t hi = u32(), lo = u32();
        return (int64_t)(hi << 32 | lo);
    }
    std::string name() {
        uint32_t n = u32();
        need(n);
        std::string s(reinterpret_cast<const char *>(&b[pos]), n);
        pos += (n + 3) & ~size_t(3);
        return s;
    }
};

size_t type_size(int t)
{
    switch (t) {
        case 1: case 2: return 1;
        case 3: return 2;
        case 4: case 5: return 4;
        case 6: return 8;
    }
    throw std::runtime_error("netcdf3: unknown nc_type");
}

double decode(const unsigned char *p, int t)
{
    switch (t) {
        case 1: return (double)(int8_t)p[0];
        case 2: return (double)p[0];
        case 3: return (double)(int16_t)((uint16_t)p[0] << 8 | p[1]);
        case 4: return (double)(int32_t)((uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]);
        case 5: {
            uint32_t u = (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
            float f;
            std::memcpy(&f, &u, 4);
            return (double)f;   // f4 -> wp promotion, exact (mo_simple_netcdf.F90:60-142)
        }
        case 6: {
            uint64_t u = 0;
            for (int i = 0; i < 8; ++i) u = u << 8 | p[i];
            double d;
            std::memcpy(&d, &u, 8);
            return d;
        }
    }
    return 0.0;
}

// attribute list; returns character attributes only (numeric ones are skipped)
std::map<std::string, std::string> read_attrs(Cursor &c)
{
    std::map<std::string, std::string> out;
    uint32_t tag = c.u32(), n = c.u32();
    if (tag == 0 && n == 0) return out;
    if (tag != 0x0C) throw std::runtime_error("netcdf3: bad attribute list tag");
    for (uint32_t i = 0; i < n; ++i) {
        std::string name = c.name();
        int t = c.i32();
        uint32_t nelems = c.u32();
        size_t nbytes = type_size(t) * nelems;
        c.need(nbytes);
        if (t == 2) out[name] = std::string(reinterpret_cast<const char *>(&c.b[c.pos]), nbytes);
        c.pos += (nbytes + 3) & ~size_t(3);
    }
    return out;
}

}  // namespace

NcFile NcFile::read(const std::string &path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("load_and_init_ecckd(): can't open file " + path);   // :42-44
    std::vector<unsigned char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (buf.size() < 8 || buf[0] != 'C' || buf[1] != 'D' || buf[2] != 'F' || (buf[3] != 1 && buf[3] != 2))
        throw std::runtime_error("netcdf3: " + path + " is not a netCDF classic (CDF-1/2) file");
    const int version = buf[3];
    Cursor c(buf);
    c.pos = 4;
    (void)c.u32();   // numrecs
    NcFile nc;
    std::vector<std::string> dim_names;
    std::vector<size_t> dim_len;
    {
        uint32_t tag = c.u32(), n = c.u32();
        if (!(tag == 0 && n == 0)) {
            if (tag != 0x0A) throw std::runtime_error("netcdf3: bad dimension list tag");
            for (uint32_t i = 0; i < n; ++i) {
                std::string name = c.name();
                size_t len = c.u32();
                dim_names.push_back(name);
                dim_len.push_back(len);
                nc.dims[name] = len;
            }
        }
    }
    nc.text_attrs = read_attrs(c);
    {
        uint32_t tag = c.u32(), n = c.u32();
        if (!(tag == 0 && n == 0)) {
            if (tag != 0x0B) throw std::runtime_error("netcdf3: bad variable list tag");
            for (uint32_t i = 0; i < n; ++i) {
                std::string name = c.name();
                uint32_t nd = c.u32();
                NcVar v;
                size_t count = 1;
                for (uint32_t d = 0; d < nd; ++d) {
                    uint32_t id = c.u32();
                    if (id >= dim_names.size()) throw std::runtime_error("netcdf3: bad dimension id");
                    if (dim_len[id] == 0) throw std::runtime_error("netcdf3: record variable " + name + " not supported");
                    v.dims.push_back(dim_names[id]);
                    v.shape.push_back(dim_len[id]);
                    count *= dim_len[id];
                }
                (void)read_attrs(c);
                v.nc_type = c.i32();
                (void)c.u32();   // vsize
                int64_t begin = version == 1 ? (int64_t)c.u32() : c.i64();
                const size_t ts = type_size(v.nc_type);
                if (begin < 0 || (size_t)begin + ts * count > buf.size())
                    throw std::runtime_error("netcdf3: variable " + name + " extends past end of file");
                v.data.resize(count);
                const unsigned char *p = &buf[(size_t)begin];
                for (size_t q = 0; q < count; ++q) v.data[q] = decode(p + q * ts, v.nc_type);
                nc.vars.emplace(name, std::move(v));
            }
        }
    }
    return nc;
}

// mo_load_coefficients.F90:244-293, including the quirk that a one-character
// final token is never emitted.
std::vector<std::string> tokenize(const std::string &buffer_in)
{
    std::string buffer = buffer_in;
    while (!buffer.empty() && (buffer.back() == ' ' || buffer.back() == '\0')) buffer.pop_back();
    const int n = (int)buffer.size();
    std::vector<std::string> tokens;
    bool found = false;
    int start = 1;
    for (int i = 1; i <= n; ++i) {
        const char ch = buffer[i - 1];
        if (found) {
            if (i == n && ch != ' ') {
                if (tokens.size() + 1 > 16) throw std::runtime_error("tokenize: more tokens than can fit in array.");
                if (i - start + 1 > 32) throw std::runtime_error("tokenize: token is cut off.");
                tokens.push_back(buffer.substr(start - 1, i - start + 1));
                break;
            } else if (ch == ' ') {
                if (tokens.size() + 1 > 16) throw std::runtime_error("tokenize: more tokens than can fit in array.");
                if (i - start > 32) throw std::runtime_error("tokenize: token is cut off.");
                tokens.push_back(buffer.substr(start - 1, i - start));
                found = false;
            }
        } else if (ch != ' ') {
            found = true;
            start = i;
        }
    }
    return tokens;
}

namespace {

const NcVar &var(const NcFile &nc, const std::string &name)
{
    auto it = nc.vars.find(name);
    if (it == nc.vars.end()) throw std::runtime_error("read_field: can't find variable " + name);
    return it->second;
}

// mo_load_coefficients.F90:149-203
HostModel::Gas read_gas_input_data(const NcFile &nc, const std::string &gas_name)
{
    HostModel::Gas g;
    const std::string mf = gas_name + "_mole_fraction";
    if (nc.has(mf) && var(nc, mf).shape.size() == 1) {                                   // :162-176
        g.code = ECCKD_CODE_LOOK_UP_TABLE;
        g.mole_fraction = var(nc, mf).data;
        const NcVar &k = var(nc, gas_name + "_molar_absorption_coeff");
        if (k.shape.size() != 4) throw std::runtime_error("load_and_init_ecckd: LUT coefficient not 4d for " + gas_name);
        g.coefficient = k.data;
        return g;
    }
    const int n = (int)var(nc, gas_name + "_conc_dependence_code").data.at(0);          // :178-179
    if (n == 0) g.code = ECCKD_CODE_NONE;
    else if (n == 1) g.code = ECCKD_CODE_LINEAR;
    else if (n == 3) {
        g.code = ECCKD_CODE_RELATIVE_LINEAR;
        g.reference_mole_fraction = var(nc, gas_name + "_reference_mole_fraction").data.at(0);
    } else throw std::runtime_error("load_and_init_ecckd: bad concentration code for " + gas_name);
    const NcVar &k = var(nc, gas_name + "_molar_absorption_coeff");
    if (k.shape.size() != 3)                                                              // :194-198
        throw std::runtime_error("load_and_init_ecckd: absorption coefficient not 3d for " + gas_name);
    g.coefficient = k.data;
    return g;
}

}  // namespace

HostModel load_and_init(const std::string &filename)
{
    const NcFile nc = NcFile::read(filename);
    HostModel m;
    m.log_pressure = var(nc, "pressure").data;                                            // :46-49
    for (double &p : m.log_pressure) p = std::log(p);
    m.npress = (int)m.log_pressure.size();
    const NcVar &T = var(nc, "temperature");                                              // :51-53
    if (T.shape.size() != 2 || (int)T.shape[1] != m.npress) throw std::runtime_error("load_and_init_ecckd: bad temperature shape");
    m.temperature = T.data;   // netCDF (temperature, pressure) C order == Fortran (pressure, temperature)
    m.ntemp = (int)T.shape[0];

    const std::vector<double> &w1 = var(nc, "wavenumber1_band").data, &w2 = var(nc, "wavenumber2_band").data;
    m.nband = (int)w1.size();                                                             // :55-58
    if (m.nband < 1 || w2.size() != w1.size()) throw std::runtime_error("load_and_init_ecckd: file has no bands");
    m.band_lims_wvn.resize(2 * (size_t)m.nband);
    for (int b = 0; b < m.nband; ++b) { m.band_lims_wvn[2 * b] = w1[b]; m.band_lims_wvn[2 * b + 1] = w2[b]; }
    const std::vector<double> &bn = var(nc, "band_number").data;                          // :60-63
    const int ng = (int)bn.size();
    m.band2gpt.assign(2 * (size_t)m.nband, 0);
    m.band2gpt[0] = 1;                                                                    // :64
    m.band2gpt[2 * (m.nband - 1) + 1] = ng;                                               // :65
    int band = 1;
    for (int i = 2; i <= ng; ++i) {                                                       // :67-73
        if ((int)bn[i - 1] + 1 > band) {
            m.band2gpt[2 * (band - 1) + 1] = i - 1;
            band += 1;
            if (band > m.nband) throw std::runtime_error("load_and_init_ecckd: band_number exceeds band count");
            m.band2gpt[2 * (band - 1)] = i;
        }
    }
    const NcVar &gf = var(nc, "gpoint_fraction");                                         // :80-82
    if (gf.shape.size() != 2) throw std::runtime_error("load_and_init_ecckd: bad gpoint_fraction shape");
    m.ngpt = (int)gf.shape[0];   // size(gpoint_fraction, 2) in Fortran order

    m.shortwave = nc.has("solar_irradiance");                                             // :84
    if (m.shortwave) {
        m.solar_irradiance = var(nc, "solar_irradiance").data;                            // :86-89
        m.rayleigh = var(nc, "rayleigh_molar_scattering_coeff").data;                     // :90-93
    } else {
        m.temperature_planck = var(nc, "temperature_planck").data;                        // :95-97
        m.nplanck = (int)m.temperature_planck.size();
        m.planck_function = var(nc, "planck_function").data;                              // :98-100
    }

    auto attr = [&](const char *name) -> const std::string & {                            // :207-219
        auto it = nc.text_attrs.find(name);
        if (it == nc.text_attrs.end()) throw std::runtime_error(std::string("get_global_attribute: error reading ") + name);
        return it->second;
    };
    const std::vector<std::string> gas = tokenize(attr("constituent_id"));                // :104-107
    std::vector<std::string> composite_gas;
    bool uses_composite = false;
    for (const std::string &t : gas)                                                      // :109-117
        if (t == "composite") {
            uses_composite = true;
            composite_gas = tokenize(attr("composite_constituent_id"));
            m.num_composite_gases = (int)composite_gas.size();
            break;
        }
    for (const std::string &t : gas)                                                      // :119-126
        if (t != "composite") {
            HostModel::Gas g = read_gas_input_data(nc, t);
            g.name = t;
            g.composite_only = false;
            m.gas.push_back(std::move(g));
        }
    if (uses_composite)                                                                   // :127-143
        for (const std::string &c : composite_gas) {
            bool found = false;
            for (const std::string &t : gas) if (t == c) { found = true; break; }
            if (!found) {
                HostModel::Gas g = read_gas_input_data(nc, "composite");
                g.name = c;
                g.composite_only = true;
                m.gas.push_back(std::move(g));
            }
        }
    if (m.gas.size() > ECCKD_MAX_GASES) throw std::runtime_error("load_and_init_ecckd: more than 16 gases");
    return m;
}

ecckd_host_tables HostModel::view()
{
    gas_view.assign(gas.size(), ecckd_gas_table{});
    for (size_t i = 0; i < gas.size(); ++i) {
        ecckd_gas_table &t = gas_view[i];
        std::memset(t.name, ' ', ECCKD_NAME_LEN);
        std::memcpy(t.name, gas[i].name.data(), gas[i].name.size() < ECCKD_NAME_LEN ? gas[i].name.size() : ECCKD_NAME_LEN);
        t.composite_only = gas[i].composite_only;
        t.concentration_dependence_code = gas[i].code;
        t.n_mole_fraction = (int32_t)gas[i].mole_fraction.size();
        t.reference_mole_fraction = gas[i].reference_mole_fraction;
        t.mole_fraction = gas[i].mole_fraction.empty() ? nullptr : gas[i].mole_fraction.data();
        t.coefficient = gas[i].coefficient.data();
    }
    ecckd_host_tables h{};
    h.shortwave = shortwave;
    h.ngpt = ngpt; h.npress = npress; h.ntemp = ntemp; h.nplanck = nplanck;
    h.num_gases = (int32_t)gas.size();
    h.log_pressure = log_pressure.data();
    h.temperature = temperature.data();
    h.temperature_planck = temperature_planck.empty() ? nullptr : temperature_planck.data();
    h.planck_function = planck_function.empty() ? nullptr : planck_function.data();
    h.rayleigh_molar_scattering_coeff = rayleigh.empty() ? nullptr : rayleigh.data();
    h.solar_irradiance = solar_irradiance.empty() ? nullptr : solar_irradiance.data();
    h.gas = gas_view.data();
    h.nband = nband;
    h.band_lims_wvn = band_lims_wvn.data();
    h.band2gpt = band2gpt.data();
    return h;
}

}  // namespace ecckd
