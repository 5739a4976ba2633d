// ecckd_device.cuh -- device-side model description and the per-(column,layer)
// index arithmetic shared by every kernel.
//
// The index arithmetic restates /root/reference/src/gas_optics_ecckd.f90:104-107,
// 120-163 and :271-280 operation by operation with the round-to-nearest
// intrinsics (__dadd_rn, __dmul_rn, __ddiv_rn), which nvcc never contracts into
// FMAs: apart from log() (CUDA's vs libm's, both < 1 ulp) every intermediate is
// the IEEE value the Fortran computes, so ip0/it0/iv0 are bit-exact.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ecckd {

constexpr int kMaxGases = 16;  // gas_optics_ecckd.f90:24
constexpr int kMaxBil = 8;     // bilinear (code 0/1/3) table slots the walking kernel keeps in registers

// Padded "walk" table geometry: every dimension that enters a corner offset is a
// compile-time constant, so the 4 (8) corner loads of every slot are immediate
// offsets from ONE base pointer per (ip0,it0[,iv0]) change.
constexpr int kGP = 32;                    // g-points per group = lanes per warp
constexpr int kNTP = 8;                    // padded temperature extent   (ntemp  <= 8)
constexpr int kNPP = 64;                   // padded pressure extent      (npress <= 64)
constexpr int kSlab = kNPP * kNTP * kGP;   // elements of one (slot | mole-fraction) slab
// Position of g-point l (0..31) of a group inside a 32-element table row.  Rows are stored permuted so that the four
// g-points {q, q+8, q+16, q+24} of one thread of the quad kernel (8 lanes per column, 4 g-points per lane) come from
// 16-byte vector loads that are contiguous across the 8 lanes, while their rows in the 128-byte-swizzled staging
// tile (swizzle key = g & 7 = q) stay conflict-free:
//   4-byte rows (128 B): one 16-byte load per lane          -> [q][i]          i = g >> 3, q = g & 7
//   8-byte rows (256 B): two 16-byte loads per lane, each a contiguous 128-byte half row -> [i >> 1][q][i & 1]
__host__ __device__ constexpr int walk_pos32(int l) { return ((l & 7) << 2) | (l >> 3); }
__host__ __device__ constexpr int walk_pos64(int l) { return ((l >> 4) << 4) | ((l & 7) << 1) | ((l >> 3) & 1); }
__host__ __device__ constexpr int walk_inv32(int p) { return (p >> 2) | ((p & 3) << 3); }
__host__ __device__ constexpr int walk_inv64(int p) { return ((p >> 1) & 7) | ((((p >> 4) << 1) | (p & 1)) << 3); }
template <typename T> __host__ __device__ constexpr int walk_pos(int l) { return sizeof(T) == 8 ? walk_pos64(l) : walk_pos32(l); }
template <typename T> __host__ __device__ constexpr int walk_inv(int p) { return sizeof(T) == 8 ? walk_inv64(p) : walk_inv32(p); }

enum Code : int { kNone = 0, kLinear = 1, kLut = 2, kRelLinear = 3 };  // :54-57

// Device-resident packed tables (one copy per GPU, read-only after create).
struct DevModel {
    int ngpt, np, nt, nplanck;
    int sm_count;      // multiprocessors of the device this copy lives on (persistent-grid size)
    double logp0;      // log_pressure(1)                              :104
    double dlogp;      // log_pressure(2) - log_pressure(1)            :105
    double dt;         // temperature(1,2) - temperature(1,1)          :106
    double gw;         // 1./(gravity*0.001*dry_air_molar_mass)        :107 (default-real literals)
    double p_clamp;    // size(log_pressure) - 1.0001_wp               :125
    double t_clamp;    // size(temperature,2) - 1.0001_wp              :137
    double tp0, dtp;   // temperature_planck(1), (2)-(1)               :271-272
    double pi;         // (double)3.14159265359f                       :53
    // correctly rounded reciprocals of the constant divisors (NaN: divisor unsuitable, div_const falls back to __ddiv_rn)
    double r_dlogp, r_dt, r_dtp, r_tp0, r_pi;
    const double *tbase;     // temperature(:,1)                       :131-132
    const float *k32;        // all coefficient tables, fp32 (exact)   element (g,p,t,v) of a gas at
    const double *k64;       //   off + g + ngpt*(p + np*(t + nt*v)), 0-based
    const float *planck32;   // planck_function(ngpt, nplanck)
    const double *planck64;
    const double *rayleigh;  // (ngpt)
    const double *solar;     // (ngpt)
    // walk tables (NULL when the model does not qualify): per g-group of 32,
    //   wbil [group][slot][ip][it][lane]   slot = distinct bilinear table, model order, composite once
    //   wlut [group][iv][ip][it][lane]     the single look-up-table gas
    int nslot, nv_lut;
    const double *wbil64, *wlut64;
    const float *wbil32, *wlut32;
};

// One active gas of this call, already resolved against the model on the host
// (gas_optical_depth, :348-374): caller order, unknown names dropped, composite once.
struct GasRef {
    const void *vmr;
    long long col_stride, lay_stride;  // elements of Real
    long long table_off;               // element offset of coefficient(1,1,1,1)
    double ref;                        // reference_mole_fraction          :146
    double mf0;                        // mole_fraction(1)                 :153
    double log_mf0;                    // log(mole_fraction(1))            :156
    double dlog_mf;                    // log(mole_fraction(2)/mole_fraction(1))  :154-155
    double r_dlog_mf;                  // RN(1 / dlog_mf) for div_const
    double v_clamp;                    // size(mole_fraction) - 1.001_wp   :160
    int code;
    int nv;
};

struct CallGases {
    int n;
    int nbil;               // number of code 0/1/3 gases
    int lut;                // index in g[] of the (single) LUT gas, or -1
    // ---- plan of the walking kernel (host-filled when the call qualifies) ----
    // Bilinear interpolation is linear in the table, so every gas whose weight
    // factor is constant over the call (scalar vmr, or code 0) is pre-mixed into
    // ONE table per call:  sum_g (sw*f_g)*bil(K_g) == sw*bil(sum_g f_g*K_g).
    int nvar;               // bilinear gases whose vmr varies over the call: register slots 1..nvar
    int var_gas[kMaxBil];   //   index in g[]
    int var_off[kMaxBil];   //   element offset of their slab inside one g-group of wbil
    int nmix;               // bilinear gases with a call-constant factor: mixed into slot 0
    int mix_gas[kMaxBil];
    int mix_off[kMaxBil];
    GasRef g[kMaxGases];
};

template <typename Real> struct TableOf;
template <> struct TableOf<double> {
    __device__ static const double *k(const DevModel &m) { return m.k64; }
    __device__ static const double *planck(const DevModel &m) { return m.planck64; }
    __device__ static const double *wbil(const DevModel &m) { return m.wbil64; }
    __device__ static const double *wlut(const DevModel &m) { return m.wlut64; }
};
template <> struct TableOf<float> {
    __device__ static const float *k(const DevModel &m) { return m.k32; }
    __device__ static const float *planck(const DevModel &m) { return m.planck32; }
    __device__ static const float *wbil(const DevModel &m) { return m.wbil32; }
    __device__ static const float *wlut(const DevModel &m) { return m.wlut32; }
};

// x / c for a divisor c that is constant over the call, given rc = RN(1/c):  q0 = RN(x*rc), r = x - c*q0 (exact in one
// FMA), q = RN(q0 + r*rc).  With a correctly rounded reciprocal and q0 within an ulp of x/c this IS the correctly rounded
// quotient (Markstein 1990; the one excluded case, a divisor whose significand is all ones, is detected on the host,
// which then passes rc = NaN) -- bit for bit what __ddiv_rn and the reference's division return, in 3 dependent fp64
// operations instead of ~25 instructions.  Non-finite dividends take q0 as it is (inf / c = inf, NaN stays NaN).
__device__ __forceinline__ double div_const(double x, double c, double rc)
{
    if (rc != rc) return __ddiv_rn(x, c);                  // uniform over the launch
    const double q0 = __dmul_rn(x, rc);
    const double q = __fma_rn(__fma_rn(-c, q0, x), rc, q0);
    return fabs(q0) < __longlong_as_double(0x7ff0000000000000ll) ? q : q0;
}

// ---- per-cell interpolation state (pressure, temperature) -------------------
struct CellIndex {
    int ip0, it0;      // 1-based, as in the reference
    double pw1, tw1;   // pressure_weight1, temperature_weight1
    double sw;         // simple_weight = global_weight*(plev(j+1)-plev(j))   :143
};

__device__ __forceinline__ CellIndex cell_index(const DevModel &m, double p_upper, double p_lower, double tlay)
{
    // p_upper = level_pressure(i,j), p_lower = level_pressure(i,j+1)
    CellIndex c;
    double lp = log(__dmul_rn(0.5, __dadd_rn(p_lower, p_upper)));                     // :120
    double x = div_const(__dsub_rn(lp, m.logp0), m.dlogp, m.r_dlogp);                 // :121
    x = __dadd_rn(1.0, fmax(0.0, fmin(x, m.p_clamp)));                                // :122-125
    c.ip0 = __double2int_rz(x);                                                       // :126
    c.pw1 = __dsub_rn(x, (double)c.ip0);
    double pw0 = __dsub_rn(1.0, c.pw1);
    double t0 = __dadd_rn(__dmul_rn(pw0, m.tbase[c.ip0 - 1]), __dmul_rn(c.pw1, m.tbase[c.ip0]));  // :131-132
    double y = div_const(__dsub_rn(tlay, t0), m.dt, m.r_dt);                          // :133
    y = __dadd_rn(1.0, fmax(0.0, fmin(y, m.t_clamp)));                                // :134-137
    c.it0 = __double2int_rz(y);                                                       // :138
    c.tw1 = __dsub_rn(y, (double)c.it0);
    c.sw = __dmul_rn(m.gw, __dsub_rn(p_lower, p_upper));                              // :143
    return c;
}

// weight of one gas (:144-149; code 0 uses simple_weight, :216)
__device__ __forceinline__ double gas_weight(const GasRef &g, double sw, double vmr)
{
    if (g.code == kRelLinear) return __dmul_rn(sw, __dsub_rn(vmr, g.ref));
    if (g.code == kNone) return sw;
    return __dmul_rn(sw, vmr);
}

// look-up-table index in mole fraction (:153-163); returns 1-based iv0, sets vw1
__device__ __forceinline__ int vmr_index(const GasRef &g, double vmr, double &vw1)
{
    double lv = log(fmax(vmr, g.mf0));                                                // :153
    double z = div_const(__dsub_rn(lv, g.log_mf0), g.dlog_mf, g.r_dlog_mf);           // :156
    z = __dadd_rn(1.0, fmax(0.0, fmin(z, g.v_clamp)));                                // :157-160
    int iv0 = __double2int_rz(z);                                                     // :161
    vw1 = __dsub_rn(z, (double)iv0);
    return iv0;
}

// Planck LUT index (:275-284).  Returns 1-based it0 (0 = the T < T_planck(1)
// branch); a0/a1 are the factors of planck_function(:,max(it0,1)) and (:,+1)
// BEFORE the division by pi: in range (w0, w1); below range (T/t0, 0).
__device__ __forceinline__ int planck_index(const DevModel &m, double T, double &a0, double &a1)
{
    double x = div_const(__dsub_rn(T, m.tp0), m.dtp, m.r_dtp);                        // :275
    if (x >= 0.0) {
        x = __dadd_rn(1.0, x);                                                        // :277
        int it0 = min(__double2int_rz(x), m.nplanck - 1);                             // :278
        a1 = __dsub_rn(x, (double)it0);                                               // :279
        a0 = __dsub_rn(1.0, a1);                                                      // :280
        return it0;
    }
    a0 = div_const(T, m.tp0, m.r_tp0);                                                // :284
    a1 = 0.0;
    return 0;
}

template <typename Real>
__device__ __forceinline__ double load_vmr(const GasRef &g, long long col, long long lay)
{
    return (double)static_cast<const Real *>(g.vmr)[col * g.col_stride + lay * g.lay_stride];
}

// streaming (evict-first) store: outputs are written once and never re-read here,
// keep them from displacing the L2-resident tables
#ifdef ECCKD_PLAIN_STORES
template <typename Real> __device__ __forceinline__ void store_cs(Real *p, Real v) { *p = v; }
#else
template <typename Real> __device__ __forceinline__ void store_cs(Real *p, Real v) { __stcs(p, v); }
#endif

}  // namespace ecckd
