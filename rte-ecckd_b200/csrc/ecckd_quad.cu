// ecckd_quad.cu -- quad_kernel, the product kernel of the ECCKD gas-optics hot path on sm_100a.
//
// One fused pass replaces the loop nests of /root/reference/src/gas_optics_ecckd.f90:
//   calculate_optical_depth :117-240 (once per gas) + accumulate :370
//   calculate_planck_function :273-288 (x3) + shifted copies :423-424
//   calculate_rayleigh_optical_depth :313-317 + ssa/g :456-460 + toa_src :468-472
//
// A persistent block (one per SM) walks tiles of 32 columns top to bottom as ONE continuous stream of
// (tile, layer) positions; four warp roles hand off through mbarriers:
//
//   index warps  (4)  lanes = columns.  Stream position s (a layer of a tile) belongs to warp s mod 4: coalesced loads
//                     of plev / tlay / vmr, the reference's index arithmetic (:120-163) with IEEE divisions in its
//                     operation order, and the interpolation record of each column written as structure-of-arrays
//                     into a ring of R layer slots.
//   compute warps (8) FOUR columns per warp: 8 lanes per column, 4 g-points per lane (g = q, q+8, q+16, q+24).
//                     The record of a layer is read with 5 shared-memory loads that serve four columns at once (one
//                     wavefront each), 16 independent bilinear chains per lane give the fp64 pipe instruction-level
//                     parallelism, and the corner rows of the three table slots (look-up-table gas, per-call mixed
//                     table, one varying gas) live in registers, refilled by predicated 16-byte loads when the
//                     column's (ip0,it0[,iv0]) key moves.  tau (and ssa) go to a 128-byte-swizzled staging tile.
//   store warp   (1)  one thread turns finished staging tiles into cp.async.bulk.tensor stores of the caller's
//                     (ncol, nlay, ngpt) arrays (SW: the g = 0 plane comes from a zeroed tile).
//   Planck warps (3)  longwave: lanes = columns, one level of a tile per unit; interpolate the Planck LUT from
//                     shared memory and store lay_source / lev_source_inc / lev_source_dec / sfc_source directly
//                     (256-byte runs).  lev_source_inc(:,j-1) and lev_source_dec(:,j) are one evaluation.
//
// Against round 1's one-column-per-warp kernel this runs 32 instead of 16 columns per SM in flight at a third of the
// instructions per cell, and a quarter of the record / barrier traffic in the L1 data pipe.
// Registers are traded between the roles with setmaxnreg (compute warps grow, the others shrink).
// The arithmetic per g-point is the same sequence of explicitly rounded operations as in walk_kernel, so both
// kernels produce bit-identical results.
#include "ecckd_device.cuh"
#include "ecckd_kernels.h"
#include "ecckd_pipeline.cuh"
#ifndef ECCKD_STORE_PEND
#define ECCKD_STORE_PEND 0
#endif
#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace ecckd {

namespace {

template <typename Real> struct PairOf;
template <> struct PairOf<double> { using type = double2; };
template <> struct PairOf<float> { using type = float2; };

constexpr int kQCols = kQuadTileCols;   // columns per tile
constexpr int kQNCW = 8;                // compute warps (4 columns each)
constexpr int kQNI = 4;                 // index warps
constexpr int kQNH = 4;                 // helper warpgroup: store warp + 3 Planck (LW) / source (SW) warps
constexpr int kQWarps = kQNCW + kQNI + kQNH;
constexpr int kNegFlag = 0x10000;       // key.x: upside-down layer whose negative-factor gases contribute (:143-149,233-238)

// One layer of the record ring: structure-of-arrays over the 32 columns of the tile, so that the index warps
// (lanes = columns) write and the compute warps (4 columns per warp) read without bank conflicts.
template <typename Real, int NV, int NX = 0> struct alignas(16) QRec {
    using P = typename PairOf<Real>::type;
    P c01[kQCols], c23[kQCols];   // tw0*pw0, tw0*pw1 | tw1*pw0, tw1*pw1
    P lut[kQCols];                // LUT gas: weight (0 when it does not contribute), vmr_weight1
    P w[(NV + NX + 2) / 2][kQCols];   // weights: mixed table, varying gases 1..NV+NX (0 when clamped away)
    P aux[kQCols];                // simple_weight (SW: moles_per_layer; upside-down layers), unused
    struct alignas(sizeof(P)) Key { int2 k; };
    Key key[kQCols];              // x: bilinear corner offset | kNegFlag; y: LUT corner offset (same pitch as the pairs: one
};                                //    lane address serves every field)

template <typename T> __device__ __forceinline__ T make_pair_(double a, double b);
template <> __device__ __forceinline__ double2 make_pair_<double2>(double a, double b) { return make_double2(a, b); }
template <> __device__ __forceinline__ float2 make_pair_<float2>(double a, double b) { return make_float2((float)a, (float)b); }

// predicated loads of four cached rows (corners C00 C10 C01 C11 of one table slot), four g-points per row, written
// straight into the cache registers (tied operands: nothing waits for the data until it is used one layer later).
// 8-byte rows: the lane's g-points {q, q+8} and {q+16, q+24} sit in the two 128-byte halves of the row (walk_pos64),
// so each 16-byte load instruction reads contiguous memory across the 8 lanes of a column.
__device__ __forceinline__ void ld_rows_if(double (&r)[4][4], const double *p, bool pred)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %17, 0;\n\t"
        "@q ld.global.nc.v2.f64 {%0, %1}, [%16];\n\t@q ld.global.nc.v2.f64 {%2, %3}, [%16+128];\n\t"
        "@q ld.global.nc.v2.f64 {%4, %5}, [%16+%18];\n\t@q ld.global.nc.v2.f64 {%6, %7}, [%16+%18+128];\n\t"
        "@q ld.global.nc.v2.f64 {%8, %9}, [%16+%19];\n\t@q ld.global.nc.v2.f64 {%10, %11}, [%16+%19+128];\n\t"
        "@q ld.global.nc.v2.f64 {%12, %13}, [%16+%20];\n\t@q ld.global.nc.v2.f64 {%14, %15}, [%16+%20+128];\n\t}"
        : "+d"(r[0][0]), "+d"(r[0][1]), "+d"(r[0][2]), "+d"(r[0][3]), "+d"(r[1][0]), "+d"(r[1][1]), "+d"(r[1][2]), "+d"(r[1][3]),
          "+d"(r[2][0]), "+d"(r[2][1]), "+d"(r[2][2]), "+d"(r[2][3]), "+d"(r[3][0]), "+d"(r[3][1]), "+d"(r[3][2]), "+d"(r[3][3])
        : "l"(p), "r"((int)pred), "n"(kC10 * 8), "n"(kC01 * 8), "n"(kC11 * 8));
}
__device__ __forceinline__ void ld_rows_if(float (&r)[4][4], const float *p, bool pred)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %17, 0;\n\t"
        "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%16];\n\t"
        "@q ld.global.nc.v4.f32 {%4, %5, %6, %7}, [%16+%18];\n\t"
        "@q ld.global.nc.v4.f32 {%8, %9, %10, %11}, [%16+%19];\n\t"
        "@q ld.global.nc.v4.f32 {%12, %13, %14, %15}, [%16+%20];\n\t}"
        : "+f"(r[0][0]), "+f"(r[0][1]), "+f"(r[0][2]), "+f"(r[0][3]), "+f"(r[1][0]), "+f"(r[1][1]), "+f"(r[1][2]), "+f"(r[1][3]),
          "+f"(r[2][0]), "+f"(r[2][1]), "+f"(r[2][2]), "+f"(r[2][3]), "+f"(r[3][0]), "+f"(r[3][1]), "+f"(r[3][2]), "+f"(r[3][3])
        : "l"(p), "r"((int)pred), "n"(kC10 * 4), "n"(kC01 * 4), "n"(kC11 * 4));
}

template <typename P> __device__ __forceinline__ P lds_pair(unsigned addr);
template <> __device__ __forceinline__ double2 lds_pair<double2>(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
template <> __device__ __forceinline__ float2 lds_pair<float2>(unsigned addr)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ int2 lds_int2(unsigned addr)
{
    int2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
template <typename Real> __device__ __forceinline__ void sts_(unsigned addr, Real v);
template <> __device__ __forceinline__ void sts_<double>(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
template <> __device__ __forceinline__ void sts_<float>(unsigned addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

}  // namespace

// Real   double | float (arithmetic and I/O type; index arithmetic is always double)
// NV     register-cached varying bilinear gases (1 or 3); the call's nvar may be smaller
// NX     further varying gases (0 or 6) whose corner rows are NOT cached: four 16-byte loads per gas and layer, used once
//        (host models that pass co2 / ch4 / n2o / cfc11 / cfc12 as profiles or fields instead of scalars)
// LUT32  cache the look-up-table rows as fp32 (exact: the files are f4) and widen them in the FMA chain
// R      layers in the record ring, NBUF staging tiles, RC / RI / RH registers of compute / index / helper warps
// L      positions (layers) per hand-off: record slots and staging tiles change hands in groups of L, so that the
//        mbarrier round trips, the proxy fence and the warp synchronisation are paid once per L layers
// HIDX   shortwave: the three helper warps (no Planck role there) join the index role -- 7 index warps.  Measured: no
//        change (2.11 vs 2.13 ms on the same box): the index role is not what the shortwave waits for.  Off.
// STAGED  read the KEY of the next record at the top of the iteration and refill each table's cache registers right after the
//         arithmetic has read them for the last time, instead of refilling everything after the arithmetic
// PHELP   longwave: an index warp that finds the record ring full (it is positions ahead of the compute warps) runs Planck
//         units until its slot is free, instead of sleeping: seven warps share the Planck role instead of three
// ZERO_BY_TMA  SW: the g = 0 plane is stored by the TMA unit from a zeroed tile instead of by the source warps
template <typename Real, bool SW, int NV, bool LUT32, int R, int NBUF, int RC, int RI, int RH, bool ZERO_BY_TMA, int L, int NX, bool STAGED, bool HIDX, bool PHELP>
__global__ void __launch_bounds__(kQWarps * 32, 1)
quad_kernel(const __grid_constant__ DevModel m, const __grid_constant__ CallGases cg, const __grid_constant__ KernelArgs<Real> a,
            const Real *__restrict__ mixtab, const __grid_constant__ TmaMaps maps)
{
    using RecT = QRec<Real, NV, NX>;
    constexpr int NW2 = (NV + NX + 2) / 2;                 // weight pairs of a record
    using P = typename PairOf<Real>::type;
    using LutT = typename std::conditional<LUT32, float, Real>::type;
    constexpr int NOUT = SW ? 2 : 1;                       // staged outputs: tau (+ ssa)
    constexpr int NBOX = sizeof(Real) == 8 ? 2 : 1;        // 128-byte rows: 16 fp64 or 32 fp32 columns per TMA box
    constexpr int BOXB = 32 * 128;                         // bytes of one box in shared memory: 32 g-points x 128 bytes
    constexpr int TILEB = NBOX * BOXB;                     // one staged output of one layer
    constexpr int NPW = kQNH - 1;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    static_assert(R % L == 0 && L <= kQNI, "record ring: whole groups; every position of a group comes from a different index warp");
    // Shortwave has no Planck role: with the g = 0 plane leaving through the TMA unit its three helper warps have nothing
    // to do and join the index role (7 index warps: the index arithmetic is a chain of ~140 dependent fp64 operations
    // per position and warp, 9.6 cycles each).  toa_src is written by the index warp that handles a tile's first layer.
    constexpr bool HELPERS_INDEX = HIDX;
    static_assert(!HIDX || (SW && ZERO_BY_TMA), "only the shortwave's helper warps are free, and only when the TMA unit writes the g = 0 plane");
    constexpr int NIW = HELPERS_INDEX ? kQNI + kQNH - 1 : kQNI;      // index warps
    static_assert(!HELPERS_INDEX || RH >= RI, "helper warps that run the index role need its registers");
    unsigned char *stage = smem_raw;                                           // [NBUF][L][NOUT][TILEB]  (1024-byte aligned)
    unsigned char *zero_tile = stage + (size_t)NBUF * L * NOUT * TILEB;        // SW: source of the g = 0 plane
    RecT *recs = reinterpret_cast<RecT *>(zero_tile + (SW ? BOXB : 0));        // [R]
    unsigned long long *tile_full = reinterpret_cast<unsigned long long *>(recs + R);
    unsigned long long *tile_empty = tile_full + NBUF;
    unsigned long long *rec_full = tile_empty + NBUF;
    unsigned long long *rec_empty = rec_full + R / L;
    // LW: Planck LUT of this g-group, [g][row], 8 bytes per entry.  fp64 kernels hold the rows widened to fp64 (exact) and
    // read rows r and r + 1 with two loads; fp32 kernels hold (row, row + 1) pairs and read both with one load.  Either
    // way there is no conversion in the Planck warps' loop (widening there, on the conversion unit or with integer
    // operations, was two thirds of their instructions)
    using PlT = typename std::conditional<sizeof(Real) == 8, double, float2>::type;
    PlT *spl = reinterpret_cast<PlT *>(rec_empty + R / L);
    const int pl_pitch = m.nplanck + 1;                        // one zero entry behind the last row

    // logical warp order [compute | index | store | helpers]; hardware warp ids run the other way round because the
    // issue arbiter prefers the highest warp id: the compute warps get priority over the throughput-oriented helpers
    const int warp = kQWarps - 1 - (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int ncol = a.ncol, nlay = a.nlay, ngpt = m.ngpt;
    const int grp = blockIdx.y, g0 = grp * kGP;
    const int ntile_all = (ncol + kQCols - 1) / kQCols;
    const int nslot = (ntile_all - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this block: b, b + grid, ...
    const int S = nslot * nlay;                                                             // stream positions of this block
    const int Spad = (S + L - 1) / L * L;                                                   // ... in whole hand-off groups
    const size_t pitch = (size_t)ncol, plane = (size_t)ncol * nlay;
    const bool has_lut = cg.lut >= 0, has_mix = cg.nmix > 0;
    const int nvar = cg.nvar;

    if (threadIdx.x == 0) {
        for (int b = 0; b < NBUF; ++b) { mbar_init(&tile_full[b], kQNCW); mbar_init(&tile_empty[b], 1); }
        for (int q = 0; q < R / L; ++q) { mbar_init(&rec_full[q], L); mbar_init(&rec_empty[q], kQNCW); }
        if (!SW) *reinterpret_cast<unsigned *>(spl + (size_t)(m.nplanck + 1) * kGP) = 0u;       // unit_ctr
    }
    if (!SW) {
        const int np_ = m.nplanck;
        for (int q = threadIdx.x; q < pl_pitch * kGP; q += blockDim.x) {
            const int row = q / kGP, gl = q % kGP;
            const bool gok = g0 + gl < ngpt;
            const float p0 = (gok && row < np_) ? m.planck32[(size_t)row * ngpt + g0 + gl] : 0.0f;
            const float p1 = (gok && row + 1 < np_) ? m.planck32[(size_t)(row + 1) * ngpt + g0 + gl] : 0.0f;
            if constexpr (sizeof(Real) == 8) spl[gl * pl_pitch + row] = (double)p0;
            else spl[gl * pl_pitch + row] = make_float2(p0, p1);
        }
    } else {
        for (int q = threadIdx.x; q < BOXB / 4; q += blockDim.x) reinterpret_cast<float *>(zero_tile)[q] = 0.0f;
        fence_proxy_async_smem();
    }
    __syncthreads();

    // ---- Planck role (LW): lanes = columns, a unit is one level of one tile.  Units are claimed from a counter in shared
    //      memory by the Planck warps and (PHELP) by index warps that have nothing else to do ----
    unsigned *unit_ctr = reinterpret_cast<unsigned *>(spl + (size_t)pl_pitch * kGP);
    const int ng = min(kGP, ngpt - g0);
    const bool has_lev = !SW && a.tlev != nullptr;
    const int nlevu = has_lev ? nlay + 1 : nlay;          // levels 0..nlay carry a value when tlev is present
    const int nunit = SW ? 0 : nslot * nlevu;
    const size_t plane_b = plane * sizeof(Real);
    const int S_ = pl_pitch;
    auto rows2 = [](const PlT *p, Real &x, Real &y) {          // table rows r and r + 1 of one g-point
        if constexpr (sizeof(Real) == 8) { x = p[0]; y = p[1]; }
        else { const float2 v = *p; x = v.x; y = v.y; }
    };
    auto place = [&](int u, int &j, long long &colp, bool &c_ok) {
        const int ts = u / nlevu;
        j = u - ts * nlevu;
        const long long cc = (long long)((int)blockIdx.x + ts * (int)gridDim.x) * kQCols + lane;
        c_ok = cc < ncol;
        colp = min(cc, (long long)ncol - 1);
    };
    auto planck_fetch = [&](int u, double &tl, double &tv) {
        int j; long long colp; bool c_ok;
        place(u, j, colp, c_ok);
        tl = (double)a.tlay[colp + (long long)ncol * min(j, nlay - 1)];
        tv = 0.0;
        if (has_lev) tv = (double)a.tlev[colp + (long long)ncol * min(j, nlay)];
    };
    int my_unit = warp - (kQNCW + kQNI) - 1;              // !PHELP: the Planck warps take every third unit, no counter
    auto claim_unit = [&]() -> int {
        if constexpr (PHELP) {
            unsigned u = 0;
            if (lane == 0) u = atomicAdd(unit_ctr, 1u);
            return (int)min(__shfl_sync(0xffffffffu, u, 0), (unsigned)nunit);
        } else {
            const int u = min(my_unit, nunit);
            my_unit += NPW;
            return u;
        }
    };
    auto planck_unit = [&](int u, double tl, double tv) {
        int j; long long colp; bool c_ok;
        place(u, j, colp, c_ok);
        const bool ok_lay = c_ok && j < nlay;                       // lay_source(:,j), lev_source_dec(:,j)
        const bool ok_inc = c_ok && has_lev && j >= 1 && j <= nlay; // lev_source_inc(:,j-1)
        double a0, a1;
        int it = planck_index(m, tl, a0, a1);
        const PlT *tl_p = spl + (max(it, 1) - 1);
        const Real l0 = (Real)div_const(a0, m.pi, m.r_pi), l1 = (Real)div_const(a1, m.pi, m.r_pi);
        const size_t off = (size_t)colp + pitch * (size_t)j + plane * (size_t)g0;
        char *p_lay = reinterpret_cast<char *>(a.out[1] + off);
        // g-points in batches of GB: all table reads of a batch are issued before its arithmetic and its
        // stores, so that a batch is GB independent chains instead of one serial LDS -> F2F -> FMA -> STG chain
        constexpr int GB = 4;
        // (rows of g-points beyond ngpt exist in the shared table and hold zeros: reads need no clamp; a whole
        //  g-group -- the usual case -- needs no per-g-point store predicate either)
        if (has_lev) {
            it = planck_index(m, tv, a0, a1);
            const PlT *tv_p = spl + (max(it, 1) - 1);
            const Real v0 = (Real)div_const(a0, m.pi, m.r_pi), v1 = (Real)div_const(a1, m.pi, m.r_pi);
            char *p_dec = reinterpret_cast<char *>(a.out[3] + off);
            char *p_inc = reinterpret_cast<char *>(a.out[2] + off) - pitch * sizeof(Real);   // only dereferenced when j >= 1
            auto gloop = [&](auto full) {
                constexpr bool FULL = decltype(full)::value;
                const int ngl = FULL ? kGP : ng;
                const PlT *tlp = tl_p, *tvp = tv_p;
                for (int gq = 0; gq < ngl; gq += GB, tlp += GB * S_, tvp += GB * S_) {
                    Real pl0[GB], pl1[GB], pv0[GB], pv1[GB];
#pragma unroll
                    for (int k = 0; k < GB; ++k) { rows2(tlp + k * S_, pl0[k], pl1[k]); rows2(tvp + k * S_, pv0[k], pv1[k]); }
                    Real vl[GB], vv[GB];
#pragma unroll
                    for (int k = 0; k < GB; ++k) {
                        vl[k] = fma_(l1, pl1[k], mul_(l0, pl0[k]));
                        vv[k] = fma_(v1, pv1[k], mul_(v0, pv0[k]));      // B(tlev(:,j))
                    }
#pragma unroll
                    for (int k = 0; k < GB; ++k) {
                        const bool gk = FULL || gq + k < ng;
                        stcs_if(reinterpret_cast<Real *>(p_lay), vl[k], ok_lay && gk);
                        stcs_if(reinterpret_cast<Real *>(p_dec), vv[k], ok_lay && gk);            // lev_source_dec(:,j)   :424
                        stcs_if(reinterpret_cast<Real *>(p_inc), vv[k], ok_inc && gk);            // lev_source_inc(:,j-1) :423
                        p_lay += plane_b; p_dec += plane_b; p_inc += plane_b;
                    }
                }
            };
            if (ng == kGP) gloop(std::true_type{}); else gloop(std::false_type{});
        } else {
            auto gloop = [&](auto full) {
                constexpr bool FULL = decltype(full)::value;
                const int ngl = FULL ? kGP : ng;
                const PlT *tlp = tl_p;
                for (int gq = 0; gq < ngl; gq += GB, tlp += GB * S_) {
                    Real pl0[GB], pl1[GB];
#pragma unroll
                    for (int k = 0; k < GB; ++k) rows2(tlp + k * S_, pl0[k], pl1[k]);
#pragma unroll
                    for (int k = 0; k < GB; ++k) {
                        const Real vl = fma_(l1, pl1[k], mul_(l0, pl0[k]));
                        stcs_if(reinterpret_cast<Real *>(p_lay), vl, ok_lay && (FULL || gq + k < ng));
                        p_lay += plane_b;
                    }
                }
            };
            if (ng == kGP) gloop(std::true_type{}); else gloop(std::false_type{});
        }
    };
    int iw = -1;                                               // index-role id of this warp (-1: another role)
    // ================================ helper warpgroup ================================
    if (warp >= kQNCW + kQNI) {
        reg_dec<RH>();
        const int hw = warp - (kQNCW + kQNI);
        if (HELPERS_INDEX && hw >= 1) iw = kQNI + hw - 1;
        else {
        if (ECCKD_DBG(128) && hw == 0) return;   // experiment: Planck warps only
        if (hw == 0) {
            // ---- store warp: staging tiles -> cp.async.bulk.tensor stores ----
            if (lane == 0) {
                // ECCKD_STORE_PEND bulk groups stay in flight: the tile of group n - pend is handed back once group n has been
                // issued, so that the TMA unit has the next store queued while it reads a tile out (measured with 1 and 2:
                // no change, the default waits for each group)
#ifdef ECCKD_TUNE
                const int pend = ECCKD_DBG(2048) ? 1 : ECCKD_DBG(4096) ? 2 : 0;
#else
                constexpr int pend = ECCKD_STORE_PEND;
#endif
                int b = 0, par = 0, ts = 0, j = 0, rb = 0;
                for (int s0 = 0; s0 < S; s0 += L) {
                    mbar_wait_relaxed(&tile_full[b], par, kStoreNap);
#pragma unroll
                    for (int l = 0; l < L; ++l) {
                        if (s0 + l >= S) break;                          // padding positions of the last group
                        const int col0 = ((int)blockIdx.x + ts * (int)gridDim.x) * kQCols;
                        const unsigned char *sb = stage + (size_t)(b * L + l) * NOUT * TILEB;
#pragma unroll
                        for (int o = 0; o < NOUT; ++o)
                            if (a.out[o] && !ECCKD_DBG(1)) {
#pragma unroll
                                for (int h = 0; h < NBOX; ++h)
                                    if (h == 0 || col0 + 16 < ncol) tma_store_3d(&maps.m[o], sb + o * TILEB + h * BOXB, col0 + 16 * h, j, g0);
                            }
                        if (SW && ZERO_BY_TMA && a.out[2] && !ECCKD_DBG(1)) {
#pragma unroll
                            for (int h = 0; h < NBOX; ++h)
                                if (h == 0 || col0 + 16 < ncol) tma_store_3d(&maps.m[2], zero_tile, col0 + 16 * h, j, g0);   // g = 0  :460
                        }
                        if (++j == nlay) { j = 0; ++ts; }
                    }
                    tma_commit();
                    // the tiles have been read out of shared memory: hand them back
                    if (pend == 0) tma_wait_read<0>(); else if (pend == 1) tma_wait_read<1>(); else tma_wait_read<2>();
                    if (s0 >= pend * L) { mbar_arrive(&tile_empty[rb]); if (++rb == NBUF) rb = 0; }
                    if (++b == NBUF) { b = 0; par ^= 1; }
                }
                tma_wait_all<0>();                       // (nobody waits for the last tiles)
            }
            return;
        }
        const int pw = hw - 1;
        if (SW) {
            // ---- source warps (SW): lanes = columns.  toa_src(:, g) = solar_irradiance(g) :468-472 and the g = 0
            //      plane :460, written with plain 256-byte row stores: it does not have to pass through the TMA unit ----
            for (int ts = 0; ts < nslot; ++ts) {
                const long long cc = (long long)((int)blockIdx.x + ts * (int)gridDim.x) * kQCols + lane;
                if (cc >= ncol) continue;
                if (a.out2d && ts % NPW == pw)
                    for (int gq = 0; gq < ng; ++gq) a.out2d[cc + (long long)ncol * (g0 + gq)] = (Real)m.solar[g0 + gq];
                if (a.out[2] && !ZERO_BY_TMA) {
                    for (int j = pw; j < nlay; j += NPW) {
                        char *pz = reinterpret_cast<char *>(a.out[2] + ((size_t)cc + pitch * (size_t)j + plane * (size_t)g0));
#pragma unroll 4
                        for (int gq = 0; gq < ng; ++gq, pz += plane * sizeof(Real)) store_cs(reinterpret_cast<Real *>(pz), (Real)0);
                    }
                }
            }
            return;
        }
        // ---- Planck warps (LW) ----
        {
            if (ECCKD_DBG(64)) return;
            // software pipeline: the next unit is claimed and its temperatures fetched before the g-loop of the current one
            double tl_n = 0.0, tv_n = 0.0;
            int u = claim_unit();
            if (u < nunit) planck_fetch(u, tl_n, tv_n);
            while (u < nunit) {
                const int un = claim_unit();
                const double tl = tl_n, tv = tv_n;
                if (un < nunit) planck_fetch(un, tl_n, tv_n);
                planck_unit(u, tl, tv);
                u = un;
            }
            if (a.out2d) {                                 // sfc_source(:, g) = B(tsfc)   :408-412
                for (int ts = pw; ts < nslot; ts += NPW) {
                    const long long cc = (long long)((int)blockIdx.x + ts * (int)gridDim.x) * kQCols + lane;
                    const long long colp = min(cc, (long long)ncol - 1);
                    double a0, a1;
                    const int it = planck_index(m, (double)a.tsfc[colp], a0, a1);
                    const int r = max(it, 1) - 1;
                    const Real s0 = (Real)div_const(a0, m.pi, m.r_pi), s1 = (Real)div_const(a1, m.pi, m.r_pi);
                    if (cc < ncol)
                        for (int gq = 0; gq < ng; ++gq) {
                            Real ps0, ps1;
                            rows2(spl + gq * S_ + r, ps0, ps1);
                            a.out2d[colp + (long long)ncol * (g0 + gq)] = fma_(s1, ps1, mul_(s0, ps0));
                        }
                }
            }
        }
        return;
        }
    }

    // ================================ index warps =====================================
    if (warp >= kQNCW && warp < kQNCW + kQNI) {
        reg_dec<RI>();
        iw = warp - kQNCW;
    }
    if (iw >= 0) {
        if (ECCKD_DBG(128)) return;
        int ts = 0, j = iw;                                    // position s = ts * nlay + j
        while (j >= nlay) { j -= nlay; ++ts; }
        int slot = iw % R, n = iw / R;                          // ring slot of s and how often it has been used before
        // The inputs of a position stream in from DRAM (latency of microseconds under the kernel's write load): they are
        // fetched one position ahead of the arithmetic, like the Planck warps do.
        struct In { double pu, pl, tl, vl, vv[NV + NX]; };
        auto fetch = [&](int ts_, int j_, In &in) {
            const long long col = min((long long)((int)blockIdx.x + ts_ * (int)gridDim.x) * kQCols + lane, (long long)ncol - 1);
            // (tail tiles recompute the last column; never stored)
            in.pu = (double)a.plev[col + (long long)ncol * j_];
            in.pl = (double)a.plev[col + (long long)ncol * (j_ + 1)];
            in.tl = (double)a.tlay[col + (long long)ncol * j_];
            in.vl = has_lut ? load_vmr<Real>(cg.g[cg.lut], col, j_) : 0.0;
#pragma unroll
            for (int v = 0; v < NV + NX; ++v) in.vv[v] = v < nvar ? load_vmr<Real>(cg.g[cg.var_gas[v]], col, j_) : 0.0;
        };
        // ... and pulled into L2 kPF positions of this warp earlier still (no registers: prefetch.global.L2), so that the
        // register fetch one position ahead finds them a few hundred cycles away instead of a DRAM round trip
        constexpr int kPF = 4;
        auto prefetch = [&](int ts_, int j_) {
            const long long col = min((long long)((int)blockIdx.x + ts_ * (int)gridDim.x) * kQCols + lane, (long long)ncol - 1);
            auto pf = [](const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); };
            pf(a.plev + col + (long long)ncol * (j_ + 1));
            pf(a.tlay + col + (long long)ncol * j_);
            if (has_lut) { const GasRef &gr = cg.g[cg.lut]; pf(static_cast<const Real *>(gr.vmr) + col * gr.col_stride + (long long)j_ * gr.lay_stride); }
#pragma unroll
            for (int v = 0; v < NV + NX; ++v)
                if (v < nvar) { const GasRef &gr = cg.g[cg.var_gas[v]]; pf(static_cast<const Real *>(gr.vmr) + col * gr.col_stride + (long long)j_ * gr.lay_stride); }
        };
        int pj = j, pts = ts;                                   // position kPF * NIW ahead of s
        for (int q2 = 0; q2 < kPF; ++q2) { pj += NIW; while (pj >= nlay) { pj -= nlay; ++pts; } }
        In nxt;
        bool units_left = PHELP && !SW;
        if (iw < S) fetch(ts, j, nxt);
        for (int s = iw; s < Spad; s += NIW) {
            const In in = nxt;
            {
                int j2 = j + NIW, ts2 = ts;
                while (j2 >= nlay) { j2 -= nlay; ++ts2; }
                if (s + NIW < S) fetch(ts2, j2, nxt);
                if (!ECCKD_DBG(512) && s + kPF * NIW < S) prefetch(pts, pj);
                pj += NIW; while (pj >= nlay) { pj -= nlay; ++pts; }
            }
            P c01, c23, lutw, wv[NW2], aux;
            int2 key;
            if ((ECCKD_DBG(2) && s >= R) || s >= S) {   // (s >= S: padding position of the last hand-off group, never stored)
                c01 = c23 = lutw = aux = make_pair_<P>(0.25, 0.25); key = make_int2(0, 0);
#pragma unroll
                for (int v = 0; v < NW2; ++v) wv[v] = c01;
            } else {
                const CellIndex ci = cell_index(m, in.pu, in.pl, in.tl);
                const double pw0 = __dsub_rn(1.0, ci.pw1), tw0 = __dsub_rn(1.0, ci.tw1);
                c01 = make_pair_<P>(__dmul_rn(tw0, pw0), __dmul_rn(tw0, ci.pw1));
                c23 = make_pair_<P>(__dmul_rn(ci.tw1, pw0), __dmul_rn(ci.tw1, ci.pw1));
                const int ip = ci.ip0 - 1, it = ci.it0 - 1;
                // Tables are >= 0 (checked at model create), so od < 0 <=> weight < 0: the per-g clamp (:233-238)
                // becomes "the gas contributes nothing when its weight is negative" -- its weight is stored as 0,
                // which leaves tau unchanged bit for bit (fma(0, finite, tau) == tau); NaN weights propagate.
                const bool up = ci.sw < 0.0;
                double W[NV + NX + 1];
                W[0] = (has_mix && !up) ? ci.sw : 0.0;
#pragma unroll
                for (int v = 0; v < NV + NX; ++v) {
                    double wd = 0.0;
                    if (v < nvar) {
                        wd = gas_weight(cg.g[cg.var_gas[v]], ci.sw, in.vv[v]);
                        if (wd < 0.0) wd = 0.0;
                    }
                    W[v + 1] = wd;
                }
#pragma unroll
                for (int v = 0; v < NW2; ++v) wv[v] = make_pair_<P>(W[2 * v], 2 * v + 1 <= NV + NX ? W[2 * v + 1] : 0.0);
                aux = make_pair_<P>(ci.sw, 0.0);
                key.x = (ip * kNTP + it) * kGP | ((has_mix && up) ? kNegFlag : 0);
                key.y = 0;
                lutw = make_pair_<P>(0.0, 0.0);
                if (has_lut) {
                    const GasRef &gr = cg.g[cg.lut];
                    double v1;
                    const int iv = vmr_index(gr, in.vl, v1) - 1;
                    key.y = ((iv * kNPP + ip) * kNTP + it) * kGP;
                    double wd = gas_weight(gr, ci.sw, in.vl);
                    if (wd < 0.0) wd = 0.0;
                    lutw = make_pair_<P>(wd, v1);
                }
            }
            if (HELPERS_INDEX && j == 0 && s < S && a.out2d) {        // toa_src(:, g) = solar_irradiance(g)   :468-472
                const long long cc = (long long)((int)blockIdx.x + ts * (int)gridDim.x) * kQCols + lane;
                if (cc < ncol)
                    for (int gq = 0; gq < min(kGP, ngpt - g0); ++gq) a.out2d[cc + (long long)ncol * (g0 + gq)] = (Real)m.solar[g0 + gq];
            }
            if (n >= 1) {
                if constexpr (PHELP) {         // ring full: this warp is R positions ahead -- time for a Planck unit
                    while (!mbar_test(&rec_empty[slot / L], (n - 1) & 1)) {
                        const int u = units_left ? claim_unit() : nunit;
                        if (u >= nunit) { units_left = false; mbar_wait_relaxed(&rec_empty[slot / L], (n - 1) & 1, kIndexNap); break; }
                        double tl, tv;
                        planck_fetch(u, tl, tv);
                        planck_unit(u, tl, tv);
                    }
                } else mbar_wait_relaxed(&rec_empty[slot / L], (n - 1) & 1, kIndexNap);
            }
            RecT &rc = recs[slot];
            rc.c01[lane] = c01; rc.c23[lane] = c23; rc.lut[lane] = lutw; rc.aux[lane] = aux; rc.key[lane].k = key;
#pragma unroll
            for (int v = 0; v < NW2; ++v) rc.w[v][lane] = wv[v];
            __syncwarp();
            if (lane == 0) mbar_arrive(&rec_full[slot / L]);
            j += NIW; while (j >= nlay) { j -= nlay; ++ts; }
            slot += NIW; while (slot >= R) { slot -= R; ++n; }
        }
        if constexpr (PHELP) {
            while (units_left) {
                const int u = claim_unit();
                if (u >= nunit) break;
                double tl, tv;
                planck_fetch(u, tl, tv);
                planck_unit(u, tl, tv);
            }
        }
        return;
    }

    // ================================ compute warps ===================================
    reg_inc<RC>();
    if (ECCKD_DBG(128)) return;
    const int cw = lane >> 3, q = lane & 7;            // column of the warp's quad, lane of the column
    const int ct = 4 * warp + cw;                      // column of the tile
    const int ngroups = gridDim.y;
    // table pointers of this lane (rows are stored permuted, see walk_pos32 / walk_pos64): 4-byte rows hold the lane's
    // g-points q, q+8, q+16, q+24 as one 16-byte vector at 4q; 8-byte rows as two vectors at 2q and 16 + 2q
    constexpr int LQ = sizeof(LutT) == 8 ? 2 : 4, MQ = sizeof(Real) == 8 ? 2 : 4;
    // (caching the varying gases' rows at full width from the fp64 tables, or widening them on the integer pipe, instead
    //  of F2F in the FMA chain: both measured, neither faster)
    const float *wbil = m.wbil32 + (size_t)grp * m.nslot * kSlab + 4 * q;
    const LutT *wlut = (LUT32 ? reinterpret_cast<const LutT *>(m.wlut32) : reinterpret_cast<const LutT *>(TableOf<Real>::wlut(m))) +
                       (size_t)grp * m.nv_lut * kSlab + LQ * q;
    const Real *wmix = mixtab + (size_t)grp * kSlab + MQ * q;    // negative-sign copy at + ngroups*kSlab
    const float *wvar[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) wvar[v] = wbil + (v < nvar ? cg.var_off[v] : 0);

    Real km[4][4];            // mixed table      [corner][g]
    float kb[NV][4][4];       // varying gases (fp32 rows, exact; widened in the FMA chain)
    LutT kh[2][4][4];         // look-up-table gas, mole fractions iv0 and iv0 + 1
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            km[r][i] = 0; kh[0][r][i] = 0; kh[1][r][i] = 0;
#pragma unroll
            for (int v = 0; v < NV; ++v) kb[v][r][i] = 0;
        }
    Real sigma[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) sigma[i] = SW ? (Real)m.rayleigh[min(g0 + q + 8 * i, ngpt - 1)] : (Real)0;

    // (re)load the register cache for corner offsets (pkoff, hkoff); every load is predicated on "this column's key
    // moved", so the four columns of the warp refill independently without divergence
    bool first_fill = true;            // (ECCKD_DBG 8192: after the first fill the loads are issued with false predicates)
    auto refill_lut = [&](int hkoff, bool moved) {
        if (ECCKD_DBG(8192) && !first_fill) moved = false;
        const bool pl = has_lut && moved;
        if (!ECCKD_DBG(8)) {
            ld_rows_if(kh[0], wlut + hkoff, pl);
            ld_rows_if(kh[1], wlut + hkoff + kSlab, pl);
        }
    };
    auto refill_mix = [&](int pkoff, bool moved) {
        if (ECCKD_DBG(8192) && !first_fill) moved = false;
        if (!ECCKD_DBG(8)) ld_rows_if(km, wmix + pkoff, moved && has_mix);
    };
    auto refill_var = [&](int pkoff, bool moved) {
        if (ECCKD_DBG(8192) && !first_fill) moved = false;
        if (!ECCKD_DBG(8)) {
#pragma unroll
            for (int v = 0; v < NV; ++v) ld_rows_if(kb[v], wvar[v] + pkoff, moved && v < nvar);
        }
    };
    auto refill_bil = [&](int pkoff, bool moved) { refill_mix(pkoff, moved); refill_var(pkoff, moved); };

    // staging address of this lane's first value (g-point q): row = g, 16-byte chunks XOR-swizzled with (row & 7) = q
    // exactly as CU_TENSOR_MAP_SWIZZLE_128B lays the box out; value i (g-point q + 8 i) is 8 rows = 1024 bytes further
    unsigned st_off;
    if (sizeof(Real) == 8) {
        const int col16 = 4 * (warp & 3) + cw;
        st_off = (warp >> 2) * BOXB + q * 128 + (((col16 >> 1) ^ q) << 4) + (col16 & 1) * 8;
    } else {
        st_off = q * 128 + ((warp ^ q) << 4) + cw * 4;
    }
    unsigned stage_a = smem_u32(stage) + st_off;
    // the four barrier arrays are contiguous: one base register, the array is chosen by an immediate offset
    unsigned a_bar = smem_u32(tile_full);
    constexpr int B_TFULL = 0, B_TEMPTY = NBUF * 8, B_RFULL = 2 * NBUF * 8, B_REMPTY = (2 * NBUF + R / L) * 8;
    // record fields of this lane's column, as shared-window addresses relative to the ring slot
    constexpr unsigned PB = sizeof(P), RB = sizeof(RecT);
    constexpr unsigned O_C01 = 0, O_C23 = kQCols * PB, O_LUT = 2 * kQCols * PB, O_W = 3 * kQCols * PB,
                       O_AUX = (3 + NW2) * kQCols * PB, O_KEY = (4 + NW2) * kQCols * PB;
    static_assert(O_KEY + kQCols * PB == RB, "record layout");
    unsigned rec_a = smem_u32(recs) + ct * PB;
    // loop invariants are pinned in registers: otherwise they are rematerialised from kernel parameters and special
    // registers in every iteration (measured in round 1, and visible in the SASS)
    keep(stage_a); keep(a_bar); keep(rec_a);
    keep(wmix); keep(wlut);
#pragma unroll
    for (int v = 0; v < NV; ++v) keep(wvar[v]);

    if (S > 0) {
        // The record of position s + 1 is read into registers one position ahead, after the arithmetic of s (reading all
        // of it ahead of the arithmetic costs 20 live registers: spills, measured slower).  STAGED reads only its KEY at the
        // top of iteration s and refills each table's cache registers as soon as the arithmetic of s has read them for the
        // last time (look-up table first, then the mixed table, then the varying gases) instead of all at the end: the
        // loads get the rest of the arithmetic as a head start.
        struct Cur { P c01, c23, lw, wv[NW2]; Real swt; int2 key; };
        auto load_key = [&](int slot_, Cur &r) { r.key = lds_int2(rec_a + slot_ * RB + O_KEY); };
        auto load_rec = [&](int slot_, Cur &r, bool with_key) {
            const unsigned ra = rec_a + slot_ * RB;
            if (with_key) r.key = lds_int2(ra + O_KEY);
            r.c01 = lds_pair<P>(ra + O_C01); r.c23 = lds_pair<P>(ra + O_C23); r.lw = lds_pair<P>(ra + O_LUT);
#pragma unroll
            for (int v = 0; v < NW2; ++v) r.wv[v] = lds_pair<P>(ra + O_W + v * kQCols * PB);
            r.swt = SW ? lds_pair<P>(ra + O_AUX).x : (Real)0;
        };
        Cur cur, nxt;
        mbar_wait_a<B_RFULL>(a_bar, 0);
        load_rec(0, cur, true);
        refill_bil(cur.key.x & 0xffff, true);
        refill_lut(cur.key.y, true);
        first_fill = false;
        nxt = cur;
        int slot = 0, rpar = 0, b = 0, tpar = 1;       // tpar: parity to wait for on tile_empty (the phase before the
                                                       // initial one counts as complete, so the first pass never blocks)
        int lane0 = lane == 0;
        keep(lane0);
        int ls = 0;                                    // position inside the hand-off group
        for (int s = 0; s < Spad; ++s) {
            const int nslot_r = slot + 1 == R ? 0 : slot + 1;
            const bool more = s + 1 < Spad;
            if (STAGED && more) {
                if (ls == L - 1) mbar_wait_a<B_RFULL>(a_bar + (nslot_r / L) * 8, nslot_r == 0 ? rpar ^ 1 : rpar);   // next group
                load_key(nslot_r, nxt);
            }
            const Real c0 = cur.c01.x, c1 = cur.c01.y, c2 = cur.c23.x, c3 = cur.c23.y;
            const Real omv = add_((Real)1, -cur.lw.y);
            Real tau[4];
            if (ECCKD_DBG(4)) { tau[0] = c0; tau[1] = c1; tau[2] = c2; tau[3] = c3; }
            else if constexpr (STAGED) {       // the same operations per g-point, table by table
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const Real b0 = bil4(c0, c1, c2, c3, kh[0][0][i], kh[0][1][i], kh[0][2][i], kh[0][3][i]);
                    const Real b1 = bil4(c0, c1, c2, c3, kh[1][0][i], kh[1][1][i], kh[1][2][i], kh[1][3][i]);
                    tau[i] = mul_(cur.lw.x, fma_(cur.lw.y, b1, mul_(omv, b0)));
                }
                refill_lut(nxt.key.y, more && nxt.key.y != cur.key.y);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    tau[i] = fma_(cur.wv[0].x, bil4(c0, c1, c2, c3, km[0][i], km[1][i], km[2][i], km[3][i]), tau[i]);
                refill_mix(nxt.key.x & 0xffff, more && nxt.key.x != cur.key.x);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const Real wgt = ((v + 1) & 1) ? cur.wv[(v + 1) >> 1].y : cur.wv[(v + 1) >> 1].x;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        tau[i] = fma_(wgt, bil4(c0, c1, c2, c3, kb[v][0][i], kb[v][1][i], kb[v][2][i], kb[v][3][i]), tau[i]);
                }
                refill_var(nxt.key.x & 0xffff, more && nxt.key.x != cur.key.x);
            } else
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const Real b0 = bil4(c0, c1, c2, c3, kh[0][0][i], kh[0][1][i], kh[0][2][i], kh[0][3][i]);
                const Real b1 = bil4(c0, c1, c2, c3, kh[1][0][i], kh[1][1][i], kh[1][2][i], kh[1][3][i]);
                const Real bm = bil4(c0, c1, c2, c3, km[0][i], km[1][i], km[2][i], km[3][i]);
                // (1-v1)*b0 + v1*b1, not b0 + v1*(b1-b0): that cancels when b1 << b0
                Real t = mul_(cur.lw.x, fma_(cur.lw.y, b1, mul_(omv, b0)));
                t = fma_(cur.wv[0].x, bm, t);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const Real bv = bil4(c0, c1, c2, c3, kb[v][0][i], kb[v][1][i], kb[v][2][i], kb[v][3][i]);
                    const Real wgt = ((v + 1) & 1) ? cur.wv[(v + 1) >> 1].y : cur.wv[(v + 1) >> 1].x;
                    t = fma_(wgt, bv, t);
                }
                tau[i] = t;
            }
            if constexpr (NX > 0) {
                // gases beyond the cached ones: corner rows straight from L1/L2, two gases in flight at a time
                const float *xb = wbil + (cur.key.x & 0xffff);
                auto rows_of = [&](int x, float4 (&r)[4]) {
                    const bool on = NV + x < nvar;
                    const float4 *p = reinterpret_cast<const float4 *>(xb + (on ? cg.var_off[NV + x] : 0));
                    r[0] = __ldg(p + kC00 / 4); r[1] = __ldg(p + kC10 / 4); r[2] = __ldg(p + kC01 / 4); r[3] = __ldg(p + kC11 / 4);
                };
                auto acc = [&](int x, const float4 (&r)[4]) {
                    const int wi = NV + x + 1;
                    const Real wgt = (wi & 1) ? cur.wv[wi >> 1].y : cur.wv[wi >> 1].x;      // 0 for absent / clamped gases
                    tau[0] = fma_(wgt, bil4(c0, c1, c2, c3, r[0].x, r[1].x, r[2].x, r[3].x), tau[0]);
                    tau[1] = fma_(wgt, bil4(c0, c1, c2, c3, r[0].y, r[1].y, r[2].y, r[3].y), tau[1]);
                    tau[2] = fma_(wgt, bil4(c0, c1, c2, c3, r[0].z, r[1].z, r[2].z, r[3].z), tau[2]);
                    tau[3] = fma_(wgt, bil4(c0, c1, c2, c3, r[0].w, r[1].w, r[2].w, r[3].w), tau[3]);
                };
                float4 ra_[4], rb_[4];
                rows_of(0, ra_);
#pragma unroll
                for (int x = 0; x < NX; x += 2) {
                    rows_of(x + 1, rb_);
                    acc(x, ra_);
                    if (x + 2 < NX) rows_of(x + 2, ra_);
                    acc(x + 1, rb_);
                }
            }
            if (cur.key.x & kNegFlag) {   // upside-down layer: the negative-factor gases give positive optical depth
                const Real *k = wmix + (size_t)ngroups * kSlab + (cur.key.x & 0xffff);
                const Real swn = lds_pair<P>(rec_a + slot * RB + O_AUX).x;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = sizeof(Real) == 8 ? (i >> 1) * 16 + (i & 1) : i;    // element of g-point q + 8 i in this lane's row slice
                    tau[i] = fma_(swn, bil4(c0, c1, c2, c3, k[kC00 + e], k[kC10 + e], k[kC01 + e], k[kC11 + e]), tau[i]);
                }
            }
            const Real swt = cur.swt;
            // the cache has been read: fetch the next position's corners under the rest of this layer's work
            if (STAGED) {
                if (more) load_rec(nslot_r, nxt, false);
            } else if (more) {
                if (ls == L - 1) mbar_wait_a<B_RFULL>(a_bar + (nslot_r / L) * 8, nslot_r == 0 ? rpar ^ 1 : rpar);
                load_rec(nslot_r, nxt, true);
                // "moved" compares the whole word: a flip of kNegFlag alone reloads the same rows, harmlessly
                refill_bil(nxt.key.x & 0xffff, nxt.key.x != cur.key.x);
                refill_lut(nxt.key.y, nxt.key.y != cur.key.y);
            }
            Real ssa[4];
            if (SW) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const Real ray = mul_(swt, sigma[i]);             // :313-316 (moles_per_layer == simple_weight, same product)
                    tau[i] = add_(tau[i], ray);                       // :456
                    ssa[i] = div_fast(ray, tau[i]);                   // :459
                }
            }
            if (ls == 0) mbar_wait_a<B_TEMPTY>(a_bar + b * 8, tpar);
            const unsigned sa = stage_a + (b * L + ls) * (NOUT * TILEB);
            if (!ECCKD_DBG(16)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    sts_<Real>(sa + i * 1024, tau[i]);
                    if (SW) sts_<Real>(sa + TILEB + i * 1024, ssa[i]);
                }
            }
            if (++ls == L) {
                ls = 0;
                if (!ECCKD_DBG(16)) fence_proxy_async_smem();      // generic-proxy writes -> visible to the async (TMA) proxy
                __syncwarp();
                if (lane0) {
                    mbar_arrive_a<B_TFULL>(a_bar + b * 8);
                    mbar_arrive_a<B_REMPTY>(a_bar + (slot / L) * 8);
                }
                if (++b == NBUF) { b = 0; tpar ^= 1; }
            }
            cur = nxt;
            slot = nslot_r;
            if (slot == 0) rpar ^= 1;
        }
    }
}

// ============================================================================
// launcher
// ============================================================================
// 3-D map of an (ncol, nlay, ngpt) output, box = 128 bytes of columns x 1 layer x 32 g-points, 128-byte swizzle
template <typename Real> static bool encode_quad_map(CUtensorMap *map, Real *ptr, int ncol, int nlay, int ngpt)
{
    auto enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t w = sizeof(Real);
    const cuuint64_t gdim[3] = {(cuuint64_t)ncol, (cuuint64_t)nlay, (cuuint64_t)ngpt};
    const cuuint64_t gstr[2] = {(cuuint64_t)ncol * w, (cuuint64_t)ncol * nlay * w};
    const cuuint32_t box[3] = {(cuuint32_t)(128 / w), 1, 32}, estr[3] = {1, 1, 1};
    return enc(map, w == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ptr, gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename Real> bool quad_eligible(const CallGases &cg, const KernelArgs<Real> &a)
{
    if (cg.nvar > kQuadMaxVar + kQuadMaxExtra) return false;
    if ((size_t)a.ncol * sizeof(Real) % 16) return false;            // tensor-map strides are multiples of 16 bytes
    for (int o = 0; o < 3; ++o) if (a.out[o] && (reinterpret_cast<uintptr_t>(a.out[o]) & 15)) return false;
    return tensor_map_encoder() != nullptr;
}

template <typename Real, bool SW, int NV, bool LUT32, int R, int NBUF, int RC = 184, int RI = 64, int RH = 64, bool ZERO_BY_TMA = false, int L = 1, int NX = 0, bool STAGED = false, bool HIDX = false, bool PHELP = false>
static cudaError_t launch_quad_t(const DevModel &m, const CallGases &cg, const KernelArgs<Real> &a, const Real *mix, int sms, cudaStream_t s)
{
    constexpr int NOUT = SW ? 2 : 1;
    constexpr int TILEB = (sizeof(Real) == 8 ? 2 : 1) * 32 * 128;
    TmaMaps maps{};
    for (int o = 0; o < (SW ? 3 : 1); ++o)
        if (a.out[o] && !encode_quad_map<Real>(&maps.m[o], a.out[o], a.ncol, a.nlay, m.ngpt)) return cudaErrorInvalidValue;
    const size_t smem = (size_t)NBUF * L * NOUT * TILEB + (SW ? 32 * 128 : 0) + sizeof(QRec<Real, NV, NX>) * R +
                        (2 * NBUF + 2 * R / L) * sizeof(unsigned long long) + (SW ? 0 : sizeof(float2) * (m.nplanck + 1) * kGP + 16);
    auto kern = quad_kernel<Real, SW, NV, LUT32, R, NBUF, RC, RI, RH, ZERO_BY_TMA, L, NX, STAGED, HIDX, PHELP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
#ifdef ECCKD_TUNE
    if (getenv("ECCKD_CARVEOUT")) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("ECCKD_CARVEOUT")));
    if (getenv("ECCKD_VERBOSE")) fprintf(stderr, "quad smem %zu bytes\n", smem);
#endif
    const int ntile = (a.ncol + kQCols - 1) / kQCols;
    int blocks = ntile < sms ? ntile : sms;
    if (blocks < 1) blocks = 1;
#ifdef ECCKD_TUNE
    if (getenv("ECCKD_BLOCKS")) blocks = std::min(ntile, atoi(getenv("ECCKD_BLOCKS")));
#endif
    dim3 grid((unsigned)blocks, (unsigned)((m.ngpt + kGP - 1) / kGP));
    kern<<<grid, kQWarps * 32, smem, s>>>(m, cg, a, mix, maps);
    return cudaGetLastError();
}

template <typename Real, bool SW>
cudaError_t launch_quad(const DevModel &m, const CallGases &cg, const KernelArgs<Real> &a, const Real *mix, int sms, cudaStream_t s)
{
#ifdef ECCKD_TUNE
    // tuning build: ECCKD_QUAD_VARIANT selects an alternative instantiation (tools/exp_quad.sh)
    static const int v = getenv("ECCKD_QUAD_VARIANT") ? atoi(getenv("ECCKD_QUAD_VARIANT")) : 0;
    if (cg.nvar <= 1) {
        //                                  NV LUT32  R NBUF   RC   RI   RH  ZERO_BY_TMA L NX STAGED HIDX
        if (v == 1) return launch_quad_t<Real, SW, 1, false, 16, 6, 184, 64, 64, true, 1, 0, false, false>(m, cg, a, mix, sms, s);
        if (v == 2) return launch_quad_t<Real, SW, 1, false, 16, 6, 168, 80, 96, false, 1, 0, false, false>(m, cg, a, mix, sms, s);
        if constexpr (SW) {
            if (v == 20) return launch_quad_t<Real, SW, 1, false, 16, 6, 184, 64, 64, true, 1, 0, true>(m, cg, a, mix, sms, s);
            if (v == 21) return launch_quad_t<Real, SW, 1, false, 8, 6, 184, 64, 64, true, 1, 0, true>(m, cg, a, mix, sms, s);
        } else {
            if (v == 20) return launch_quad_t<Real, SW, 1, false, 16, 6, 168, 80, 96, false, 1, 0, true>(m, cg, a, mix, sms, s);
            if (v == 21) return launch_quad_t<Real, SW, 1, false, 8, 6, 168, 80, 96, false, 1, 0, true>(m, cg, a, mix, sms, s);
            if (v == 22) return launch_quad_t<Real, SW, 1, false, 8, 6, 176, 72, 88, false, 1, 0, true>(m, cg, a, mix, sms, s);
        }
        // shared-memory footprint (what is left of the 256 KB is the L1 that holds the table rows between columns)
        if constexpr (SW) {
            if (v == 10) return launch_quad_t<Real, SW, 1, false, 8, 6, 184, 64, 64, true>(m, cg, a, mix, sms, s);
            if (v == 11) return launch_quad_t<Real, SW, 1, false, 8, 4, 184, 64, 64, true>(m, cg, a, mix, sms, s);
            if (v == 12) return launch_quad_t<Real, SW, 1, false, 16, 4, 184, 64, 64, true>(m, cg, a, mix, sms, s);
            if (v == 13) return launch_quad_t<Real, SW, 1, false, 8, 3, 184, 64, 64, true>(m, cg, a, mix, sms, s);
            if (v == 14) return launch_quad_t<Real, SW, 1, false, 8, 5, 184, 64, 64, true>(m, cg, a, mix, sms, s);
        } else {
            if (v == 10) return launch_quad_t<Real, SW, 1, false, 8, 6, 168, 80, 96, false>(m, cg, a, mix, sms, s);
            if (v == 11) return launch_quad_t<Real, SW, 1, false, 8, 4, 168, 80, 96, false>(m, cg, a, mix, sms, s);
            if (v == 12) return launch_quad_t<Real, SW, 1, false, 16, 4, 168, 80, 96, false>(m, cg, a, mix, sms, s);
            if (v == 13) return launch_quad_t<Real, SW, 1, false, 8, 3, 168, 80, 96, false>(m, cg, a, mix, sms, s);
            if (v == 14) return launch_quad_t<Real, SW, 1, false, 8, 5, 168, 80, 96, false>(m, cg, a, mix, sms, s);
        }
        if constexpr (SW) {
            if (v == 3) return launch_quad_t<Real, SW, 1, false, 16, 6, 184, 64, 64, true, 1, 0, false, true>(m, cg, a, mix, sms, s);
            if (v == 4) return launch_quad_t<Real, SW, 1, false, 16, 6, 176, 80, 80, true, 1, 0, false, true>(m, cg, a, mix, sms, s);
            if (v == 5) return launch_quad_t<Real, SW, 1, false, 16, 6, 184, 72, 72, true, 1, 0, false, true>(m, cg, a, mix, sms, s);
            if (v == 6) return launch_quad_t<Real, SW, 1, false, 32, 6, 184, 64, 64, true, 1, 0, false, true>(m, cg, a, mix, sms, s);
        }
    }
#endif
    // Production instantiations, as measured on B200 (DESIGN.md sections 4.2 and 5; every figure an A/B inside one gpurun call):
    //   nvar <= 1   LW: 168 registers per compute thread, 96 for the Planck
    //               warps (batches of four g-points in flight), 80 for the index warps; record ring of 8 (131 KB of shared
    //               memory, so that 124 KB of L1 are left for the table rows: 1.98 -> 1.97 ms).  SW: no Planck role, the
    //               g = 0 plane leaves through the TMA unit, the compute warps take the registers nobody else needs
    //   nvar <= 3   three varying gases cached, look-up-table rows cached as fp32 to make room
    //   nvar 4..7   the other gases read uncached every layer
    //                                          NV  LUT32   R NBUF   RC   RI   RH  ZERO_BY_TMA  L NX STAGED HIDX PHELP
    if (cg.nvar <= 1) {
        // staged refills: LW 2.05 -> 2.03 ms, fp32 LW 1.97 -> 1.91, fp32 SW 1.20 -> 1.13; fp64 SW 1.61 -> 1.65 with 184 registers
        // per compute thread, 1.55 -> 1.54 with 192
        constexpr bool kStaged = true;
        // SW leaves 16 registers per thread unclaimed at 184 / 64 / 64: the fp64 compute warps take them (1.58 -> 1.55 ms; fp32 loses 2 %)
        constexpr int kSwRc = sizeof(Real) == 8 ? 192 : 184;
        // fp64 LW caches the look-up-table rows as fp32 (exact: the tables are fp32 on file): half the bytes of the largest
        // refill for 32 conversions per iteration, which the conversion unit can take now that the Planck warps do not use it
        // (1.99 -> 1.89 ms; with other register splits -- 160/88/104, 152/96/112, 144/96/128 -- 1.96 - 2.03 ms).  SW keeps
        // fp64 rows: its ssa divisions share that unit (1.56 -> 1.66 ms with fp32 rows)
        constexpr bool kLwLut32 = sizeof(Real) == 8;
        // fp32 SW (bound by issue slots) hands tiles and records over in groups of two layers: half the barrier operations
        // per layer (1.125 -> 1.099 ms; fp32 LW 1.60 -> 1.65, fp64 no change: left at one)
        constexpr int kSwL = sizeof(Real) == 4 ? 2 : 1;
        if constexpr (SW) return launch_quad_t<Real, SW, 1, false, 16, 6, kSwRc, 64, 64, true, kSwL, 0, kStaged>(m, cg, a, mix, sms, s);
        // index warps helping with the Planck units: fp32 LW 1.85 -> 1.60 ms (there the three Planck warps set the pace);
        // fp64 LW 2.03 -> 2.06 ms (there the compute warps do, and the helpers' extra instructions compete with them)
        else return launch_quad_t<Real, SW, 1, kLwLut32, 8, 6, 168, 80, 96, false, 1, 0, kStaged, false, sizeof(Real) == 4>(m, cg, a, mix, sms, s);
    }
    if (cg.nvar <= kQuadMaxVar) {      // up to three varying gases, all cached (fp32 LUT rows to make room)
        if constexpr (SW) return launch_quad_t<Real, SW, kQuadMaxVar, true, 16, 6, 184, 64, 64, true>(m, cg, a, mix, sms, s);
        else return launch_quad_t<Real, SW, kQuadMaxVar, true, 16, 6, 176, 72, 88, false>(m, cg, a, mix, sms, s);
    }
    // four to seven varying gases: three cached, the others read uncached every layer
    if constexpr (SW) return launch_quad_t<Real, SW, kQuadMaxVar, true, 8, 6, 184, 96, 48, true, 1, kQuadMaxExtra>(m, cg, a, mix, sms, s);
    else return launch_quad_t<Real, SW, kQuadMaxVar, true, 8, 6, 176, 80, 80, false, 1, kQuadMaxExtra>(m, cg, a, mix, sms, s);
}

#define ECCKD_QUAD_INST(Real)                                                                                                   \
    template bool quad_eligible<Real>(const CallGases &, const KernelArgs<Real> &);                                              \
    template cudaError_t launch_quad<Real, false>(const DevModel &, const CallGases &, const KernelArgs<Real> &, const Real *, int, cudaStream_t); \
    template cudaError_t launch_quad<Real, true>(const DevModel &, const CallGases &, const KernelArgs<Real> &, const Real *, int, cudaStream_t);
ECCKD_QUAD_INST(double)
ECCKD_QUAD_INST(float)

}  // namespace ecckd
