// ecckd_kernels.cu -- sm_100a kernels of the ECCKD gas-optics hot path.
//
// Replaces the loop nests of /root/reference/src/gas_optics_ecckd.f90:
//   calculate_optical_depth :117-240 (once per gas) + accumulate :370
//   calculate_planck_function :273-288 (x3) + shifted copies :423-424
//   calculate_rayleigh_optical_depth :313-317 + ssa/g :456-460 + toa_src :468-472
// with ONE fused pass per call: interpolation indices once per (column, layer),
// all gases accumulated in registers, every output element stored exactly once.
//
// This file: the fallback kernels, the per-call table mixing and the launch dispatch.  The product kernel is
// quad_kernel in ecckd_quad.cu (four columns per compute warp, TMA tensor stores).
//   walk_kernel    -- round 1's product kernel, now the fallback for calls the quad kernel cannot map (odd column
//                     counts, unaligned outputs, more than seven varying gases).  A warp-specialised pipeline in a
//                     persistent block: index warps (lane = layer) -> compute warps (one column each, lanes =
//                     g-points, corner rows cached in registers) -> TMA store warp or writer warps; Planck warps
//                     (lanes = columns) write the sources directly.  Same arithmetic as quad_kernel, bit for bit.
//   generic_kernel -- thread per (column, layer), loop over g; handles any gas mix (several LUT gases, tables with
//                     negative entries or not exact in fp32, more than 64 pressure levels) with the reference's
//                     per-g clamp.  Also an independent second formulation in the GPU tests.
//   mix_kernel     -- per call: gases with a call-constant weight factor pre-mixed into one table.
#include "ecckd_device.cuh"
#include "ecckd_kernels.h"
#include "ecckd_pipeline.cuh"
#include <cstdlib>
#include <cstdint>

namespace ecckd {

// ============================================================================
// generic kernel
// ============================================================================
template <typename Real, bool SW>
__global__ void __launch_bounds__(256)
generic_kernel(const __grid_constant__ DevModel m, const __grid_constant__ CallGases cg,
               const __grid_constant__ KernelArgs<Real> a)
{
    const long long ncell = (long long)a.ncol * a.nlay;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncell) return;
    const long long col = idx % a.ncol, lay = idx / a.ncol;
    const int ngpt = m.ngpt, np = m.np, nt = m.nt;
    const long long ncol = a.ncol, nlay = a.nlay;
    const Real *ktab = TableOf<Real>::k(m);

    const double pu = (double)a.plev[col + ncol * lay], pl = (double)a.plev[col + ncol * (lay + 1)];
    const CellIndex ci = cell_index(m, pu, pl, (double)a.tlay[col + ncol * lay]);
    const Real pw1 = (Real)ci.pw1, pw0 = (Real)__dsub_rn(1.0, ci.pw1);
    const Real tw1 = (Real)ci.tw1, tw0 = (Real)__dsub_rn(1.0, ci.tw1);
    const long long s_t = (long long)ngpt * np, s_v = s_t * nt;
    const long long corner = (long long)ngpt * (ci.ip0 - 1) + s_t * (ci.it0 - 1);

    // per-gas weights and LUT indices, once per cell
    Real w[kMaxGases], vw1[kMaxGases];
    long long base[kMaxGases];
    for (int q = 0; q < cg.n; ++q) {
        const GasRef &gr = cg.g[q];
        double vmr = gr.code == kNone ? 0.0 : load_vmr<Real>(gr, col, lay);
        w[q] = (Real)gas_weight(gr, ci.sw, vmr);
        base[q] = gr.table_off + corner;
        vw1[q] = 0;
        if (gr.code == kLut) {
            double v1;
            int iv0 = vmr_index(gr, vmr, v1);
            vw1[q] = (Real)v1;
            base[q] += s_v * (iv0 - 1);
        }
    }

    // Planck factors (LW)
    Real a_lay0 = 0, a_lay1 = 0, a_up0 = 0, a_up1 = 0, a_dn0 = 0, a_dn1 = 0;
    int r_lay = 0, r_up = 0, r_dn = 0;
    const Real *ptab = TableOf<Real>::planck(m);
    if (!SW) {
        double a0, a1;
        int it = planck_index(m, (double)a.tlay[col + ncol * lay], a0, a1);
        r_lay = max(it, 1) - 1; a_lay0 = (Real)a0; a_lay1 = (Real)a1;
        if (a.tlev) {
            it = planck_index(m, (double)a.tlev[col + ncol * lay], a0, a1);
            r_up = max(it, 1) - 1; a_up0 = (Real)a0; a_up1 = (Real)a1;
            it = planck_index(m, (double)a.tlev[col + ncol * (lay + 1)], a0, a1);
            r_dn = max(it, 1) - 1; a_dn0 = (Real)a0; a_dn1 = (Real)a1;
        }
    }
    const Real pi = (Real)m.pi;

    for (int g = 0; g < ngpt; ++g) {
        Real tau = 0;
        for (int q = 0; q < cg.n; ++q) {
            const Real *k = ktab + base[q] + g;
            // :198-203 / :216-221 grouping
            Real od = tw0 * (pw0 * k[0] + pw1 * k[ngpt]) + tw1 * (pw0 * k[s_t] + pw1 * k[s_t + ngpt]);
            if (cg.g[q].code == kLut) {
                const Real *k2 = k + s_v;                                              // :167-178
                Real od1 = tw0 * (pw0 * k2[0] + pw1 * k2[ngpt]) + tw1 * (pw0 * k2[s_t] + pw1 * k2[s_t + ngpt]);
                od = ((Real)1 - vw1[q]) * od + vw1[q] * od1;
            }
            od = w[q] * od;
            if (od < (Real)0) od = 0;                                                  // :233-238
            tau += od;                                                                 // :370
        }
        const long long o = col + ncol * (lay + nlay * g);
        if (SW) {
            const Real ray = (Real)ci.sw * (Real)m.rayleigh[g];                        // :313-316
            tau += ray;                                                                // :456
            store_cs(a.out[0] + o, tau);
            if (a.out[1]) store_cs(a.out[1] + o, ray / tau);                           // :459
            if (a.out[2]) store_cs(a.out[2] + o, (Real)0);                             // :460
        } else {
            store_cs(a.out[0] + o, tau);
            store_cs(a.out[1] + o, (a_lay0 * ptab[r_lay * ngpt + g] + a_lay1 * ptab[(r_lay + 1) * ngpt + g]) / pi);
            if (a.tlev) {
                store_cs(a.out[2] + o, (a_dn0 * ptab[r_dn * ngpt + g] + a_dn1 * ptab[(r_dn + 1) * ngpt + g]) / pi); // :423
                store_cs(a.out[3] + o, (a_up0 * ptab[r_up * ngpt + g] + a_up1 * ptab[(r_up + 1) * ngpt + g]) / pi); // :424
            }
        }
    }
    if (lay == 0) {
        if (SW) {
            if (a.out2d)
                for (int g = 0; g < ngpt; ++g) a.out2d[col + ncol * g] = (Real)m.solar[g];      // :468-472
        } else {
            double a0, a1;
            int it = planck_index(m, (double)a.tsfc[col], a0, a1);                               // :408-412
            const int r = max(it, 1) - 1;
            for (int g = 0; g < ngpt; ++g)
                a.out2d[col + ncol * g] = ((Real)a0 * ptab[r * ngpt + g] + (Real)a1 * ptab[(r + 1) * ngpt + g]) / pi;
        }
    }
}

// ============================================================================
// column-walking kernel
// ============================================================================
// ---- per-call table mixing ----------------------------------------------------
// out[sign][group][ip][it][lane]: sign 0 = sum of f_g*K_g over gases with f_g not < 0,
// sign 1 = the same over gases with f_g < 0 (only reachable by cells whose
// simple_weight is negative, i.e. upside-down layers -- kept so that the
// reference's sign logic :143-149,233-238 is reproduced for them too).
// f_g: 1 (code 0, :216), vmr (code 1, :148), vmr - reference (code 3, :146).
template <typename Real>
__global__ void __launch_bounds__(256)
mix_kernel(const __grid_constant__ DevModel m, const __grid_constant__ CallGases cg, Real *out)
{
    const int ngroups = (m.ngpt + kGP - 1) / kGP;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ngroups * kSlab) return;
    const int grp = e / kSlab, cell = e % kSlab;
    // e is a position of the OUTPUT row layout (Real-wide rows); the fp64 source tables use the 8-byte layout
    const int src_cell = (cell & ~(kGP - 1)) + walk_pos64(walk_inv<Real>(cell & (kGP - 1)));
    const double *src = m.wbil64 + (size_t)grp * m.nslot * kSlab + src_cell;
    double plus = 0.0, minus = 0.0;
    for (int q = 0; q < cg.nmix; ++q) {
        const GasRef &gr = cg.g[cg.mix_gas[q]];
        const double vmr = gr.code == kNone ? 0.0 : load_vmr<Real>(gr, 0, 0);
        const double f = gr.code == kNone ? 1.0 : (gr.code == kRelLinear ? __dsub_rn(vmr, gr.ref) : vmr);
        const double kv = src[cg.mix_off[q]];
        if (f < 0.0) minus += f * kv; else plus += f * kv;
    }
    out[e] = (Real)plus;
    out[(size_t)ngroups * kSlab + e] = (Real)minus;
}

// ---- records ------------------------------------------------------------------
// Per-(column, layer) record produced by the index warps and read back by the
// compute warp of that column as warp-wide broadcasts (every lane reads the same
// address: one wavefront per 16-byte vector load).
template <typename Real, int NB> struct alignas(16) Rec {
    Real c[4];          // tw0*pw0, tw0*pw1, tw1*pw0, tw1*pw1
    Real w[NB];         // w[0] = simple_weight (mixed table; also SW's moles_per_layer); w[s] = weight of varying gas s-1
    Real wl, vw1;       // LUT gas: weight, vmr_weight1
    alignas(16) int key;// bilinear corner offset (ip*kNTP + it)*kGP  |  mask << 16
                        //   mask bit s: slot s contributes; bit 8: LUT gas; bit 9: negative-weight mix (slow path)
    int hkoff;          // LUT corner offset ((iv*kNPP + ip)*kNTP + it)*kGP
    int pad0, pad1;
};
constexpr int kMaskShift = 16, kMaskLut = 0x100 << kMaskShift, kMaskNeg = 0x200 << kMaskShift;

// ============================================================================
// walk_kernel: a warp-specialised pipeline inside one persistent block
//
//   index warps   (NI)   lane l computes the interpolation record of layer j0+l of
//                        one column: log-pressure / temperature / mole-fraction
//                        indices and weights -- long dependent chains (log, IEEE
//                        divisions), one chunk of 32 layers ahead of the walk
//   compute warps (TC/CPW) one column each, lanes = g-points: walk the layers top to
//                        bottom, corner rows cached in registers, accumulate tau (and
//                        ssa), write them transposed into a ring of staging tiles
//                        [output][layer][g][column]
//   store warp    (TMA)  one thread turns finished tiles into cp.async.bulk.tensor
//                        stores of the caller's (ncol, nlay, ngpt) arrays; or
//   writer warps  (NWW)  the same with full-line streaming stores (fp32, odd ncol)
//   Planck warps  (NPW)  longwave: lanes = (column, level), interpolate the Planck
//                        LUT from shared memory and store lay_source / lev_source_inc /
//                        lev_source_dec / sfc_source directly, bypassing the tiles
//
// Hand-offs are mbarrier full/empty pairs, so no role ever waits for a peer of
// the same role: index -> compute through double-buffered per-column record
// chunks, compute -> store through NBUF tiles of L layers.  A block walks chunks of
// one or two 16-column tiles, strided by the grid (see the comment at the top of the
// kernel body); DESIGN.md 4.2 has the measured rationale of every choice below.
// NB = register-cached bilinear slots: slot 0 is the per-call mixed table, slots
// 1..NB-1 the gases whose concentration varies over the call.
// RF = flavour of the compute warps' cache refill, a bit set tuned per band on B200:
//   1 branch-free refill (all loads predicated)      2 loop invariants pinned in registers
//   4 tile-free barrier probed before the arithmetic  8 refill issued after the staging fence ("late")
//  16 refill issued as soon as the bilinears have read the cache, ahead of the rest of the layer's arithmetic
// ============================================================================
// Register rebalancing (RC > 0): the roles are padded to whole warpgroups (4 warps) and trade registers with
// setmaxnreg -- index warps keep RI, store / Planck warps drop to RH, compute warps grow to RC.  Each setmaxnreg sits
// inside its role's branch, which never rejoins the others: ptxas budgets registers per control-flow region.
template <int NWW, int NPW, int RC> struct PadWarps { static constexpr int value = RC ? (4 - (NWW + NPW) % 4) % 4 : 0; };

template <typename Real, bool SW, int NB, int TC, int L, int NBUF, int NI, int NWW, bool TMA, int CPW, int NPW, int RF = 7,
          int RC = 0, int RI = 64, int RH = 56>
__global__ void __launch_bounds__((TC / CPW + NI + NWW + NPW + PadWarps<NWW, NPW, RC>::value) * 32, 1)
walk_kernel(const __grid_constant__ DevModel m, const __grid_constant__ CallGases cg,
            const __grid_constant__ KernelArgs<Real> a, const Real *__restrict__ mixtab, const __grid_constant__ TmaMaps maps,
            const int tiles_per_chunk)
{
    constexpr int NOUT = SW ? 2 : 1;      // staged outputs: tau (+ ssa).  SW's g plane is zero-filled by the store warp;
                                          // LW's three Planck sources are written by the Planck warps, straight to global memory
    static_assert(SW || NPW >= 1, "longwave needs Planck warps");
    constexpr bool ZERO_BY_WARP = SW && NPW >= 1;   // SW: helper warps write the g = 0 plane with plain full-line stores
    // writer-warp path: padded column stride (conflict-free transpose both ways).
    // TMA path: dense 128-byte rows [g][column] in the 128B-swizzle pattern the tensor map expects.
    constexpr int TCP = TMA ? TC : TC + 1;
    constexpr int ROWS = 32 * TCP;        // one (output, layer) tile: 32 g-points x TC columns
    constexpr int TILE = NOUT * L * ROWS; // one ring entry
    constexpr int NCW = TC / CPW;          // compute warps
    
    static_assert(CPW == 1 || (CPW == 2 && TMA), "two columns per warp needs the 16-byte swizzled staging of the TMA path");
    static_assert(!(RF & 8) || L == 1, "the late refill is issued at tile hand-over, i.e. once per layer only when L == 1");
    static_assert(!TMA || (L == 1 && NWW == 1 && sizeof(Real) * TC == 128), "TMA store path: one layer per tile, 128-byte rows, one store warp");
    using RecT = Rec<Real, NB>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // shared-memory carve-up (walk_smem_bytes() on the host mirrors it)
    RecT *recs = reinterpret_cast<RecT *>(smem_raw);                                  // [2][TC][32]
    Real *stage = reinterpret_cast<Real *>(smem_raw + sizeof(RecT) * 2 * TC * 32);     // [NBUF][TILE]   (1024-byte aligned)
    Real *zero_tile = stage + (size_t)NBUF * TILE;                                     // TMA && SW: source of the g = 0 plane
    unsigned long long *bar_full = reinterpret_cast<unsigned long long *>(zero_tile + ((TMA && SW) ? ROWS : 0));
    unsigned long long *bar_empty = bar_full + NBUF;
    unsigned long long *rec_full = bar_empty + NBUF;       // [2][TC]
    unsigned long long *rec_empty = rec_full + 2 * TC;     // [2][TC]
    float *spl = reinterpret_cast<float *>(rec_empty + 2 * TC);   // LW: planck_function of this g-group, fp32 (exact), [g][row]

    // Roles are laid out in LOGICAL warp order [compute | index | store | Planck]; hardware warp ids run the other way
    // round because the SM's issue arbiter prefers the highest warp id: the latency-bound compute warps get priority
    // over the throughput-oriented Planck warps.
    constexpr int HB = NCW + NI + NWW;              // first helper (Planck / zero-plane) warp
    constexpr int NWARPS = HB + NPW + PadWarps<NWW, NPW, RC>::value;   // (+ idle warps that complete the helper warpgroup)
    static_assert(RC == 0 || ((TC / CPW) % 4 == 0 && NI % 4 == 0), "register rebalancing works on whole warpgroups");
    const int warp = NWARPS - 1 - (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int ncol = a.ncol, nlay = a.nlay;
    // Persistent blocks: the columns are cut into chunks of CH tiles of TC columns (CH = 2 for large calls: the
    // Planck warps then store 256-byte runs) and block b walks chunks b, b + gridDim.x, ...  The barriers, the Planck
    // table in shared memory and the register caches are set up once; the index warps run ahead into the next tile
    // while the compute warps finish the current one; and because the blocks advance through neighbouring chunks
    // together, the stores of all SMs stay close together in memory (measured: giving each block one contiguous
    // column range instead costs 7 %).
    const int CH = tiles_per_chunk;
    const int ntile_all = (ncol + TC - 1) / TC;
    const int nchunk_all = (ntile_all + CH - 1) / CH;
    const int nchunk = (nchunk_all - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // chunks of this block (>= 1)
    const int nslot = nchunk * CH;                                                  // tile slots of this block
    // tile slot -> global tile (>= ntile_all: the empty second half of the last chunk; every role skips it)
    auto tile_of = [&](int ts) { const int k = ts / CH; return ((int)blockIdx.x + k * (int)gridDim.x) * CH + (ts - k * CH); };
    const int ngpt = m.ngpt;
    const int grp = blockIdx.y, g0 = grp * kGP;
    const bool has_lut = cg.lut >= 0, has_mix = cg.nmix > 0;
    const int nvar = cg.nvar;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < NBUF; ++b) { mbar_init(&bar_full[b], NCW); mbar_init(&bar_empty[b], TMA ? 1 : NWW); }
        for (int q = 0; q < 2 * TC; ++q) { mbar_init(&rec_full[q], 1); mbar_init(&rec_empty[q], 1); }
    }
    if (!SW) {   // Planck LUT of this g-group -> shared memory, transposed to [g][row] (29.6 KB for 231 rows): the Planck
                 // warps (lanes = columns) gather rows that differ per lane, which then fall into different banks
        // each entry holds the row AND its upper neighbour, so one 8-byte LDS fetches both interpolation operands
        const int np_ = m.nplanck;
        float2 *spl2 = reinterpret_cast<float2 *>(spl);
        for (int q = threadIdx.x; q < np_ * kGP; q += blockDim.x) {
            const int row = q / kGP, gl = q % kGP;
            const bool gok = g0 + gl < ngpt;
            const float p0 = gok ? m.planck32[(size_t)row * ngpt + g0 + gl] : 0.0f;
            const float p1 = (gok && row + 1 < np_) ? m.planck32[(size_t)(row + 1) * ngpt + g0 + gl] : 0.0f;
            spl2[gl * np_ + row] = make_float2(p0, p1);
        }
    }
    if (TMA && SW) {
        for (int q = threadIdx.x; q < ROWS; q += blockDim.x) zero_tile[q] = 0;
        fence_proxy_async_smem();
    }
    __syncthreads();
    const int nround = (nlay + 31) / 32;

    if (ECCKD_DBG(128) && warp < HB) return;   // experiment: Planck warps only
    if (warp >= NCW + NI) {      // ---- store / writer / helper warps (every path below returns) ----
    if constexpr (RC > 0) reg_dec<RH>();
    // =========================== TMA store warp ==================================
    if (TMA && warp >= NCW + NI && warp < HB) {
        if (lane == 0) {
            int b = 0, par = 0;                             // ring slot and its phase, running over all tiles
            for (int ts = 0; ts < nslot; ++ts) {
                const int gt = tile_of(ts);
                if (gt >= ntile_all) continue;
                const int col0 = gt * TC;
                for (int n = 0; n < nlay; ++n) {            // one staging tile per layer
                    mbar_wait_relaxed(&bar_full[b], par, kStoreNap);
                    const Real *sb = stage + (size_t)b * TILE;
#pragma unroll
                    for (int o = 0; o < NOUT; ++o)
                        if (a.out[o]) tma_store_3d(&maps.m[o], sb + o * ROWS, col0, n, g0);
                    if (SW && !ZERO_BY_WARP && a.out[2]) tma_store_3d(&maps.m[2], zero_tile, col0, n, g0);     // g = 0  :460
                    tma_commit();
                    tma_wait_read<0>();                     // the tile has been read out of shared memory: hand it back
                    mbar_arrive(&bar_empty[b]);
                    if (++b == NBUF) { b = 0; par ^= 1; }
                }
            }
            tma_wait_all<0>();
        }
        return;
    }

    // =========================== writer warps ====================================
    if (!TMA && warp >= NCW + NI && warp < NCW + NI + NWW) {
        constexpr int NWT = NWW * 32;                 // writer threads
        constexpr int GSTEP = NWT / TC;               // g-points covered per pass
        const int t = (warp - (NCW + NI)) * 32 + lane;
        const int c = t % TC, gsub = t / TC;
        const long long plane = (long long)ncol * nlay * (long long)sizeof(Real);   // bytes between g-points
        const long long pitch = (long long)ncol * (long long)sizeof(Real);          // bytes between layers
        const int nchunk_lay = (nlay + L - 1) / L;
        int b = 0, par = 0;                               // ring slot and its phase, running over all tiles
        for (int tn = 0; tn < nslot * nchunk_lay; ++tn) {
            const int n = tn % nchunk_lay;
            const int gt = tile_of(tn / nchunk_lay);
            if (gt >= ntile_all) continue;
            const long long col0 = (long long)gt * TC;
            const bool c_ok = col0 + c < ncol;
            const int lay0 = n * L, nl = min(L, nlay - lay0);
            mbar_wait_relaxed(&bar_full[b], par, kStoreNap);
            const Real *sb = stage + (size_t)b * TILE + c;
            const long long off0 = ((col0 + c) + (long long)ncol * lay0) * (long long)sizeof(Real) + plane * (g0 + gsub);
            if (c_ok && !ECCKD_DBG(1)) {
#pragma unroll 2
                for (int gi = 0; gi < 32 / GSTEP; ++gi) {
                    const int gg = gsub + gi * GSTEP;
                    if (g0 + gg < ngpt) {
                        Real v[NOUT * L];
#pragma unroll
                        for (int q = 0; q < NOUT * L; ++q) v[q] = sb[q * ROWS + gg * TCP];
                        const long long off = off0 + plane * (gi * GSTEP);
#pragma unroll
                        for (int o = 0; o < NOUT; ++o) {
                            char *dst = reinterpret_cast<char *>(a.out[o]) + off;
#pragma unroll
                            for (int l2 = 0; l2 < L; ++l2)
                                if (a.out[o] && l2 < nl) store_cs(reinterpret_cast<Real *>(dst + pitch * l2), v[o * L + l2]);
                        }
                        if (SW && !ZERO_BY_WARP && a.out[2]) {
                            char *dst = reinterpret_cast<char *>(a.out[2]) + off;
#pragma unroll
                            for (int l2 = 0; l2 < L; ++l2)
                                if (l2 < nl) store_cs(reinterpret_cast<Real *>(dst + pitch * l2), (Real)0);   // g = 0  :460
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[b]);
            if (++b == NBUF) { b = 0; par ^= 1; }
        }
        return;
    }

    // =========================== zero-plane warps (SW) ===========================
    if (SW && NPW >= 1 && warp >= HB && warp < HB + NPW) {
        constexpr int LPU = 32 / TC;
        const int pw = warp - HB;
        const int c = lane % TC, lsel = lane / TC;
        const int ng = min(kGP, ngpt - g0);
        for (int ts = 0; ts < nslot; ++ts) {
            const long long col0 = (long long)tile_of(ts) * TC;
            if (a.out[2] && col0 + c < ncol) {
                const unsigned plane32 = (unsigned)((long long)ncol * nlay);
                for (int j = pw * LPU + lsel; j < nlay; j += NPW * LPU) {
                    unsigned off = (unsigned)(col0 + c + (long long)ncol * j) + plane32 * (unsigned)g0;
#pragma unroll 4
                    for (int gq = 0; gq < ng; ++gq, off += plane32) store_cs(a.out[2] + off, (Real)0);     // g = 0  :460
                }
            }
        }
        return;
    }

    // =========================== Planck warps (LW) ===============================
    // lanes = (column, level).  The block's column range is cut into strips of W columns -- 32 when the block holds
    // more than one tile (one level per unit, 256-byte store runs), else TC -- and a unit is one strip x 32/W levels.
    // Per g-point every lane interpolates its own rows of the shared-memory LUT and the warp stores full column runs
    // of lay_source, lev_source_inc and lev_source_dec directly: these three outputs (75 % of the longwave bytes)
    // never touch the staging tiles.  lev_source_inc(:,j-1) and lev_source_dec(:,j) are both B(tlev(:,j)) (:423-424):
    // the lane of level j evaluates it once and stores it twice, so a cell costs two Planck evaluations, not three.
    if (!SW && warp >= HB && warp < HB + NPW) {
        if (ECCKD_DBG(64)) return;
        const int pw = warp - HB;
        const int W = CH * TC;                        // strip = chunk (W divides 32)
        const int LPU = 32 / W;                       // levels per unit
        const int c = lane % W, lsel = lane / W;
        const int nstrip = nchunk;
        const int ng = min(kGP, ngpt - g0);
        const int S = m.nplanck;
        const bool has_lev = a.tlev != nullptr;
        const int nlevu = has_lev ? nlay + 1 : nlay;  // levels 0..nlay carry a value when tlev is present
        const int ups = (nlevu + LPU - 1) / LPU;      // units per strip
        const int nunit = nstrip * ups;
        // unit u -> level j of this lane and its (clamped) column
        auto place = [&](int u, int &j, long long &colp, bool &c_ok) {
            const int st = u / ups;
            j = (u - st * ups) * LPU + lsel;
            const long long cc = (long long)((int)blockIdx.x + st * (int)gridDim.x) * W + c;
            c_ok = cc < ncol;
            colp = min(cc, (long long)ncol - 1);
        };
        // software pipeline: the temperatures of the next unit are fetched before the g-loop of the current one
        auto fetch = [&](int u, double &tl, double &tv) {
            int j; long long colp; bool c_ok;
            place(u, j, colp, c_ok);
            tl = (double)a.tlay[colp + (long long)ncol * min(j, nlay - 1)];
            tv = 0.0;
            if (has_lev) tv = (double)a.tlev[colp + (long long)ncol * min(j, nlay)];
        };
        double tl_n, tv_n;
        if (pw < nunit) fetch(pw, tl_n, tv_n);
        const unsigned plane32 = (unsigned)((long long)ncol * nlay);   // elements between g-points (the walking kernel needs ncol*nlay*ngpt < 2^31)
        const Real *o_lay = a.out[1], *o_inc = a.out[2], *o_dec = a.out[3];
        keep(o_lay); keep(o_inc); keep(o_dec);
        const float2 *spl2 = reinterpret_cast<const float2 *>(spl);
        for (int u = pw; u < nunit; u += NPW) {
            int j; long long colp; bool c_ok;
            place(u, j, colp, c_ok);
            const bool ok_lay = c_ok && j < nlay;                       // lay_source(:,j), lev_source_dec(:,j)
            const bool ok_inc = c_ok && has_lev && j >= 1 && j <= nlay; // lev_source_inc(:,j-1)
            const double tl = tl_n, tv = tv_n;
            if (u + NPW < nunit) fetch(u + NPW, tl_n, tv_n);
            double a0, a1;
            int it = planck_index(m, tl, a0, a1);
            const float2 *tl_p = spl2 + (max(it, 1) - 1);
            const Real l0 = (Real)div_const(a0, m.pi, m.r_pi), l1 = (Real)div_const(a1, m.pi, m.r_pi);
            unsigned off = (unsigned)colp + (unsigned)ncol * (unsigned)j + plane32 * (unsigned)g0;
            int so = 0;
            if (has_lev) {
                it = planck_index(m, tv, a0, a1);
                const float2 *tv_p = spl2 + (max(it, 1) - 1);
                const Real v0 = (Real)div_const(a0, m.pi, m.r_pi), v1 = (Real)div_const(a1, m.pi, m.r_pi);
#pragma unroll 4
                for (int gq = 0; gq < ng; ++gq, off += plane32, so += S) {
                    const float2 pl = tl_p[so], pv = tv_p[so];
                    const Real vl = fma_(l1, (Real)pl.y, mul_(l0, (Real)pl.x));
                    const Real vv = fma_(v1, (Real)pv.y, mul_(v0, (Real)pv.x));      // B(tlev(:,j))
                    stcs_if(const_cast<Real *>(o_lay) + off, vl, ok_lay);
                    stcs_if(const_cast<Real *>(o_dec) + off, vv, ok_lay);                        // lev_source_dec(:,j)   :424
                    stcs_if(const_cast<Real *>(o_inc) + (off - (unsigned)ncol), vv, ok_inc);      // lev_source_inc(:,j-1) :423
                }
            } else {
#pragma unroll 4
                for (int gq = 0; gq < ng; ++gq, off += plane32, so += S) {
                    const float2 pl = tl_p[so];
                    const Real vl = fma_(l1, (Real)pl.y, mul_(l0, (Real)pl.x));
                    stcs_if(const_cast<Real *>(o_lay) + off, vl, ok_lay);
                }
            }
        }
        if (a.out2d) {                                 // sfc_source(:, g) = B(tsfc)   :408-412
            for (int st = pw; st < nstrip; st += NPW) {
                const long long cc = (long long)((int)blockIdx.x + st * (int)gridDim.x) * W + c;
                const long long colp = min(cc, (long long)ncol - 1);
                double a0, a1;
                const int it = planck_index(m, (double)a.tsfc[colp], a0, a1);
                const int r = max(it, 1) - 1;
                const Real s0 = (Real)div_const(a0, m.pi, m.r_pi), s1 = (Real)div_const(a1, m.pi, m.r_pi);
                for (int gq = lsel; gq < ng; gq += LPU)
                    if (cc < ncol) {
                        const float2 ps = spl2[gq * S + r];
                        a.out2d[colp + (long long)ncol * (g0 + gq)] = fma_(s1, (Real)ps.y, mul_(s0, (Real)ps.x));
                    }
            }
        }
        return;
    }

    return;                      // idle warps of a padded helper warpgroup
    }

    // =========================== index warps =====================================
    if (warp >= NCW && warp < NCW + NI) {
        if constexpr (RC > 0) reg_dec<RI>();
        const int iw = warp - NCW;
        int tr = 0;                                        // running round over all tiles of the block
        for (int ts = 0; ts < nslot; ++ts) {
          const int gt = tile_of(ts);
          if (gt >= ntile_all) continue;
          const long long col0 = (long long)gt * TC;
          for (int r = 0; r < nround; ++r, ++tr) {
            const int rb = tr & 1, u = tr >> 1;
            const int j = r * 32 + lane;
            for (int cs = iw; cs < TC; cs += NI) {
                const long long col = min(col0 + cs, (long long)ncol - 1);   // tail tiles recompute the last column; never stored
                RecT rec;
                if (j < nlay && !(ECCKD_DBG(2) && tr > 0)) {
                    const double pu = (double)a.plev[col + (long long)ncol * j], pl = (double)a.plev[col + (long long)ncol * (j + 1)];
                    const double tl = (double)a.tlay[col + (long long)ncol * j];
                    const CellIndex ci = cell_index(m, pu, pl, tl);
                    const double pw0 = __dsub_rn(1.0, ci.pw1), tw0 = __dsub_rn(1.0, ci.tw1);
                    rec.c[0] = (Real)__dmul_rn(tw0, pw0); rec.c[1] = (Real)__dmul_rn(tw0, ci.pw1);
                    rec.c[2] = (Real)__dmul_rn(ci.tw1, pw0); rec.c[3] = (Real)__dmul_rn(ci.tw1, ci.pw1);
                    const int ip = ci.ip0 - 1, it = ci.it0 - 1;
                    // Tables are >= 0 (checked at model create), so od < 0 <=> weight < 0: the per-g clamp
                    // (:233-238) becomes "skip the gas when its weight is negative"; NaN weights propagate.
                    int mask = 0;
                    rec.w[0] = (Real)ci.sw;
                    if (has_mix) mask |= (ci.sw < 0.0) ? 0x200 : 1;
#pragma unroll
                    for (int s = 1; s < NB; ++s) {
                        Real wv = 0;
                        if (s <= nvar) {
                            const GasRef &gr = cg.g[cg.var_gas[s - 1]];
                            const double wd = gas_weight(gr, ci.sw, load_vmr<Real>(gr, col, j));
                            wv = (Real)wd;
                            if (!(wd < 0.0)) mask |= 1 << s;
                        }
                        rec.w[s] = wv;
                    }
                    rec.wl = 0; rec.vw1 = 0; rec.hkoff = 0;
                    if (has_lut) {
                        const GasRef &gr = cg.g[cg.lut];
                        const double vmr = load_vmr<Real>(gr, col, j);
                        double v1;
                        const int iv = vmr_index(gr, vmr, v1) - 1;
                        rec.hkoff = ((iv * kNPP + ip) * kNTP + it) * kGP;
                        rec.vw1 = (Real)v1;
                        const double wd = gas_weight(gr, ci.sw, vmr);
                        rec.wl = (Real)wd;
                        if (!(wd < 0.0)) mask |= 0x100;
                    }
                    rec.key = (ip * kNTP + it) * kGP | mask << kMaskShift;
                    rec.pad0 = rec.pad1 = 0;
                }
                if (u >= 1) mbar_wait_relaxed(&rec_empty[rb * TC + cs], (u - 1) & 1, kIndexNap);
                if (j < nlay && !(ECCKD_DBG(2) && tr > 1)) recs[(rb * TC + cs) * 32 + lane] = rec;
                __syncwarp();
                if (lane == 0) mbar_arrive(&rec_full[rb * TC + cs]);
            }
          }
        }
        return;
    }

    // =========================== compute warps ===================================
    if constexpr (RC > 0) reg_inc<RC>();
    // warp w owns columns CPW*w .. CPW*w + CPW-1 of the tile (CPW = 2: two independent dependency chains per
    // lane, and the pair of values goes to the staging tile as one conflict-free 16-byte store)
    const int g = min(g0 + lane, ngpt - 1);     // inactive lanes shadow the last g-point; never stored
    const int ngroups = gridDim.y;
    // The file tables are exact in fp32.  The varying gases are cached as fp32 rows and widened in the FMA chain; the
    // LUT gas (eight values per layer) is cached at full width, because the conversion pipe (16 lanes/clk/SM) is
    // shared with the Planck warps and eight more F2F per layer measurably slow both; the per-call mixed table, a
    // sum of products, needs the full Real width anyway.
    const float *wbil = m.wbil32 + (size_t)grp * m.nslot * kSlab + walk_pos32(lane);
    const Real *wlut = TableOf<Real>::wlut(m) + (size_t)grp * m.nv_lut * kSlab + walk_pos<Real>(lane);
    const Real *wmix = mixtab + (size_t)grp * kSlab + walk_pos<Real>(lane);          // negative-sign copy at + ngroups*kSlab
    const float *wvar = wbil + (NB == 2 ? cg.var_off[0] : 0);       // NB == 2: the one varying gas
    if constexpr ((RF & 2) != 0) { keep(wvar); keep(wlut); keep(wmix); }

    Real km[CPW][4];                               // slot 0: mixed table
    float kb[CPW][NB > 1 ? NB - 1 : 1][4];         // slots 1..NB-1: varying gases
    Real kh[CPW][8];                               // LUT gas rows at full width: keeps eight F2F per layer off the conversion pipe
    int cur_pk[CPW], cur_hk[CPW];
    const Real sigma = SW ? (Real)m.rayleigh[g] : (Real)0;
#pragma unroll
    for (int q = 0; q < CPW; ++q) {
        km[q][0] = km[q][1] = km[q][2] = km[q][3] = 0;
#pragma unroll
        for (int s = 0; s < NB - 1; ++s) kb[q][s][0] = kb[q][s][1] = kb[q][s][2] = kb[q][s][3] = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) kh[q][i] = 0;
        cur_pk[q] = cur_hk[q] = -0x10000;
    }

    // (re)load the register cache of column q for corner offsets (pkoff, hkoff); no-ops when unchanged.
    // All addresses are formed before the first load issues, so that the look-up-table rows never queue behind the
    // scoreboard of the bilinear rows (measured: a false dependency through a reused address register cost one L2
    // round trip per pressure-row change).  RF & 1: branch-free (every load predicated).
    auto refill = [&](int q, int pkoff, int hkoff) {
        const bool changed = pkoff != cur_pk[q];
        const bool shift = pkoff == cur_pk[q] + kNTP * kGP;      // ip0 advanced by one, it0 unchanged
        const bool full = changed && !shift;
        const bool hch = has_lut && hkoff != cur_hk[q];
        const Real *kmp = wmix + pkoff;
        const Real *khp = wlut + hkoff;
        keep(kmp); keep(khp);
        if ((RF & 1) || changed) {
            if (shift) { km[q][0] = km[q][1]; km[q][2] = km[q][3]; }
            ldg_if(km[q][0], kmp + kC00, full && has_mix); ldg_if(km[q][2], kmp + kC01, full && has_mix);
            ldg_if(km[q][1], kmp + kC10, changed && has_mix); ldg_if(km[q][3], kmp + kC11, changed && has_mix);
#pragma unroll
            for (int s = 0; s < NB - 1; ++s) {
                const bool on = s < nvar;
                const float *k = (NB == 2 ? wvar : wbil + cg.var_off[s]) + pkoff;
                if (shift) { kb[q][s][0] = kb[q][s][1]; kb[q][s][2] = kb[q][s][3]; }
                ldg_if(kb[q][s][0], k + kC00, full && on); ldg_if(kb[q][s][2], k + kC01, full && on);
                ldg_if(kb[q][s][1], k + kC10, changed && on); ldg_if(kb[q][s][3], k + kC11, changed && on);
            }
        }
        if ((RF & 1) || hch) {
            // the same one-row shift for the look-up-table gas (its key moves with ip0 too): four loads instead of eight
            const bool hshift = hkoff == cur_hk[q] + kNTP * kGP;
            const bool hfull = hch && !hshift;
            if (hshift) { kh[q][0] = kh[q][1]; kh[q][2] = kh[q][3]; kh[q][4] = kh[q][5]; kh[q][6] = kh[q][7]; }
            ldg_if(kh[q][0], khp + kC00, hfull); ldg_if(kh[q][2], khp + kC01, hfull);
            ldg_if(kh[q][4], khp + kSlab + kC00, hfull); ldg_if(kh[q][6], khp + kSlab + kC01, hfull);
            ldg_if(kh[q][1], khp + kC10, hch); ldg_if(kh[q][3], khp + kC11, hch);
            ldg_if(kh[q][5], khp + kSlab + kC10, hch); ldg_if(kh[q][7], khp + kSlab + kC11, hch);
        }
        cur_pk[q] = pkoff;
        cur_hk[q] = hkoff;
    };

    // running staging pointer of this thread: + b*TILE + ls*ROWS.  TMA: row = g-point (lane), 16-byte chunk
    // XOR-swizzled with (row & 7) exactly as CU_TENSOR_MAP_SWIZZLE_128B lays the box out
    Real *st;
    if (TMA) st = CPW == 2 ? stage + lane * TC + ((warp ^ (lane & 7)) << 1)
                           : stage + lane * TC + ((((warp >> 1) ^ (lane & 7)) << 1) | (warp & 1));
    else st = stage + lane * TCP + warp;
    int ls = 0, b = 0, fill = 0;               // layer inside the tile, ring slot, fills of slot 0 so far
    unsigned a_full = smem_u32(bar_full), a_empty = smem_u32(bar_empty);   // shared-window addresses, kept in registers
    int lane_k = lane;
    if constexpr ((RF & 2) != 0) { keep(a_full); keep(a_empty); keep(lane_k); }
    int tr = 0;                                        // running round over all tiles of the block; the register
    for (int ts = 0; ts < nslot; ++ts) {               // cache, the staging ring and the barrier phases carry over
      if (tile_of(ts) >= ntile_all) continue;
      for (int r = 0; r < nround; ++r, ++tr) {
        const int rb = tr & 1;
        const int len = min(32, nlay - r * 32);
        const RecT *rp[CPW];
#pragma unroll
        for (int q = 0; q < CPW; ++q) {
            mbar_wait(&rec_full[rb * TC + warp * CPW + q], (tr >> 1) & 1);
            rp[q] = recs + (rb * TC + warp * CPW + q) * 32;
            const int2 k0 = *reinterpret_cast<const int2 *>(&rp[q]->key);
            refill(q, k0.x & 0xffff, k0.y);
        }
        for (int l = 0; l < len; ++l) {
            // tile free again?  Probe now, look at the answer only before the staging store: the barrier round
            // trip hides under the interpolation arithmetic
            // (parity (fill & 1) ^ 1: the first pass over the ring asks for the phase before the initial one, which
            //  a fresh barrier reports as complete)
            const bool need_tile = ls == 0;
            const bool tile_ok = !need_tile || ((RF & 4) && mbar_test_a(a_empty + b * 8, (fill & 1) ^ 1));
            Real val[NOUT][CPW];
            int2 k1[CPW];
            constexpr bool early = !(RF & 8);          // measured: LW prefers the early fetch, SW the late one
#pragma unroll
            for (int q = 0; q < CPW; ++q) {
                const RecT *rq = rp[q];
                const int4 ix = *reinterpret_cast<const int4 *>(&rq->key);
                Real c0, c1, c2, c3;
                lds4<Real>(rq->c, c0, c1, c2, c3);
                const int key = ix.x;
                k1[q] = make_int2(cur_pk[q], cur_hk[q]);
                if (l + 1 < len) k1[q] = *reinterpret_cast<const int2 *>(&rq[1].key);
                constexpr bool asap = (RF & 16) != 0;     // refill as soon as the bilinears have read the cache
                Real tau = 0;
                constexpr int kCommon = kMaskLut | (3 << kMaskShift);   // LUT gas + mixed table + one varying gas, all contributing
                if (NB == 2 && (key & (kCommon | kMaskNeg)) == kCommon) {
                    // the overwhelmingly common case as straight-line code: four independent bilinear chains
                    Real wl, v1, w0, w1;
                    lds2<Real>(&rq->wl, wl, v1);
                    lds2<Real>(&rq->w[0], w0, w1);
                    const Real b0 = bil4(c0, c1, c2, c3, kh[q][0], kh[q][1], kh[q][2], kh[q][3]);
                    const Real b1 = bil4(c0, c1, c2, c3, kh[q][4], kh[q][5], kh[q][6], kh[q][7]);
                    const Real bm = bil4(c0, c1, c2, c3, km[q][0], km[q][1], km[q][2], km[q][3]);
                    const Real bv = bil4(c0, c1, c2, c3, kb[q][0][0], kb[q][0][1], kb[q][0][2], kb[q][0][3]);
                    if (asap && !ECCKD_DBG(8)) refill(q, k1[q].x & 0xffff, k1[q].y);
                    tau = mul_(wl, fma_(v1, b1, mul_(add_((Real)1, -v1), b0)));
                    tau = fma_(w0, bm, tau);
                    tau = fma_(w1, bv, tau);
                } else {
                    if (key & kMaskLut) {
                        Real wl, v1;
                        lds2<Real>(&rq->wl, wl, v1);
                        const Real b0 = bil4(c0, c1, c2, c3, kh[q][0], kh[q][1], kh[q][2], kh[q][3]);
                        const Real b1 = bil4(c0, c1, c2, c3, kh[q][4], kh[q][5], kh[q][6], kh[q][7]);
                        // (1-v1)*b0 + v1*b1, not b0 + v1*(b1-b0): that cancels when b1 << b0
                        tau = mul_(wl, fma_(v1, b1, mul_(add_((Real)1, -v1), b0)));
                    }
#pragma unroll
                    for (int s = 0; s < NB; s += 2) {
                        if (key & (3 << (s + kMaskShift))) {
                            Real w0, w1;
                            lds2<Real>(&rq->w[s], w0, w1);
                            if (key & (1 << (s + kMaskShift))) {
                                if (s == 0) tau = fma_(w0, bil4(c0, c1, c2, c3, km[q][0], km[q][1], km[q][2], km[q][3]), tau);
                                else tau = fma_(w0, bil4(c0, c1, c2, c3, kb[q][s > 0 ? s - 1 : 0][0], kb[q][s > 0 ? s - 1 : 0][1],
                                                         kb[q][s > 0 ? s - 1 : 0][2], kb[q][s > 0 ? s - 1 : 0][3]), tau);
                            }
                            if (key & (2 << (s + kMaskShift)))
                                tau = fma_(w1, bil4(c0, c1, c2, c3, kb[q][s][0], kb[q][s][1], kb[q][s][2], kb[q][s][3]), tau);
                        }
                    }
                    if (key & kMaskNeg) {   // upside-down layer: the negative-factor gases give positive optical depth
                        const Real *k = wmix + (size_t)ngroups * kSlab + (key & 0xffff);
                        tau = fma_(rq->w[0], bil4(c0, c1, c2, c3, k[kC00], k[kC10], k[kC01], k[kC11]), tau);
                    }
                    if (asap && !ECCKD_DBG(8)) refill(q, k1[q].x & 0xffff, k1[q].y);
                }
                if (early && !asap && !ECCKD_DBG(8)) refill(q, k1[q].x & 0xffff, k1[q].y);   // cache is free: fetch the next layer's corners under the Planck / staging work
                if (SW) {
                    const Real ray = mul_(rq->w[0], sigma);         // :313-316 (moles_per_layer == simple_weight, same product)
                    tau = add_(tau, ray);                           // :456
                    val[0][q] = tau;
                    val[1][q] = div_fast(ray, tau);                 // :459
                } else {
                    val[0][q] = tau;
                }
                rp[q] = rq + 1;
            }
            if (!tile_ok) mbar_wait_a(a_empty + b * 8, (fill & 1) ^ 1);
#pragma unroll
            for (int o = 0; o < NOUT; ++o) {
                if constexpr (CPW == 2) {
                    if constexpr (sizeof(Real) == 8) *reinterpret_cast<double2 *>(st + o * L * ROWS) = make_double2((double)val[o][0], (double)val[o][1]);
                    else *reinterpret_cast<float2 *>(st + o * L * ROWS) = make_float2((float)val[o][0], (float)val[o][1]);
                } else {
                    st[o * L * ROWS] = val[o][0];
                }
            }
            st += ROWS;
            if (++ls == L || (r == nround - 1 && l == len - 1)) {
                if (TMA) fence_proxy_async_smem();     // generic-proxy writes -> visible to the async (TMA) proxy
                __syncwarp();
                if (lane_k == 0) mbar_arrive_a(a_full + b * 8);
                // TMA path: the proxy fence drains this thread's outstanding loads, so the next layer's corner
                // fetch is issued after it (it then overlaps the next record read and Planck arithmetic)
                if (!early && !(RF & 16) && !ECCKD_DBG(8)) {
#pragma unroll
                    for (int q = 0; q < CPW; ++q) refill(q, k1[q].x & 0xffff, k1[q].y);
                }
                st += TILE - ls * ROWS;
                ls = 0;
                if (++b == NBUF) { b = 0; st -= NBUF * TILE; ++fill; }
            }
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < CPW; ++q) mbar_arrive(&rec_empty[rb * TC + warp * CPW + q]);
        }
      }
    }

    // ---- per-column 2-D output (SW; the LW surface source is written by the Planck warps) ----------
    if (SW && a.out2d && g0 + lane < ngpt) {
        for (int ts = 0; ts < nslot; ++ts) {
#pragma unroll
            for (int q = 0; q < CPW; ++q) {
                const long long cq = (long long)tile_of(ts) * TC + warp * CPW + q;
                if (cq < ncol) a.out2d[cq + (long long)ncol * g] = (Real)m.solar[g];                       // :468-472
            }
        }
    }
}

// ============================================================================
// debug kernels: the indices cell_index / vmr_index / planck_index produce
// ============================================================================
template <typename Real>
__global__ void indices_kernel(const __grid_constant__ DevModel m, const __grid_constant__ GasRef lut, int has_lut,
                               int ncol_, int nlay_, const Real *plev, const Real *tlay,
                               int *ip0, double *pw1, int *it0, double *tw1, int *iv0, double *vw1)
{
    const long long ncol = ncol_, ncell = ncol * nlay_;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncell) return;
    const long long col = idx % ncol, lay = idx / ncol;
    const CellIndex ci = cell_index(m, (double)plev[col + ncol * lay], (double)plev[col + ncol * (lay + 1)],
                                    (double)tlay[col + ncol * lay]);
    ip0[idx] = ci.ip0; pw1[idx] = ci.pw1; it0[idx] = ci.it0; tw1[idx] = ci.tw1;
    int iv = 1; double v1 = 0.0;
    if (has_lut) iv = vmr_index(lut, load_vmr<Real>(lut, col, lay), v1);
    iv0[idx] = iv; vw1[idx] = v1;
}

__global__ void planck_indices_kernel(const __grid_constant__ DevModel m, long long n, const double *T, int *it0, double *w1)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    double a0, a1;
    const int it = planck_index(m, T[idx], a0, a1);
    it0[idx] = it; w1[idx] = it > 0 ? a1 : 0.0;
}

// div_const against the IEEE division it replaces: number of dividends whose quotients differ in any bit
__global__ void div_const_check_kernel(long long n, const double *x, double c, double rc, unsigned long long *mismatches)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const double a = div_const(x[idx], c, rc), b = __ddiv_rn(x[idx], c);
    const bool same = (a != a && b != b) || __double_as_longlong(a) == __double_as_longlong(b) || (a == 0.0 && b == 0.0);   // (+0 / -0: see div_const)
    if (!same) atomicAdd(mismatches, 1ull);
}

cudaError_t launch_div_const_check(long long n, const double *x, double c, double rc, unsigned long long *mismatches, cudaStream_t s)
{
    div_const_check_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, x, c, rc, mismatches);
    return cudaGetLastError();
}

// ============================================================================
// launchers
// ============================================================================
// cuTensorMapEncodeTiled through the runtime's driver-entry-point lookup (no link-time dependency on libcuda)
PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }();
    return fn;
}

// 3-D map of an (ncol, nlay, ngpt) fp64 output, box = 16 columns x 1 layer x 32 g-points, 128-byte swizzle
static bool encode_output_map(CUtensorMap *map, double *ptr, int ncol, int nlay, int ngpt)
{
    auto enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)ncol, (cuuint64_t)nlay, (cuuint64_t)ngpt};
    const cuuint64_t gstr[2] = {(cuuint64_t)ncol * 8, (cuuint64_t)ncol * nlay * 8};
    const cuuint32_t box[3] = {16, 1, 32}, estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Launch geometry of the persistent walk kernel: tiles per chunk and blocks per g-group.  One block is resident
// per SM (shared memory).  Two-tile chunks pay in the longwave only (wider Planck stores, measured 8 %), and only
// when the coarser granularity does not unbalance the blocks: compare the longest block under both choices.
// Pure host arithmetic (exported for the CPU tests as ecckd_debug_walk_plan).
void walk_plan(int ntile, bool sw, int sms, int &tiles_per_chunk, int &blocks)
{
    if (sms < 1) sms = 1;
    tiles_per_chunk = 1;
    if (!sw && ntile >= 4 * sms) {
        const int n2 = (ntile + 1) / 2;
        const double t1 = 1.08 * ((ntile + sms - 1) / sms), t2 = 2.0 * ((n2 + sms - 1) / sms);
        if (t2 <= t1) tiles_per_chunk = 2;
    }
    const int nchunk = (ntile + tiles_per_chunk - 1) / tiles_per_chunk;
    blocks = nchunk < sms ? nchunk : sms;
    if (blocks < 1) blocks = 1;
}

static void walk_geometry(int ntile, bool sw, int sms, int &tiles_per_chunk, int &blocks)
{
    walk_plan(ntile, sw, sms, tiles_per_chunk, blocks);
#ifdef ECCKD_TUNE
    if (getenv("ECCKD_CH") || getenv("ECCKD_BLOCKS")) {
        if (getenv("ECCKD_CH")) tiles_per_chunk = atoi(getenv("ECCKD_CH"));
        const int cap = getenv("ECCKD_BLOCKS") ? atoi(getenv("ECCKD_BLOCKS")) : sms;
        const int nchunk = (ntile + tiles_per_chunk - 1) / tiles_per_chunk;
        blocks = nchunk < cap ? nchunk : cap;
    }
#endif
}

template <typename Real, bool SW, int NB, int TC, int L, int NBUF, int NI, int NWW, bool TMA = false, int CPW = 1, int NPW_LW = 8, int NPW_SW = 0,
          int RF = 7, int RC = 0, int RI = 64, int RH = 56>
static cudaError_t launch_walk_t(const DevModel &m, const CallGases &cg, const KernelArgs<Real> &a, const Real *mix, cudaStream_t s,
                                 const TmaMaps *maps = nullptr)
{
    constexpr int NOUT = SW ? 2 : 1;
    constexpr int NPW = SW ? NPW_SW : NPW_LW;  // helper warps: Planck sources (LW) / zero g-plane (SW)
    constexpr int TCP = TMA ? TC : TC + 1;
    const size_t smem = sizeof(Rec<Real, NB>) * 2 * TC * 32 + sizeof(Real) * NBUF * NOUT * L * 32 * TCP +
                        ((TMA && SW) ? sizeof(Real) * 32 * TCP : 0) +
                        (2 * NBUF + 4 * TC) * sizeof(unsigned long long) + (SW ? 0 : sizeof(float2) * m.nplanck * kGP);
    auto kern = walk_kernel<Real, SW, NB, TC, L, NBUF, NI, NWW, TMA, CPW, NPW, RF, RC, RI, RH>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int ntile = (a.ncol + TC - 1) / TC;
    int tpc, blocks;
    walk_geometry(ntile, SW, m.sm_count, tpc, blocks);
    dim3 grid((unsigned)blocks, (unsigned)((m.ngpt + kGP - 1) / kGP));
    static const TmaMaps no_maps{};
    kern<<<grid, (TC / CPW + NI + NWW + NPW + PadWarps<NWW, NPW, RC>::value) * 32, smem, s>>>(m, cg, a, mix, maps ? *maps : no_maps, tpc);
    return cudaGetLastError();
}

// The TMA store path needs fp64, 16-byte aligned outputs and a column pitch that is a multiple of 16 bytes.
template <typename Real> static bool tma_eligible(const KernelArgs<Real> &a)
{
    if (sizeof(Real) != 8 || (a.ncol & 1)) return false;
    for (int o = 0; o < 4; ++o) if (a.out[o] && (reinterpret_cast<uintptr_t>(a.out[o]) & 15)) return false;
    return tensor_map_encoder() != nullptr;
}

size_t walk_scratch_bytes(const DevModel &m, const CallGases &, size_t real_size)
{
    return (size_t)2 * ((m.ngpt + kGP - 1) / kGP) * kSlab * real_size;
}

template <typename Real, bool SW>
cudaError_t launch_walk(const DevModel &m, const CallGases &cg, const KernelArgs<Real> &a, void *scratch, cudaStream_t s, int which)
{
    if (which == 1 && !quad_eligible<Real>(cg, a)) return cudaErrorInvalidValue;
    Real *mix = static_cast<Real *>(scratch);
    if (cg.nmix > 0) {
        const int n = ((m.ngpt + kGP - 1) / kGP) * kSlab;
        mix_kernel<Real><<<(n + 255) / 256, 256, 0, s>>>(m, cg, mix);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    {   // the product path: four columns per compute warp, TMA tensor stores (ecckd_quad.cu)
        static const bool no_quad = getenv("ECCKD_NO_QUAD") != nullptr;
        if (which == 1 || (which == 0 && !no_quad && quad_eligible<Real>(cg, a))) return launch_quad<Real, SW>(m, cg, a, mix, m.sm_count, s);
    }
    if (cg.nvar <= 1) {
        if constexpr (sizeof(Real) == 8) {
            static const bool no_tma = getenv("ECCKD_NO_TMA") != nullptr;
            if (!no_tma && tma_eligible(a)) {
                // staging tiles leave the SM through cp.async.bulk.tensor: no writer warps, one tile per layer
                TmaMaps maps{};
                bool ok = true;
                for (int o = 0; o < (SW ? 3 : 1); ++o)
                    if (a.out[o]) ok = ok && encode_output_map(&maps.m[o], reinterpret_cast<double *>(a.out[o]), a.ncol, a.nlay, m.ngpt);
                if (ok) {
#ifdef ECCKD_TUNE_WALK1
                    // tuning build: ECCKD_WALK_VARIANT selects an alternative instantiation (round-1 tuning sweeps; build with -DECCKD_TUNE_WALK1)
                    //   100 + RF      refill flavour RF at the default geometry
                    //   LW 1..5       2 / 4 / 6 / 8 Planck warps, 3 index warps      LW 6..8   ring of 4 / 8 / 16 tiles
                    //   SW 1..2       ring of 4 / 16 tiles      SW 3      3 index warps (20 warps: 96 registers per thread)
                    static const int v = getenv("ECCKD_WALK_VARIANT") ? atoi(getenv("ECCKD_WALK_VARIANT")) : 0;
#define ECCKD_VARIANT(n, ...) if (v == (n)) return launch_walk_t<Real, SW, __VA_ARGS__>(m, cg, a, mix, s, &maps);
#define ECCKD_RF_CASE(rf) \
                    if constexpr (SW) { ECCKD_VARIANT(100 + rf, 2, 16, 1, 8, 4, 1, true, 1, 8, 0, rf) } \
                    else { ECCKD_VARIANT(100 + rf, 2, 16, 1, 12, 4, 1, true, 1, 3, 0, rf) }
                    ECCKD_RF_CASE(6) ECCKD_RF_CASE(7) ECCKD_RF_CASE(14) ECCKD_RF_CASE(18) ECCKD_RF_CASE(22) ECCKD_RF_CASE(23)
                    if constexpr (SW) {
                        ECCKD_VARIANT(1, 2, 16, 1, 4, 4, 1, true, 1, 8, 0, 14)
                        ECCKD_VARIANT(2, 2, 16, 1, 16, 4, 1, true, 1, 8, 0, 14)
                        ECCKD_VARIANT(3, 2, 16, 1, 8, 3, 1, true, 1, 8, 0, 14)
                        ECCKD_VARIANT(31, 2, 16, 1, 8, 4, 1, true, 2, 8, 0, 14, 168, 64, 40)
                        ECCKD_VARIANT(32, 2, 16, 1, 8, 4, 1, true, 2, 8, 0, 6, 168, 64, 40)
                        ECCKD_VARIANT(33, 2, 16, 1, 8, 4, 1, true, 2, 8, 0, 22, 168, 64, 40)
                        ECCKD_VARIANT(34, 2, 16, 1, 8, 4, 1, true, 2, 8, 0, 18, 168, 64, 40)
                    } else {
                        ECCKD_VARIANT(1, 2, 16, 1, 12, 4, 1, true, 1, 2, 0, 18)
                        ECCKD_VARIANT(2, 2, 16, 1, 12, 4, 1, true, 1, 4, 0, 18)
                        ECCKD_VARIANT(3, 2, 16, 1, 12, 4, 1, true, 1, 6, 0, 18)
                        ECCKD_VARIANT(4, 2, 16, 1, 12, 4, 1, true, 1, 8, 0, 18)
                        ECCKD_VARIANT(5, 2, 16, 1, 12, 3, 1, true, 1, 3, 0, 18)
                        ECCKD_VARIANT(6, 2, 16, 1, 4, 4, 1, true, 1, 3, 0, 18)
                        ECCKD_VARIANT(7, 2, 16, 1, 8, 4, 1, true, 1, 3, 0, 18)
                        ECCKD_VARIANT(8, 2, 16, 1, 16, 4, 1, true, 1, 3, 0, 18)
                        ECCKD_VARIANT(31, 2, 16, 1, 12, 4, 1, true, 2, 3, 0, 18, 168, 64, 56)
                        ECCKD_VARIANT(32, 2, 16, 1, 12, 4, 1, true, 2, 3, 0, 6, 168, 64, 56)
                        ECCKD_VARIANT(33, 2, 16, 1, 12, 4, 1, true, 2, 3, 0, 22, 168, 64, 56)
                        ECCKD_VARIANT(34, 2, 16, 1, 12, 4, 1, true, 2, 3, 0, 14, 168, 64, 56)
                    }
#undef ECCKD_RF_CASE
#undef ECCKD_VARIANT
#endif
                    // one column per compute warp (16 warps): measured faster than two columns x 8 warps
                    // refill flavour (RF) and Planck-warp count as measured on B200: SW prefers the branchy, late refill
                    if constexpr (SW) return launch_walk_t<Real, SW, 2, 16, 1, 8, 4, 1, true, 1, 8, 0, 14>(m, cg, a, mix, s, &maps);
                    else return launch_walk_t<Real, SW, 2, 16, 1, 12, 4, 1, true, 1, 3, 0, 18>(m, cg, a, mix, s, &maps);
                }
            }
        }
#ifdef ECCKD_TUNE_WALK1
        if constexpr (!SW) {   // writer-warp path, longwave: 21..23 = 2 writer warps + 3 Planck warps with 2- / 4-layer tiles, 3 writer warps
            static const int v = getenv("ECCKD_WALK_VARIANT") ? atoi(getenv("ECCKD_WALK_VARIANT")) : 0;
            if (v == 21) return launch_walk_t<Real, SW, 2, 16, 2, 4, 4, 2, false, 1, 3, 0, 18>(m, cg, a, mix, s);
            if (v == 22) return launch_walk_t<Real, SW, 2, 16, 4, 2, 4, 2, false, 1, 3, 0, 18>(m, cg, a, mix, s);
            if (v == 23) return launch_walk_t<Real, SW, 2, 16, 2, 4, 4, 3, false, 1, 3, 0, 18>(m, cg, a, mix, s);
        }
#endif
        // writer-warp path (fp32, odd ncol, unaligned outputs)
        if constexpr (SW) return launch_walk_t<Real, SW, 2, 16, 2, 4, 4, 4, false, 1, 8, 0, 6>(m, cg, a, mix, s);
        else return launch_walk_t<Real, SW, 2, 16, 2, 4, 4, 4, false, 1, 4, 0, 18>(m, cg, a, mix, s);
    }
    return launch_walk_t<Real, SW, 8, 8, 2, 3, 4, 1, false, 1, 4>(m, cg, a, mix, s);
}

template <typename Real, bool SW>
cudaError_t launch_generic(const DevModel &m, const CallGases &cg, const KernelArgs<Real> &a, cudaStream_t s)
{
    const long long ncell = (long long)a.ncol * a.nlay;
    const unsigned blocks = (unsigned)((ncell + 255) / 256);
    generic_kernel<Real, SW><<<blocks, 256, 0, s>>>(m, cg, a);
    return cudaGetLastError();
}

template <typename Real>
cudaError_t launch_indices(const DevModel &m, const GasRef *lut, int ncol, int nlay, const Real *plev, const Real *tlay,
                           int *ip0, double *pw1, int *it0, double *tw1, int *iv0, double *vw1, cudaStream_t s)
{
    const long long ncell = (long long)ncol * nlay;
    GasRef dummy{};
    indices_kernel<Real><<<(unsigned)((ncell + 255) / 256), 256, 0, s>>>(m, lut ? *lut : dummy, lut != nullptr, ncol, nlay,
                                                                        plev, tlay, ip0, pw1, it0, tw1, iv0, vw1);
    return cudaGetLastError();
}

cudaError_t launch_planck_indices(const DevModel &m, long long n, const double *T, int *it0, double *w1, cudaStream_t s)
{
    planck_indices_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(m, n, T, it0, w1);
    return cudaGetLastError();
}

#define ECCKD_INST(Real)                                                                                               \
    template cudaError_t launch_walk<Real, false>(const DevModel &, const CallGases &, const KernelArgs<Real> &, void *, cudaStream_t, int); \
    template cudaError_t launch_walk<Real, true>(const DevModel &, const CallGases &, const KernelArgs<Real> &, void *, cudaStream_t, int);  \
    template cudaError_t launch_generic<Real, false>(const DevModel &, const CallGases &, const KernelArgs<Real> &, cudaStream_t); \
    template cudaError_t launch_generic<Real, true>(const DevModel &, const CallGases &, const KernelArgs<Real> &, cudaStream_t);  \
    template cudaError_t launch_indices<Real>(const DevModel &, const GasRef *, int, int, const Real *, const Real *, int *,   \
                                              double *, int *, double *, int *, double *, cudaStream_t);
ECCKD_INST(double)
ECCKD_INST(float)

}  // namespace ecckd
