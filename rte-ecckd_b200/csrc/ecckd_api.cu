// ecckd_api.cu -- implementation of the C ABI in include/ecckd_b200.h:
// table packing, per-call gas resolution (gas_optical_depth's orchestration,
// /root/reference/src/gas_optics_ecckd.f90:323-376), kernel selection, the
// host-array path (column shards over devices, pipelined column blocks) and the
// getters (:477-553).  No CPU compute path exists in this library.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "ck_loader.h"
#include "ecckd_b200.h"
#include "ecckd_device.cuh"
#include "ecckd_kernels.h"

using namespace ecckd;

namespace {

constexpr int kSlots = 3;   // pipeline depth of the host-array path

struct Slot {
    cudaStream_t stream = nullptr;
    unsigned char *arena = nullptr;
    size_t capacity = 0;
};

struct DevState {
    int device = 0;
    DevModel dm{};
    std::vector<void *> allocs;
    Slot slots[kSlots];
    void *rte_scratch = nullptr;      // grow-only work space of ecckd_rte_sw_f64 (direct beam, albedo, source of the adding method)
    size_t rte_scratch_bytes = 0;
};

// NVTX range for the profilers' timelines (H2D / table mix + walk / D2H); a no-op without an attached tool
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// RN(1/c) for div_const (ecckd_device.cuh), or NaN when the Markstein argument does not cover c: zero / non-finite
// divisors and significands that are all ones.
double const_reciprocal(double c)
{
    unsigned long long b;
    std::memcpy(&b, &c, 8);
    if (!(std::fabs(c) > 0.0) || !std::isfinite(c) || (b & 0xfffffffffffffull) == 0xfffffffffffffull) return std::nan("");
    const double r = 1.0 / c;
    return (std::isfinite(r) && std::fabs(r) > 2.3e-308) ? r : std::nan("");
}

void set_err(char *err, const std::string &msg)
{
    if (!err) return;
    std::snprintf(err, ECCKD_ERRLEN, "%s", msg.c_str());
}

std::string trim_name(const char *s, size_t maxlen)
{
    size_t n = 0;
    while (n < maxlen && s[n] != '\0') ++n;
    while (n > 0 && s[n - 1] == ' ') --n;
    return std::string(s, n);
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            set_err(err, std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #expr); \
            return ECCKD_ERR_CUDA;                                                          \
        }                                                                                   \
    } while (0)

}  // namespace

struct ecckd_model {
    bool shortwave = false;
    int ngpt = 0, np = 0, nt = 0, nplanck = 0, nband = 0;
    std::vector<std::string> gas;
    std::vector<int> composite_only, code, nv;
    std::vector<double> ref, mf0, log_mf0, dlog_mf;
    std::vector<long long> table_off;
    std::vector<int> slot;      // walk-table slot of each bilinear gas (-1: LUT gas)
    bool walk_model = false;    // walk tables were built (see ecckd_model_create)
    int nslot_ = 0, nv_lut_ = 0;
    std::vector<double> log_pressure, temperature, band_lims;
    std::vector<int32_t> band2gpt;
    double total_solar = 0.0;
    bool tables_nonneg = true, fp32_exact = true;
    std::string warnings;       // grids that are not uniform although the algorithm assumes so (see ecckd_get_table_warnings)
    std::vector<std::unique_ptr<DevState>> devs;

    ~ecckd_model()
    {
        for (auto &d : devs) {
            cudaSetDevice(d->device);
            for (Slot &s : d->slots) {
                if (s.arena) cudaFree(s.arena);
                if (s.stream) cudaStreamDestroy(s.stream);
            }
            for (void *p : d->allocs) cudaFree(p);
            if (d->rte_scratch) cudaFree(d->rte_scratch);
        }
    }
};

namespace {

template <typename T>
int upload(DevState &d, const std::vector<T> &h, const T **out, char *err)
{
    *out = nullptr;
    if (h.empty()) return ECCKD_OK;
    void *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, h.size() * sizeof(T)));
    d.allocs.push_back(p);
    CUDA_TRY(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<const T *>(p);
    return ECCKD_OK;
}

// gas_optical_depth's gas loop (:348-374) resolved on the host.
int resolve_gases(const ecckd_model *M, int ngas, const char *const *names, const ecckd_vmr_view *vmr,
                  CallGases &cg, bool &walk_ok, char *err)
{
    std::memset(&cg, 0, sizeof(cg));
    cg.lut = -1;
    int nlut = 0;
    bool first_calc = true, many = false;
    if (ngas < 0 || (ngas > 0 && (!names || !vmr))) { set_err(err, "ecckd: bad gas list"); return ECCKD_ERR_ARG; }
    for (int j = 0; j < ngas; ++j) {                                                   // :348
        const std::string want = trim_name(names[j], 4096);
        int i = -1;
        for (size_t q = 0; q < M->gas.size(); ++q)                                     // :349-357
            if (M->gas[q] == want) { i = (int)q; break; }
        if (i < 0) continue;                                                           // :358-364 unknown gas: skipped
        if (M->composite_only[i] && !first_calc) continue;                             // :365-367
        if (cg.n >= kMaxGases) { set_err(err, "ecckd: more than 16 active gases"); return ECCKD_ERR_ARG; }
        GasRef &g = cg.g[cg.n];
        g.code = M->code[i];
        if (g.code != kNone && !vmr[j].ptr) { set_err(err, "ecckd: NULL vmr for gas " + want); return ECCKD_ERR_ARG; }
        g.vmr = vmr[j].ptr; g.col_stride = vmr[j].col_stride; g.lay_stride = vmr[j].lay_stride;
        g.table_off = M->table_off[i];
        g.ref = M->ref[i]; g.nv = M->nv[i];
        g.mf0 = M->mf0[i]; g.log_mf0 = M->log_mf0[i]; g.dlog_mf = M->dlog_mf[i];
        g.v_clamp = (double)M->nv[i] - 1.001;                                          // :160
        g.r_dlog_mf = const_reciprocal(g.dlog_mf);
        if (g.code == kLut) { if (nlut++ == 0) cg.lut = cg.n; }
        else {
            ++cg.nbil;
            if (M->walk_model) {
                // call-constant weight factor (code 0, or a scalar concentration) -> mixed table; else a register slot
                const bool constant = g.code == kNone || (g.col_stride == 0 && g.lay_stride == 0);
                if (constant) {
                    if (cg.nmix < kMaxBil) { cg.mix_gas[cg.nmix] = cg.n; cg.mix_off[cg.nmix] = M->slot[i] * kSlab; ++cg.nmix; }
                    else many = true;
                } else {
                    if (cg.nvar < kMaxBil - 1) { cg.var_gas[cg.nvar] = cg.n; cg.var_off[cg.nvar] = M->slot[i] * kSlab; ++cg.nvar; }
                    else many = true;
                }
            }
        }
        ++cg.n;
        if (M->composite_only[i]) first_calc = false;                                  // :371-373
    }
    walk_ok = M->walk_model && nlut <= 1 && !many;
    return ECCKD_OK;
}

template <typename Real>
int run_device(ecckd_model *M, int dev_index, bool sw, int ncol, int nlay, const Real *plev, const Real *tlay,
               const Real *tlev, const Real *tsfc, int ngas, const char *const *names, const ecckd_vmr_view *vmr,
               Real *o0, Real *o1, Real *o2, Real *o3, Real *o2d, int kernel, cudaStream_t stream, char *err)
{
    if (err) err[0] = '\0';
    if (!M) { set_err(err, "ecckd: NULL model"); return ECCKD_ERR_ARG; }
    if (dev_index < 0 || dev_index >= (int)M->devs.size()) { set_err(err, "ecckd: bad device index"); return ECCKD_ERR_ARG; }
    if (sw != M->shortwave) {
        set_err(err, sw ? "ecckd: model has no solar source (longwave tables)" : "ecckd: model has no Planck source (shortwave tables)");
        return ECCKD_ERR_ARG;
    }
    if (ncol < 0 || nlay < 0) { set_err(err, "ecckd: negative extent"); return ECCKD_ERR_ARG; }
    if (!std::is_same<Real, double>::value && !M->fp32_exact) {
        set_err(err, "ecckd: fp32 entry points need tables that are exact in fp32");
        return ECCKD_ERR_ARG;
    }
    CallGases cg;
    bool walk_ok = false;
    int rc = resolve_gases(M, ngas, names, vmr, cg, walk_ok, err);
    if (rc != ECCKD_OK) return rc;
    const bool lev_missing = !sw && !tlev;
    const bool not_2str = sw && (!o1 || !o2);
    if (ncol > 0 && nlay > 0) {
        if (!plev || !tlay || !o0 || (!sw && (!tsfc || !o1 || !o2d)) || (!sw && tlev && (!o2 || !o3))) {
            set_err(err, "ecckd: NULL array argument");
            return ECCKD_ERR_ARG;
        }
        DevState &d = *M->devs[dev_index];
        CUDA_TRY(cudaSetDevice(d.device));
        KernelArgs<Real> a{};
        a.ncol = ncol; a.nlay = nlay;
#ifdef ECCKD_TUNE
        a.dbg = getenv("ECCKD_DBG") ? atoi(getenv("ECCKD_DBG")) : 0;
#endif
        a.plev = plev; a.tlay = tlay; a.tlev = tlev; a.tsfc = tsfc;
        a.out[0] = o0;
        if (sw) { a.out[1] = not_2str ? nullptr : o1; a.out[2] = not_2str ? nullptr : o2; a.out[3] = nullptr; a.out2d = not_2str ? nullptr : o2d; }
        else { a.out[1] = o1; a.out[2] = lev_missing ? nullptr : o2; a.out[3] = lev_missing ? nullptr : o3; a.out2d = o2d; }
        // the one-column-per-warp walk kernel indexes outputs with 32-bit element offsets; the quad kernel has no such limit
        const bool big = (long long)ncol * nlay * M->ngpt >= (1ll << 31);
        const bool quad_ok = walk_ok && quad_eligible<Real>(cg, a);
        const bool colwalk_ok = walk_ok && !big;
        bool use_walk;
        int which = 0;
        if (kernel == ECCKD_KERNEL_AUTO) { use_walk = quad_ok || colwalk_ok; which = quad_ok ? 1 : 2; }
        else if (kernel == ECCKD_KERNEL_WALK) {
            if (!quad_ok && !colwalk_ok) { set_err(err, "ecckd: call does not qualify for the walking kernel"); return ECCKD_ERR_ARG; }
            use_walk = true; which = quad_ok ? 1 : 2;
        } else if (kernel == ECCKD_KERNEL_QUAD) {
            if (!quad_ok) { set_err(err, "ecckd: call does not qualify for the quad kernel"); return ECCKD_ERR_ARG; }
            use_walk = true; which = 1;
        } else if (kernel == ECCKD_KERNEL_COLWALK) {
            if (!colwalk_ok) { set_err(err, "ecckd: call does not qualify for the walking kernel"); return ECCKD_ERR_ARG; }
            use_walk = true; which = 2;
        } else if (kernel == ECCKD_KERNEL_GENERIC) use_walk = false;
        else { set_err(err, "ecckd: unknown kernel selector"); return ECCKD_ERR_ARG; }
        cudaError_t e;
        NvtxRange nvtx(use_walk ? (which == 1 ? "ecckd: mix + quad kernel" : "ecckd: mix + column-walk kernel") : "ecckd: generic kernel");
        if (use_walk) {
            // stream-ordered scratch for the per-call mixed table: safe for concurrent calls on other streams
            void *scratch = nullptr;
            CUDA_TRY(cudaMallocAsync(&scratch, walk_scratch_bytes(d.dm, cg, sizeof(Real)), stream));
            e = sw ? launch_walk<Real, true>(d.dm, cg, a, scratch, stream, which) : launch_walk<Real, false>(d.dm, cg, a, scratch, stream, which);
            cudaError_t e2 = cudaFreeAsync(scratch, stream);
            CUDA_TRY(e);
            CUDA_TRY(e2);
        } else {
            e = sw ? launch_generic<Real, true>(d.dm, cg, a, stream) : launch_generic<Real, false>(d.dm, cg, a, stream);
            CUDA_TRY(e);
        }
    }
    if (lev_missing) { set_err(err, "tlev is required for ecckd"); return ECCKD_ERR_REFERENCE; }              // :414-417
    if (not_2str) { set_err(err, "shortwave must use ty_optical_props_2str"); return ECCKD_ERR_REFERENCE; }    // :461-463
    return ECCKD_OK;
}

// ---- host-array path --------------------------------------------------------
struct Carve {
    unsigned char *base;
    size_t off = 0;
    template <typename T> T *take(size_t n)
    {
        off = (off + 255) & ~size_t(255);
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

template <typename Real>
int run_host(ecckd_model *M, bool sw, int ncol, int nlay, const Real *plev, const Real *tlay, const Real *tlev,
             const Real *tsfc, int ngas, const char *const *names, const ecckd_vmr_view *vmr,
             Real *o0, Real *o1, Real *o2, Real *o3, Real *o2d, int block_cols, char *err)
{
    if (err) err[0] = '\0';
    if (!M) { set_err(err, "ecckd: NULL model"); return ECCKD_ERR_ARG; }
    if (ncol < 0 || nlay < 0) { set_err(err, "ecckd: negative extent"); return ECCKD_ERR_ARG; }
    if (ngas < 0 || ngas > ECCKD_MAX_GASES) {
        // more names than the model can hold: unknown ones would be skipped anyway, but the view table is bounded
        set_err(err, "ecckd: the concentration list of the host path holds 0..16 gases");
        return ECCKD_ERR_ARG;
    }
    if (ngas > 0 && (!names || !vmr)) { set_err(err, "ecckd: bad gas list"); return ECCKD_ERR_ARG; }
    for (int j = 0; j < ngas; ++j)
        if (!names[j]) { set_err(err, "ecckd: NULL gas name"); return ECCKD_ERR_ARG; }
    const int ndev = (int)M->devs.size();
    const int ngpt = M->ngpt;
    const bool lev_missing = !sw && !tlev;
    const bool not_2str = sw && (!o1 || !o2);
    for (int j = 0; j < ngas; ++j) {
        const ecckd_vmr_view &v = vmr[j];
        const bool ok = (v.col_stride == 0 && (v.lay_stride == 0 || v.lay_stride == 1)) ||
                        (v.col_stride == 1 && (v.lay_stride == 0 || v.lay_stride >= ncol));
        if (v.ptr && !ok) { set_err(err, "ecckd: host path takes scalar, (nlay), (ncol) or (ncol,nlay) concentrations"); return ECCKD_ERR_ARG; }
    }
    int rc_final = ECCKD_OK;
    if (ncol > 0 && nlay > 0) {
        if (block_cols <= 0) block_cols = 8192;
        const long long per_dev = ((long long)ncol + ndev - 1) / ndev;
        const int B = (int)std::min<long long>(block_cols, per_dev);
        // device block buffers get a column pitch that is a multiple of 4 elements (16 bytes in fp32): every block is
        // then eligible for the TMA-store (quad) kernel whatever the caller's ncol; padding columns are zero-filled
        // once, computed and never copied back
        const int Bp = (B + 3) & ~3;
        // slot layout (same for every slot)
        Carve probe{nullptr};
        auto carve = [&](Carve &c, Real *&d_plev, Real *&d_tlay, Real *&d_tlev, Real *&d_tsfc, Real **d_vmr, Real **d_out, Real *&d_2d) {
            d_plev = c.take<Real>((size_t)Bp * (nlay + 1));
            d_tlay = c.take<Real>((size_t)Bp * nlay);
            d_tlev = c.take<Real>((size_t)Bp * (nlay + 1));
            d_tsfc = c.take<Real>((size_t)Bp);
            for (int j = 0; j < ngas; ++j) {
                const ecckd_vmr_view &v = vmr[j];
                size_t n = !v.ptr ? 0 : (v.col_stride == 0 ? (v.lay_stride == 0 ? 1 : (size_t)nlay)
                                                            : (v.lay_stride == 0 ? (size_t)Bp : (size_t)Bp * nlay));
                d_vmr[j] = c.take<Real>(n);
            }
            for (int o = 0; o < 4; ++o) d_out[o] = c.take<Real>((size_t)Bp * nlay * ngpt);
            d_2d = c.take<Real>((size_t)Bp * ngpt);
        };
        {
            Real *a, *b, *c2, *d2, *e2, *v2[ECCKD_MAX_GASES], *o[4];
            carve(probe, a, b, c2, d2, v2, o, e2);
        }
        const size_t need = probe.off + 256;
        for (int d = 0; d < ndev; ++d) {
            DevState &ds = *M->devs[d];
            CUDA_TRY(cudaSetDevice(ds.device));
            for (Slot &s : ds.slots) {
                if (!s.stream) CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
                if (s.capacity < need) {
                    if (s.arena) { CUDA_TRY(cudaStreamSynchronize(s.stream)); CUDA_TRY(cudaFree(s.arena)); s.arena = nullptr; s.capacity = 0; }
                    CUDA_TRY(cudaMalloc(&s.arena, need));
                    s.capacity = need;
                    CUDA_TRY(cudaMemsetAsync(s.arena, 0, need, s.stream));
                }
            }
        }
        const long long nblk = (per_dev + B - 1) / B;
        const size_t rs = sizeof(Real);
        // Pageable caller arrays (what a Fortran host model passes) make every cudaMemcpy2DAsync below a synchronous,
        // driver-staged copy: the three pipeline slots then do not overlap.  A host model that wants the overlapped path
        // page-locks its arrays once (ecckd_host_register, or its own cudaHostRegister / pinned allocator).
        struct Pin {
            std::vector<void *> ptrs;
            ~Pin() { for (void *p : ptrs) cudaHostUnregister(p); }
            void add(const void *p, size_t bytes) {
                // opt-in: page-locking costs about as much as the copy it speeds up (measured on the B200 box: ~2.5 GB/s),
                // so it only pays when the same arrays are passed again and again -- and then ecckd_host_register() once is
                // the better tool.  ECCKD_PIN_MIN_BYTES=<n> turns per-call registration on for arrays of at least n bytes.
                static const size_t min_bytes = [] { const char *e = getenv("ECCKD_PIN_MIN_BYTES"); return e ? (size_t)atoll(e) : ~(size_t)0; }();
                if (!p || bytes < min_bytes) return;
                if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterPortable) == cudaSuccess) ptrs.push_back(const_cast<void *>(p));
                else (void)cudaGetLastError();   // already pinned by the caller, or not registrable: copies still work
            }
        } pin;
        {
            NvtxRange nvtx("ecckd: page-lock caller arrays");
            const size_t n3 = (size_t)ncol * nlay * ngpt * rs;
            pin.add(o0, n3);
            if (!(sw && not_2str)) { pin.add(o1, n3); pin.add(o2, n3); }
            if (!sw && !lev_missing) pin.add(o3, n3);
            pin.add(plev, (size_t)ncol * (nlay + 1) * rs); pin.add(tlay, (size_t)ncol * nlay * rs);
            pin.add(tlev, (size_t)ncol * (nlay + 1) * rs);
            for (int j = 0; j < ngas; ++j)
                if (vmr[j].ptr && vmr[j].col_stride == 1 && vmr[j].lay_stride != 0) pin.add(vmr[j].ptr, (size_t)vmr[j].lay_stride * nlay * rs);
        }
        for (long long blk = 0; blk < nblk; ++blk) {
            for (int d = 0; d < ndev; ++d) {
                const long long c_lo = std::min<long long>((long long)d * per_dev, ncol);
                const long long c_hi = std::min<long long>(c_lo + per_dev, ncol);
                const long long c0 = c_lo + blk * B;
                if (c0 >= c_hi) continue;
                const int nb = (int)std::min<long long>(B, c_hi - c0);
                const int nbp = (nb + 3) & ~3;                      // device pitch of this block (<= Bp)
                DevState &ds = *M->devs[d];
                CUDA_TRY(cudaSetDevice(ds.device));
                Slot &s = ds.slots[blk % kSlots];
                Carve c{s.arena};
                Real *d_plev, *d_tlay, *d_tlev, *d_tsfc, *d_vmr[ECCKD_MAX_GASES], *d_out[4], *d_2d;
                carve(c, d_plev, d_tlay, d_tlev, d_tsfc, d_vmr, d_out, d_2d);
                // H2D: a column block of a (ncol, n) array is n rows of nb elements with pitch ncol
                nvtxRangePushA("ecckd: H2D block");
                CUDA_TRY(cudaMemcpy2DAsync(d_plev, nbp * rs, plev + c0, (size_t)ncol * rs, nb * rs, nlay + 1, cudaMemcpyHostToDevice, s.stream));
                CUDA_TRY(cudaMemcpy2DAsync(d_tlay, nbp * rs, tlay + c0, (size_t)ncol * rs, nb * rs, nlay, cudaMemcpyHostToDevice, s.stream));
                if (tlev) CUDA_TRY(cudaMemcpy2DAsync(d_tlev, nbp * rs, tlev + c0, (size_t)ncol * rs, nb * rs, nlay + 1, cudaMemcpyHostToDevice, s.stream));
                if (tsfc) CUDA_TRY(cudaMemcpyAsync(d_tsfc, tsfc + c0, nb * rs, cudaMemcpyHostToDevice, s.stream));
                ecckd_vmr_view dv[ECCKD_MAX_GASES];
                for (int j = 0; j < ngas; ++j) {
                    const ecckd_vmr_view &v = vmr[j];
                    dv[j] = ecckd_vmr_view{nullptr, 0, 0};
                    if (!v.ptr) continue;
                    const Real *hp = static_cast<const Real *>(v.ptr);
                    if (v.col_stride == 0 && v.lay_stride == 0) {
                        CUDA_TRY(cudaMemcpyAsync(d_vmr[j], hp, rs, cudaMemcpyHostToDevice, s.stream));
                        dv[j] = ecckd_vmr_view{d_vmr[j], 0, 0};
                    } else if (v.col_stride == 0) {
                        CUDA_TRY(cudaMemcpyAsync(d_vmr[j], hp, nlay * rs, cudaMemcpyHostToDevice, s.stream));
                        dv[j] = ecckd_vmr_view{d_vmr[j], 0, 1};
                    } else if (v.lay_stride == 0) {
                        CUDA_TRY(cudaMemcpyAsync(d_vmr[j], hp + c0, nb * rs, cudaMemcpyHostToDevice, s.stream));
                        dv[j] = ecckd_vmr_view{d_vmr[j], 1, 0};
                    } else {
                        CUDA_TRY(cudaMemcpy2DAsync(d_vmr[j], nbp * rs, hp + c0, (size_t)v.lay_stride * rs, nb * rs, nlay, cudaMemcpyHostToDevice, s.stream));
                        dv[j] = ecckd_vmr_view{d_vmr[j], 1, nbp};
                    }
                }
                nvtxRangePop();
                char lerr[ECCKD_ERRLEN];
                int rc = run_device<Real>(M, d, sw, nbp, nlay, d_plev, d_tlay, tlev ? d_tlev : nullptr, d_tsfc, ngas, names, dv,
                                          d_out[0], (sw && not_2str) ? nullptr : d_out[1], (sw && not_2str) ? nullptr : d_out[2],
                                          d_out[3], d_2d, ECCKD_KERNEL_AUTO, s.stream, lerr);
                if (rc != ECCKD_OK && rc != ECCKD_ERR_REFERENCE) { set_err(err, lerr); return rc; }
                if (rc == ECCKD_ERR_REFERENCE) { rc_final = rc; set_err(err, lerr); }
                // D2H: every output that the reference would have written
                NvtxRange nvtx_d2h("ecckd: D2H block");
                Real *host_out[4] = {o0, o1, o2, o3};
                const int nout = sw ? 3 : 4;
                for (int o = 0; o < nout; ++o) {
                    bool written = true;
                    if (sw && not_2str && o > 0) written = false;
                    if (!sw && lev_missing && o >= 2) written = false;
                    if (!written || !host_out[o]) continue;
                    CUDA_TRY(cudaMemcpy2DAsync(host_out[o] + c0, (size_t)ncol * rs, d_out[o], nbp * rs, nb * rs, (size_t)nlay * ngpt, cudaMemcpyDeviceToHost, s.stream));
                }
                if (o2d && !(sw && not_2str))
                    CUDA_TRY(cudaMemcpy2DAsync(o2d + c0, (size_t)ncol * rs, d_2d, nbp * rs, nb * rs, ngpt, cudaMemcpyDeviceToHost, s.stream));
            }
        }
        for (int d = 0; d < ndev; ++d) {
            DevState &ds = *M->devs[d];
            CUDA_TRY(cudaSetDevice(ds.device));
            for (Slot &s : ds.slots) CUDA_TRY(cudaStreamSynchronize(s.stream));
        }
    } else {
        if (lev_missing) { set_err(err, "tlev is required for ecckd"); return ECCKD_ERR_REFERENCE; }
        if (not_2str) { set_err(err, "shortwave must use ty_optical_props_2str"); return ECCKD_ERR_REFERENCE; }
    }
    return rc_final;
}

}  // namespace

// ============================================================================
// C ABI
// ============================================================================
extern "C" {

void ecckd_debug_walk_plan(int ncol, int shortwave, int sm_count, int *tiles_per_chunk, int *blocks)
{
    int tpc = 1, nb = 1;
    const int ntile = ncol > 0 ? (ncol + ecckd::kWalkTileCols - 1) / ecckd::kWalkTileCols : 0;
    ecckd::walk_plan(ntile, shortwave != 0, sm_count, tpc, nb);
    if (tiles_per_chunk) *tiles_per_chunk = tpc;
    if (blocks) *blocks = nb;
}

const char *ecckd_version(void) { return "ecckd_b200 0.1 (sm_100a; fp32 tables promoted in-kernel, fp64 copy for the f64 path)"; }

int ecckd_model_create(const ecckd_host_tables *t, int ndev, const int *devices, ecckd_model **out, char err[ECCKD_ERRLEN])
{
    if (err) err[0] = '\0';
    if (!t || !out) { set_err(err, "ecckd_model_create: NULL argument"); return ECCKD_ERR_ARG; }
    *out = nullptr;
    if (t->ngpt < 1 || t->npress < 2 || t->ntemp < 2 || t->num_gases < 0 || t->num_gases > ECCKD_MAX_GASES || !t->log_pressure ||
        !t->temperature || (t->num_gases > 0 && !t->gas)) {
        set_err(err, "ecckd_model_create: bad table extents");
        return ECCKD_ERR_ARG;
    }
    if (t->shortwave ? (!t->solar_irradiance || !t->rayleigh_molar_scattering_coeff)
                     : (!t->temperature_planck || !t->planck_function || t->nplanck < 2)) {
        set_err(err, "ecckd_model_create: missing source tables");
        return ECCKD_ERR_ARG;
    }
    if (ndev < 1) { set_err(err, "ecckd_model_create: ndev < 1"); return ECCKD_ERR_ARG; }
    std::unique_ptr<ecckd_model> M(new ecckd_model);
    M->shortwave = t->shortwave != 0;
    M->ngpt = t->ngpt; M->np = t->npress; M->nt = t->ntemp; M->nplanck = t->shortwave ? 0 : t->nplanck;
    M->log_pressure.assign(t->log_pressure, t->log_pressure + t->npress);
    M->temperature.assign(t->temperature, t->temperature + (size_t)t->npress * t->ntemp);
    if (t->nband > 0 && t->band_lims_wvn && t->band2gpt) {
        M->nband = t->nband;
        M->band_lims.assign(t->band_lims_wvn, t->band_lims_wvn + 2 * (size_t)t->nband);
        M->band2gpt.assign(t->band2gpt, t->band2gpt + 2 * (size_t)t->nband);
    }
    if (t->shortwave) {
        double s = 0.0;                                                   // sum(solar_irradiance), mo_load_coefficients.F90:89
        for (int g = 0; g < t->ngpt; ++g) s += t->solar_irradiance[g];
        M->total_solar = s;
    }

    // ---- loader hardening ---------------------------------------------------
    // The reference derives every grid spacing from the first two entries and silently assumes the grid is uniform
    // (gas_optics_ecckd.f90:104-106,154-155,271).  The same is done here -- results must match the reference -- but
    // tables that break the assumption are flagged (ecckd_get_table_warnings) or, with ECCKD_STRICT_GRIDS=1, refused.
    {
        const double tol = 1e-3;
        auto nonuniform = [&](const double *x, int n, int stride) {
            if (n < 3) return false;
            const double d0 = x[stride] - x[0];
            for (int i = 2; i < n; ++i)
                if (std::fabs((x[(size_t)i * stride] - x[(size_t)(i - 1) * stride]) - d0) > tol * std::fabs(d0)) return true;
            return false;
        };
        if (nonuniform(t->log_pressure, t->npress, 1)) M->warnings += "log_pressure is not uniformly spaced (:104-105); ";
        bool tbad = false;
        for (int ip = 0; ip < t->npress && !tbad; ++ip) tbad = nonuniform(t->temperature + ip, t->ntemp, t->npress);
        const double dt0 = t->temperature[t->npress] - t->temperature[0];
        for (int ip = 1; ip < t->npress && !tbad; ++ip)
            tbad = std::fabs((t->temperature[t->npress + ip] - t->temperature[ip]) - dt0) > tol * std::fabs(dt0);
        if (tbad) M->warnings += "temperature(p,:) is not a uniform grid with one spacing for all pressures (:106,131-133); ";
        if (!t->shortwave && nonuniform(t->temperature_planck, t->nplanck, 1)) M->warnings += "temperature_planck is not uniformly spaced (:271); ";
        for (int i = 0; i < t->num_gases; ++i) {
            const ecckd_gas_table &g = t->gas[i];
            if (g.concentration_dependence_code != ECCKD_CODE_LOOK_UP_TABLE || g.n_mole_fraction < 3 || !g.mole_fraction) continue;
            std::vector<double> lg(g.n_mole_fraction);
            bool pos = true;
            for (int q = 0; q < g.n_mole_fraction; ++q) { pos = pos && g.mole_fraction[q] > 0.0; lg[q] = std::log(g.mole_fraction[q]); }
            if (!pos || nonuniform(lg.data(), g.n_mole_fraction, 1))
                M->warnings += "mole_fraction of " + trim_name(g.name, ECCKD_NAME_LEN) + " is not uniform in log space (:154-155); ";
        }
        if (!M->warnings.empty() && getenv("ECCKD_STRICT_GRIDS")) {
            set_err(err, "ecckd_model_create: " + M->warnings);
            return ECCKD_ERR_ARG;
        }
    }

    // pack all coefficient tables into one array; composite aliases share one copy
    std::vector<double> k64;
    const size_t slab = (size_t)t->ngpt * t->npress * t->ntemp;
    long long composite_off = -1;
    const double *composite_src = nullptr;
    for (int i = 0; i < t->num_gases; ++i) {
        const ecckd_gas_table &g = t->gas[i];
        const int code = g.concentration_dependence_code;
        if (code < 0 || code > 3 || !g.coefficient) { set_err(err, "ecckd_model_create: bad gas table"); return ECCKD_ERR_ARG; }
        const int nv = code == ECCKD_CODE_LOOK_UP_TABLE ? g.n_mole_fraction : 1;
        if (code == ECCKD_CODE_LOOK_UP_TABLE && (nv < 2 || !g.mole_fraction)) {
            set_err(err, "ecckd_model_create: look-up-table gas needs >= 2 mole fractions");
            return ECCKD_ERR_ARG;
        }
        M->gas.push_back(trim_name(g.name, ECCKD_NAME_LEN));
        M->composite_only.push_back(g.composite_only != 0);
        M->code.push_back(code);
        M->nv.push_back(nv);
        M->ref.push_back(g.reference_mole_fraction);
        if (code == ECCKD_CODE_LOOK_UP_TABLE) {
            M->mf0.push_back(g.mole_fraction[0]);
            M->log_mf0.push_back(std::log(g.mole_fraction[0]));                          // :156
            M->dlog_mf.push_back(std::log(g.mole_fraction[1] / g.mole_fraction[0]));     // :154-155
        } else { M->mf0.push_back(0.0); M->log_mf0.push_back(0.0); M->dlog_mf.push_back(1.0); }
        const size_t n = slab * nv;
        if (g.composite_only && composite_off >= 0 && composite_src &&
            std::memcmp(composite_src, g.coefficient, n * sizeof(double)) == 0) {
            M->table_off.push_back(composite_off);
            continue;
        }
        M->table_off.push_back((long long)k64.size());
        if (g.composite_only) { composite_off = (long long)k64.size(); composite_src = g.coefficient; }
        k64.insert(k64.end(), g.coefficient, g.coefficient + n);
    }
    if (k64.size() >= (size_t)1 << 30) { set_err(err, "ecckd_model_create: tables too large"); return ECCKD_ERR_ARG; }
    std::vector<float> k32(k64.size());
    for (size_t q = 0; q < k64.size(); ++q) {
        k32[q] = (float)k64[q];
        if ((double)k32[q] != k64[q]) M->fp32_exact = false;
        if (k64[q] < 0.0) M->tables_nonneg = false;
    }
    std::vector<double> p64;
    std::vector<float> p32;
    if (!t->shortwave) {
        p64.assign(t->planck_function, t->planck_function + (size_t)t->ngpt * t->nplanck);
        p32.resize(p64.size());
        for (size_t q = 0; q < p64.size(); ++q) { p32[q] = (float)p64[q]; if ((double)p32[q] != p64[q]) M->fp32_exact = false; }
    }
    // ---- walk tables: padded, g-grouped copies with compile-time corner offsets ----
    // Qualifies when extents fit the padding, at most one look-up-table gas, at most
    // kMaxBil distinct bilinear tables, and no negative coefficient (so that the
    // per-g clamp :233-238 reduces to a per-gas sign test of the weight).
    std::vector<double> wbil64, wlut64;
    {
        int nlut = 0, lut_gas = -1, nslot = 0;
        std::vector<long long> slot_off;
        M->slot.assign(t->num_gases, -1);
        for (int i = 0; i < t->num_gases; ++i) {
            if (M->code[i] == ECCKD_CODE_LOOK_UP_TABLE) { ++nlut; lut_gas = i; continue; }
            int sidx = -1;
            for (int q = 0; q < nslot; ++q) if (slot_off[q] == M->table_off[i]) sidx = q;
            if (sidx < 0) { sidx = nslot++; slot_off.push_back(M->table_off[i]); }
            M->slot[i] = sidx;
        }
        M->walk_model = M->tables_nonneg && M->fp32_exact && nlut <= 1 && nslot <= kMaxBil && t->npress <= kNPP && t->ntemp <= kNTP &&
                        M->nplanck <= 256;   // the per-block shared-memory copy of the Planck LUT must fit beside records and tiles
        if (M->walk_model) {
            const int ngroups = (t->ngpt + kGP - 1) / kGP, ng = t->ngpt, np = t->npress, nt = t->ntemp;
            const int nvl = lut_gas >= 0 ? M->nv[lut_gas] : 0;
            wbil64.assign((size_t)ngroups * std::max(nslot, 1) * kSlab, 0.0);
            wlut64.assign((size_t)ngroups * std::max(nvl, 1) * kSlab, 0.0);
            for (int grp = 0; grp < ngroups; ++grp)
                for (int ln = 0; ln < kGP; ++ln) {
                    const int g = grp * kGP + ln;
                    if (g >= ng) continue;
                    for (int ip = 0; ip < np; ++ip)
                        for (int it = 0; it < nt; ++it) {
                            const size_t cell = ((size_t)ip * kNTP + it) * kGP + walk_pos64(ln);
                            const size_t src = (size_t)g + (size_t)ng * (ip + (size_t)np * it);
                            for (int q = 0; q < nslot; ++q)
                                wbil64[((size_t)grp * nslot + q) * kSlab + cell] = k64[slot_off[q] + src];
                            for (int iv = 0; iv < nvl; ++iv)
                                wlut64[((size_t)grp * nvl + iv) * kSlab + cell] =
                                    k64[M->table_off[lut_gas] + src + (size_t)ng * np * nt * iv];
                        }
                }
        }
        M->nslot_ = nslot; M->nv_lut_ = lut_gas >= 0 ? M->nv[lut_gas] : 0;
    }
    // fp32 copies: rows re-permuted from the 8-byte to the 4-byte layout (see walk_pos32 / walk_pos64)
    auto narrow = [](const std::vector<double> &v) {
        std::vector<float> f(v.size());
        for (size_t q = 0; q < v.size(); ++q) f[(q & ~size_t(kGP - 1)) + walk_pos32(walk_inv64((int)(q & (kGP - 1))))] = (float)v[q];
        return f;
    };
    const std::vector<float> wbil32 = narrow(wbil64), wlut32 = narrow(wlut64);
    std::vector<double> tbase(t->temperature, t->temperature + t->npress);               // temperature(:,1)
    std::vector<double> ray, sol;
    if (t->shortwave) {
        ray.assign(t->rayleigh_molar_scattering_coeff, t->rayleigh_molar_scattering_coeff + t->ngpt);
        sol.assign(t->solar_irradiance, t->solar_irradiance + t->ngpt);
    }

    int ndev_avail = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev_avail));
    for (int d = 0; d < ndev; ++d) {
        const int dev = devices ? devices[d] : d;
        if (dev < 0 || dev >= ndev_avail) { set_err(err, "ecckd_model_create: no such CUDA device"); return ECCKD_ERR_CUDA; }
        CUDA_TRY(cudaSetDevice(dev));
        {   // per-call scratch comes from the stream-ordered pool: keep freed blocks cached instead of returning them to
            // the driver at every synchronisation (the default release threshold is 0, which makes each call re-allocate)
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = 64ull << 20;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
        }
        std::unique_ptr<DevState> ds(new DevState);
        ds->device = dev;
        DevModel &dm = ds->dm;
        dm.ngpt = t->ngpt; dm.np = t->npress; dm.nt = t->ntemp; dm.nplanck = M->nplanck;
        dm.sm_count = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&dm.sm_count, cudaDevAttrMultiProcessorCount, dev));
        dm.logp0 = t->log_pressure[0];
        dm.dlogp = t->log_pressure[1] - t->log_pressure[0];
        dm.dt = t->temperature[t->npress] - t->temperature[0];                           // temperature(1,2)-temperature(1,1)
        dm.gw = 1.0 / (((double)9.80665f * (double)0.001f) * (double)28.970f);           // :107 with :51-52
        dm.p_clamp = (double)t->npress - 1.0001;
        dm.t_clamp = (double)t->ntemp - 1.0001;
        dm.pi = (double)3.14159265359f;                                                  // :53
        if (!t->shortwave) { dm.tp0 = t->temperature_planck[0]; dm.dtp = t->temperature_planck[1] - t->temperature_planck[0]; }
        dm.r_dlogp = const_reciprocal(dm.dlogp); dm.r_dt = const_reciprocal(dm.dt); dm.r_pi = const_reciprocal(dm.pi);
        dm.r_dtp = t->shortwave ? std::nan("") : const_reciprocal(dm.dtp);
        dm.r_tp0 = t->shortwave ? std::nan("") : const_reciprocal(dm.tp0);
        int rc;
        M->devs.push_back(std::move(ds));
        DevState &d2 = *M->devs.back();
        if ((rc = upload(d2, tbase, &d2.dm.tbase, err))) return rc;
        if ((rc = upload(d2, k64, &d2.dm.k64, err))) return rc;
        if ((rc = upload(d2, k32, &d2.dm.k32, err))) return rc;
        if ((rc = upload(d2, p64, &d2.dm.planck64, err))) return rc;
        if ((rc = upload(d2, p32, &d2.dm.planck32, err))) return rc;
        if ((rc = upload(d2, ray, &d2.dm.rayleigh, err))) return rc;
        if ((rc = upload(d2, sol, &d2.dm.solar, err))) return rc;
        d2.dm.nslot = M->nslot_; d2.dm.nv_lut = M->nv_lut_;
        if ((rc = upload(d2, wbil64, &d2.dm.wbil64, err))) return rc;
        if ((rc = upload(d2, wlut64, &d2.dm.wlut64, err))) return rc;
        if (M->fp32_exact) {
            if ((rc = upload(d2, wbil32, &d2.dm.wbil32, err))) return rc;
            if ((rc = upload(d2, wlut32, &d2.dm.wlut32, err))) return rc;
        }
    }
    *out = M.release();
    return ECCKD_OK;
}

int ecckd_model_load(const char *ckd_file, int ndev, const int *devices, ecckd_model **model, char err[ECCKD_ERRLEN])
{
    if (err) err[0] = '\0';
    if (!ckd_file || !model) { set_err(err, "ecckd_model_load: NULL argument"); return ECCKD_ERR_ARG; }
    try {
        HostModel hm = load_and_init(ckd_file);
        ecckd_host_tables t = hm.view();
        return ecckd_model_create(&t, ndev, devices, model, err);
    } catch (const std::exception &e) {
        set_err(err, e.what());
        return ECCKD_ERR_IO;
    }
}

void ecckd_model_destroy(ecckd_model *model) { delete model; }

namespace {
struct TablesHolder {
    HostModel hm;
    ecckd_host_tables view;
};
std::mutex g_tables_mutex;
std::map<const ecckd_host_tables *, std::unique_ptr<TablesHolder>> g_tables;
}  // namespace

int ecckd_tables_load(const char *ckd_file, const ecckd_host_tables **tables, char err[ECCKD_ERRLEN])
{
    if (err) err[0] = '\0';
    if (!ckd_file || !tables) { set_err(err, "ecckd_tables_load: NULL argument"); return ECCKD_ERR_ARG; }
    *tables = nullptr;
    try {
        std::unique_ptr<TablesHolder> h(new TablesHolder);
        h->hm = load_and_init(ckd_file);
        h->view = h->hm.view();
        const ecckd_host_tables *key = &h->view;
        std::lock_guard<std::mutex> lock(g_tables_mutex);
        g_tables[key] = std::move(h);
        *tables = key;
        return ECCKD_OK;
    } catch (const std::exception &e) {
        set_err(err, e.what());
        return ECCKD_ERR_IO;
    }
}

void ecckd_tables_free(const ecckd_host_tables *tables)
{
    std::lock_guard<std::mutex> lock(g_tables_mutex);
    g_tables.erase(tables);
}

int ecckd_get_ngas(const ecckd_model *m) { return m ? (int)m->gas.size() : 0; }                       // :477-483
int ecckd_get_gas_name(const ecckd_model *m, int i, char name[ECCKD_NAME_LEN])                         // :507-513
{
    if (!m || i < 0 || i >= (int)m->gas.size() || !name) return ECCKD_ERR_ARG;
    std::memset(name, 0, ECCKD_NAME_LEN);
    std::memcpy(name, m->gas[i].data(), std::min<size_t>(m->gas[i].size(), ECCKD_NAME_LEN - 1));
    return ECCKD_OK;
}
int ecckd_get_ngpt(const ecckd_model *m) { return m ? m->ngpt : 0; }
int ecckd_get_nband(const ecckd_model *m) { return m ? m->nband : 0; }
int ecckd_get_band2gpt(const ecckd_model *m, int32_t *b)
{
    if (!m || !b) return ECCKD_ERR_ARG;
    std::copy(m->band2gpt.begin(), m->band2gpt.end(), b);
    return ECCKD_OK;
}
int ecckd_get_band_lims_wvn(const ecckd_model *m, double *b)
{
    if (!m || !b) return ECCKD_ERR_ARG;
    std::copy(m->band_lims.begin(), m->band_lims.end(), b);
    return ECCKD_OK;
}
int ecckd_host_register(void *ptr, size_t bytes, char err[ECCKD_ERRLEN])
{
    if (err) err[0] = '\0';
    if (!ptr || bytes == 0) { set_err(err, "ecckd_host_register: NULL / empty range"); return ECCKD_ERR_ARG; }
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { (void)cudaGetLastError(); return ECCKD_OK; }
    if (e != cudaSuccess) { (void)cudaGetLastError(); set_err(err, std::string("ecckd_host_register: ") + cudaGetErrorString(e)); return ECCKD_ERR_CUDA; }
    return ECCKD_OK;
}

int ecckd_host_unregister(void *ptr)
{
    if (!ptr) return ECCKD_ERR_ARG;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return ECCKD_ERR_CUDA; }
    return ECCKD_OK;
}

const char *ecckd_get_table_warnings(const ecckd_model *m) { return m ? m->warnings.c_str() : ""; }
int ecckd_get_ndev(const ecckd_model *m) { return m ? (int)m->devs.size() : 0; }
int ecckd_get_device(const ecckd_model *m, int i) { return (m && i >= 0 && i < (int)m->devs.size()) ? m->devs[i]->device : -1; }
int ecckd_source_is_internal(const ecckd_model *m) { return m && !m->shortwave; }                      // :487-493
int ecckd_source_is_external(const ecckd_model *m) { return m && m->shortwave; }                       // :497-503
double ecckd_get_press_min(const ecckd_model *m) { return m ? std::exp(m->log_pressure.front()) : 0.0; }   // :517-523
double ecckd_get_press_max(const ecckd_model *m) { return m ? std::exp(m->log_pressure.back()) : 0.0; }    // :527-533
double ecckd_get_temp_min(const ecckd_model *m) { return m ? *std::min_element(m->temperature.begin(), m->temperature.end()) : 0.0; }  // :537-543
double ecckd_get_temp_max(const ecckd_model *m) { return m ? *std::max_element(m->temperature.begin(), m->temperature.end()) : 0.0; }  // :547-553
double ecckd_get_total_solar_irradiance(const ecckd_model *m) { return m ? m->total_solar : 0.0; }

int ecckd_gas_optics_lw_f64(ecckd_model *model, int dev_index, int ncol, int nlay, const double *plev, const double *tlay,
                            const double *tlev, const double *tsfc, int ngas, const char *const *gas_names,
                            const ecckd_vmr_view *vmr, double *tau, double *lay_source, double *lev_source_inc,
                            double *lev_source_dec, double *sfc_source, int kernel, void *stream, char err[ECCKD_ERRLEN])
{
    return run_device<double>(model, dev_index, false, ncol, nlay, plev, tlay, tlev, tsfc, ngas, gas_names, vmr, tau, lay_source,
                              lev_source_inc, lev_source_dec, sfc_source, kernel, (cudaStream_t)stream, err);
}
int ecckd_gas_optics_lw_f32(ecckd_model *model, int dev_index, int ncol, int nlay, const float *plev, const float *tlay,
                            const float *tlev, const float *tsfc, int ngas, const char *const *gas_names,
                            const ecckd_vmr_view *vmr, float *tau, float *lay_source, float *lev_source_inc,
                            float *lev_source_dec, float *sfc_source, int kernel, void *stream, char err[ECCKD_ERRLEN])
{
    return run_device<float>(model, dev_index, false, ncol, nlay, plev, tlay, tlev, tsfc, ngas, gas_names, vmr, tau, lay_source,
                             lev_source_inc, lev_source_dec, sfc_source, kernel, (cudaStream_t)stream, err);
}
int ecckd_gas_optics_sw_f64(ecckd_model *model, int dev_index, int ncol, int nlay, const double *plev, const double *tlay,
                            int ngas, const char *const *gas_names, const ecckd_vmr_view *vmr, double *tau, double *ssa,
                            double *g, double *toa_src, int kernel, void *stream, char err[ECCKD_ERRLEN])
{
    return run_device<double>(model, dev_index, true, ncol, nlay, plev, tlay, nullptr, nullptr, ngas, gas_names, vmr, tau, ssa, g,
                              nullptr, toa_src, kernel, (cudaStream_t)stream, err);
}
int ecckd_gas_optics_sw_f32(ecckd_model *model, int dev_index, int ncol, int nlay, const float *plev, const float *tlay,
                            int ngas, const char *const *gas_names, const ecckd_vmr_view *vmr, float *tau, float *ssa,
                            float *g, float *toa_src, int kernel, void *stream, char err[ECCKD_ERRLEN])
{
    return run_device<float>(model, dev_index, true, ncol, nlay, plev, tlay, nullptr, nullptr, ngas, gas_names, vmr, tau, ssa, g,
                             nullptr, toa_src, kernel, (cudaStream_t)stream, err);
}

int ecckd_gas_optics_lw_host_f64(ecckd_model *model, int ncol, int nlay, const double *plev, const double *tlay,
                                 const double *tlev, const double *tsfc, int ngas, const char *const *gas_names,
                                 const ecckd_vmr_view *vmr, double *tau, double *lay_source, double *lev_source_inc,
                                 double *lev_source_dec, double *sfc_source, int block_cols, char err[ECCKD_ERRLEN])
{
    return run_host<double>(model, false, ncol, nlay, plev, tlay, tlev, tsfc, ngas, gas_names, vmr, tau, lay_source,
                            lev_source_inc, lev_source_dec, sfc_source, block_cols, err);
}
int ecckd_gas_optics_lw_host_f32(ecckd_model *model, int ncol, int nlay, const float *plev, const float *tlay,
                                 const float *tlev, const float *tsfc, int ngas, const char *const *gas_names,
                                 const ecckd_vmr_view *vmr, float *tau, float *lay_source, float *lev_source_inc,
                                 float *lev_source_dec, float *sfc_source, int block_cols, char err[ECCKD_ERRLEN])
{
    return run_host<float>(model, false, ncol, nlay, plev, tlay, tlev, tsfc, ngas, gas_names, vmr, tau, lay_source,
                           lev_source_inc, lev_source_dec, sfc_source, block_cols, err);
}
int ecckd_gas_optics_sw_host_f64(ecckd_model *model, int ncol, int nlay, const double *plev, const double *tlay, int ngas,
                                 const char *const *gas_names, const ecckd_vmr_view *vmr, double *tau, double *ssa, double *g,
                                 double *toa_src, int block_cols, char err[ECCKD_ERRLEN])
{
    return run_host<double>(model, true, ncol, nlay, plev, tlay, nullptr, nullptr, ngas, gas_names, vmr, tau, ssa, g, nullptr,
                            toa_src, block_cols, err);
}
int ecckd_gas_optics_sw_host_f32(ecckd_model *model, int ncol, int nlay, const float *plev, const float *tlay, int ngas,
                                 const char *const *gas_names, const ecckd_vmr_view *vmr, float *tau, float *ssa, float *g,
                                 float *toa_src, int block_cols, char err[ECCKD_ERRLEN])
{
    return run_host<float>(model, true, ncol, nlay, plev, tlay, nullptr, nullptr, ngas, gas_names, vmr, tau, ssa, g, nullptr,
                           toa_src, block_cols, err);
}

int ecckd_interpolation_indices_f64(ecckd_model *M, int dev_index, int ncol, int nlay, const double *plev, const double *tlay,
                                    const char *lut_gas, const ecckd_vmr_view *lut_vmr, int32_t *ip0, double *pw1,
                                    int32_t *it0, double *tw1, int32_t *iv0, double *vw1, void *stream, char err[ECCKD_ERRLEN])
{
    if (err) err[0] = '\0';
    if (!M || dev_index < 0 || dev_index >= (int)M->devs.size()) { set_err(err, "ecckd: bad model/device"); return ECCKD_ERR_ARG; }
    if (ncol <= 0 || nlay <= 0) return ECCKD_OK;
    GasRef lut{};
    bool has = false;
    if (lut_gas) {
        const std::string want = trim_name(lut_gas, 4096);
        for (size_t q = 0; q < M->gas.size(); ++q)
            if (M->gas[q] == want && M->code[q] == kLut) {
                if (!lut_vmr || !lut_vmr->ptr) { set_err(err, "ecckd: NULL vmr for look-up-table gas"); return ECCKD_ERR_ARG; }
                lut.vmr = lut_vmr->ptr; lut.col_stride = lut_vmr->col_stride; lut.lay_stride = lut_vmr->lay_stride;
                lut.code = kLut; lut.nv = M->nv[q]; lut.mf0 = M->mf0[q]; lut.log_mf0 = M->log_mf0[q]; lut.dlog_mf = M->dlog_mf[q];
                lut.v_clamp = (double)M->nv[q] - 1.001;
                lut.r_dlog_mf = const_reciprocal(lut.dlog_mf);
                has = true;
                break;
            }
        if (!has) { set_err(err, "ecckd: " + want + " is not a look-up-table gas of this model"); return ECCKD_ERR_ARG; }
    }
    DevState &d = *M->devs[dev_index];
    CUDA_TRY(cudaSetDevice(d.device));
    CUDA_TRY(launch_indices<double>(d.dm, has ? &lut : nullptr, ncol, nlay, plev, tlay, ip0, pw1, it0, tw1, iv0, vw1, (cudaStream_t)stream));
    return ECCKD_OK;
}

int ecckd_rte_lw_f64(ecckd_model *M, int dev_index, int ncol, int nlay, int ngpt, const double *tau, const double *lay_source,
                     const double *lev_source_inc, const double *lev_source_dec, const double *sfc_source, double sfc_emis,
                     double diffusivity, double *flux_up, double *flux_dn, void *stream, char err[ECCKD_ERRLEN])
{
    if (err) err[0] = '\0';
    if (!M || dev_index < 0 || dev_index >= (int)M->devs.size()) { set_err(err, "ecckd: bad model/device"); return ECCKD_ERR_ARG; }
    if (ncol < 0 || nlay < 0 || ngpt < 1) { set_err(err, "ecckd: bad extent"); return ECCKD_ERR_ARG; }
    if (ncol == 0 || nlay == 0) return ECCKD_OK;
    if (!tau || !lay_source || !lev_source_inc || !lev_source_dec || !sfc_source || !flux_up || !flux_dn) { set_err(err, "ecckd: NULL array argument"); return ECCKD_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(M->devs[dev_index]->device));
    NvtxRange nvtx("ecckd: rte_lw (device flux solver)");
    CUDA_TRY(launch_rte_lw(ncol, nlay, ngpt, tau, lay_source, lev_source_inc, lev_source_dec, sfc_source, sfc_emis, diffusivity, flux_up, flux_dn,
                           (cudaStream_t)stream));
    return ECCKD_OK;
}

int ecckd_rte_sw_f64(ecckd_model *M, int dev_index, int ncol, int nlay, int ngpt, const double *tau, const double *ssa, const double *g,
                     const double *toa_src, double mu0, double sfc_albedo, double *flux_up, double *flux_dn, double *flux_dn_dir,
                     void *stream, char err[ECCKD_ERRLEN])
{
    if (err) err[0] = '\0';
    if (!M || dev_index < 0 || dev_index >= (int)M->devs.size()) { set_err(err, "ecckd: bad model/device"); return ECCKD_ERR_ARG; }
    if (ncol < 0 || nlay < 0 || ngpt < 1 || !(mu0 > 0.0)) { set_err(err, "ecckd: bad extent / mu0"); return ECCKD_ERR_ARG; }
    if (ncol == 0 || nlay == 0) return ECCKD_OK;
    if (!tau || !ssa || !g || !toa_src || !flux_up || !flux_dn || !flux_dn_dir) { set_err(err, "ecckd: NULL array argument"); return ECCKD_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(M->devs[dev_index]->device));
    NvtxRange nvtx("ecckd: rte_sw (device flux solver)");
    // Work space: three (ncol, nlay+1, ngpt) arrays, kept per device and only ever grown (a stream-ordered allocation of
    // several GB per call would be handed back to the driver at every synchronisation).  Calls on one device must
    // therefore not overlap on different streams.
    DevState &d = *M->devs[dev_index];
    const size_t need = rte_sw_scratch_bytes(ncol, nlay, ngpt);
    if (d.rte_scratch_bytes < need) {
        CUDA_TRY(cudaDeviceSynchronize());
        if (d.rte_scratch) { CUDA_TRY(cudaFree(d.rte_scratch)); d.rte_scratch = nullptr; d.rte_scratch_bytes = 0; }
        CUDA_TRY(cudaMalloc(&d.rte_scratch, need));
        d.rte_scratch_bytes = need;
    }
    CUDA_TRY(launch_rte_sw(ncol, nlay, ngpt, tau, ssa, g, toa_src, mu0, sfc_albedo, d.rte_scratch, flux_up, flux_dn, flux_dn_dir, (cudaStream_t)stream));
    return ECCKD_OK;
}

int ecckd_debug_div_const(int64_t n, const double *x_dev, double divisor, uint64_t *mismatches, char err[ECCKD_ERRLEN])
{
    if (err) err[0] = '\0';
    if (n < 0 || !x_dev || !mismatches) { set_err(err, "ecckd_debug_div_const: bad argument"); return ECCKD_ERR_ARG; }
    unsigned long long *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(d, 0, sizeof(unsigned long long)));
    cudaError_t e = n > 0 ? launch_div_const_check(n, x_dev, divisor, const_reciprocal(divisor), d, nullptr) : cudaSuccess;
    unsigned long long h = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    CUDA_TRY(e);
    *mismatches = h;
    return ECCKD_OK;
}

int ecckd_planck_indices_f64(ecckd_model *M, int dev_index, int64_t n, const double *temperature, int32_t *it0, double *w1,
                             void *stream, char err[ECCKD_ERRLEN])
{
    if (err) err[0] = '\0';
    if (!M || dev_index < 0 || dev_index >= (int)M->devs.size()) { set_err(err, "ecckd: bad model/device"); return ECCKD_ERR_ARG; }
    if (M->shortwave) { set_err(err, "ecckd: model has no Planck source (shortwave tables)"); return ECCKD_ERR_ARG; }
    if (n <= 0) return ECCKD_OK;
    DevState &d = *M->devs[dev_index];
    CUDA_TRY(cudaSetDevice(d.device));
    CUDA_TRY(launch_planck_indices(d.dm, n, temperature, it0, w1, (cudaStream_t)stream));
    return ECCKD_OK;
}

}  // extern "C"
