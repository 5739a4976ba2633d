// ecckd_pipeline.cuh -- device helpers shared by the walking kernels: explicit-rounding arithmetic, mbarrier
// hand-offs, TMA (cp.async.bulk.tensor) stores, predicated loads / stores.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cudaTypedefs.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include "ecckd_device.cuh"

namespace ecckd {

template <typename Real> __device__ __forceinline__ void lds4(const Real *p, Real &a, Real &b, Real &c, Real &d);
template <> __device__ __forceinline__ void lds4<double>(const double *p, double &a, double &b, double &c, double &d)
{
    const double2 u = *reinterpret_cast<const double2 *>(p), v = *reinterpret_cast<const double2 *>(p + 2);
    a = u.x; b = u.y; c = v.x; d = v.y;
}
template <> __device__ __forceinline__ void lds4<float>(const float *p, float &a, float &b, float &c, float &d)
{
    const float4 u = *reinterpret_cast<const float4 *>(p);
    a = u.x; b = u.y; c = u.z; d = u.w;
}
template <typename Real> __device__ __forceinline__ void lds2(const Real *p, Real &a, Real &b);
template <> __device__ __forceinline__ void lds2<double>(const double *p, double &a, double &b)
{
    const double2 u = *reinterpret_cast<const double2 *>(p);
    a = u.x; b = u.y;
}
template <> __device__ __forceinline__ void lds2<float>(const float *p, float &a, float &b)
{
    const float2 u = *reinterpret_cast<const float2 *>(p);
    a = u.x; b = u.y;
}

// Explicit rounding / fusion: every instantiation of the walking kernel (TMA or writer-warp path, any tile
// shape) must produce bit-identical values, so that column shards, host blocks and odd/even ncol agree bitwise.
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double div_(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }
// a / b to within an ulp without the IEEE-division slow path: hardware reciprocal seed (20 bits), one Newton step
// (40 bits), and a residual correction of the quotient that squares that error again.  ssa = tau_rayleigh / tau only
// has to meet 1e-12, not correct rounding.
__device__ __forceinline__ double div_fast(double a, double b)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = __fma_rn(r, __fma_rn(-b, r, 1.0), r);
    const double q = __dmul_rn(a, r);
    return __fma_rn(__fma_rn(-b, q, a), r, q);
}
__device__ __forceinline__ float div_fast(float a, float b) { return __fdiv_rn(a, b); }

// c0*k0 + c1*k1 + c2*k2 + c3*k3 as one fused chain
template <typename Real, typename K>
__device__ __forceinline__ Real bil4(Real c0, Real c1, Real c2, Real c3, K k0, K k1, K k2, K k3)
{
    return fma_(c3, (Real)k3, fma_(c2, (Real)k2, fma_(c1, (Real)k1, mul_(c0, (Real)k0))));
}

// sleep between barrier probes (ns) of the store warp (a ring of tiles ahead of it) and the index warps (32 layers ahead)
#ifndef ECCKD_STORE_NAP
#define ECCKD_STORE_NAP 100
#endif
#ifndef ECCKD_INDEX_NAP
#define ECCKD_INDEX_NAP 400
#endif
constexpr unsigned kStoreNap = ECCKD_STORE_NAP, kIndexNap = ECCKD_INDEX_NAP;

// corner offsets inside a slab (elements): (ip,it) (ip+1,it) (ip,it+1) (ip+1,it+1)
constexpr int kC00 = 0, kC10 = kNTP * kGP, kC01 = kGP, kC11 = kNTP * kGP + kGP;

// experiment switches of the tuning build (ECCKD_DBG environment variable); compiled out of the production library
#ifdef ECCKD_TUNE
#define ECCKD_DBG(bit) ((a.dbg & (bit)) != 0)
#else
#define ECCKD_DBG(bit) false
#endif

// ---- mbarrier helpers (shared::cta, generic proxy only) -------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)   // release.cta
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)   // acquire.cta
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// The same wait for roles that run far ahead of (or behind) their partner: sleep between probes instead of spinning.
// Every probe is an LSU instruction, and the L1 data pipe is this kernel's busiest unit (measured: the spinning store
// and index warps added 6 probes per layer and column to ~22 useful shared/global accesses).
__device__ __forceinline__ void mbar_wait_relaxed(unsigned long long *bar, unsigned parity, unsigned ns)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "nanosleep.u32 %2;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long *bar, unsigned parity)   // non-blocking probe (acquire.cta)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Make a loop-invariant value opaque to the optimiser, so that it is kept in a register instead of being
// rematerialised from kernel parameters and special registers in every iteration of the layer loop.
template <typename T> __device__ __forceinline__ void keep(const T *&p) { asm volatile("" : "+l"(p)); }
template <typename T> __device__ __forceinline__ void keep(T *&p) { asm volatile("" : "+l"(p)); }
__device__ __forceinline__ void keep(int &v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void keep(unsigned &v) { asm volatile("" : "+r"(v)); }
// the same barrier operations on a precomputed shared-window address (saves the generic -> shared conversion per use)
template <int OFF = 0> __device__ __forceinline__ void mbar_arrive_a(unsigned addr)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0+%1];\n\t}" ::"r"(addr), "n"(OFF) : "memory");
}
template <int OFF = 0> __device__ __forceinline__ void mbar_wait_a(unsigned addr, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0+%2], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(addr), "r"(parity), "n"(OFF) : "memory");
}
__device__ __forceinline__ bool mbar_test_a(unsigned addr, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok != 0;
}

// ---- TMA (cp.async.bulk.tensor) store helpers ------------------------------------
struct TmaMaps { alignas(64) CUtensorMap m[4]; };
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *smem, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// Predicated global load that writes its destination register directly (tied operand): the refill of
// the register cache stays branch-free and, more importantly, nothing has to wait for the data until
// the value is really used one layer later (a compiler-generated load + MOV would stall on the MOV).
__device__ __forceinline__ void ldg_if(double &dst, const double *p, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.nc.f64 %0, [%1];\n\t}" : "+d"(dst) : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ void ldg_if(float &dst, const float *p, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}" : "+f"(dst) : "l"(p), "r"((int)pred));
}

// predicated streaming store (keeps the Planck g-loop free of divergent branches)
__device__ __forceinline__ void stcs_if(double *p, double v, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.cs.f64 [%0], %1;\n\t}" ::"l"(p), "d"(v), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void stcs_if(float *p, float v, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.cs.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"((int)pred) : "memory");
}

// register rebalancing between warp roles (whole warpgroups)
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// cuTensorMapEncodeTiled through the runtime's driver-entry-point lookup (no link-time dependency on libcuda); ecckd_kernels.cu
PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder();

}  // namespace ecckd
