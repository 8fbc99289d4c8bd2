// spec.cuh -- numeric specification shared by the sm_100a kernels of libamps_b200.
//
// Everything that decides a hard symbol is written with explicit round-to-nearest intrinsics
// (__fmaf_rn / __fmul_rn / __fadd_rn / __fdiv_rn / packed f32x2 FMA), never with contractible
// a*b+c expressions, so the result is a pure function of the inputs and of the operation order
// documented in DESIGN.md section 3 -- the same order tests/ reproduce on the CPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace amps {

// ---- rates and geometry of the 10 MS/s RECC receive chain --------------------------------
constexpr int   kD1        = 25;    // CIC^3 decimation 10 MS/s -> 400 kS/s (the reference's rate, grc/ampsbs.grc:263)
constexpr int   kNCic      = 73;    // boxcar25 (*) boxcar25 (*) boxcar25
constexpr int   kD2        = 2;     // freq_xlating_fir_filter_ccc decimation (grc/ampsbs.grc:1834)
constexpr int   kMaxLpf    = 299;   // lpf_taps length (grc/ampsbs.grc:138-184)
constexpr int   kOS        = 10;    // demod samples per Manchester half-symbol (clock_recovery omega, :1807)
constexpr int   kTrig      = 74;    // lib/recc_impl.cc:76-77
constexpr int   kCapture   = 3374;  // lib/recc_impl.cc:70
constexpr int   kSpan      = kOS * (kTrig + kCapture - 1);   // last demod offset a capture touches
constexpr int   kBurstLen  = kOS * (kTrig + kCapture);       // search resumes this far after a hit

// ---- packed fp32x2 helpers (Blackwell FFMA2 / FMUL2 / FADD2) ------------------------------
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 splat(float s) { return make_float2(s, s); }

// x * w for complex x, w:  re = fma(-x.im, w.im, x.re*w.re), im = fma(x.im, w.re, x.re*w.im)
__device__ __forceinline__ float2 cmul(float2 x, float2 w) {
    float2 t = mul2(splat(x.x), w);
    return fma2(splat(x.y), make_float2(-w.y, w.x), t);
}

// sin/cos of a 32-bit phase (2^32 = one turn): nearest-quadrant reduction, minimax polynomials on
// [-pi/4, pi/4] (Cephes sinf/cosf coefficients), fixed evaluation order.
__device__ __forceinline__ float2 sincos_phase(uint32_t psi) {
    uint32_t quad = (psi + 0x20000000u) >> 30;
    int32_t  frac = (int32_t)(psi - (quad << 30));
    float a = __fmul_rn(__int2float_rn(frac), 1.46291807926715968e-9f);   // pi / 2^31
    float z = __fmul_rn(a, a);
    float sp = __fmaf_rn(__fmaf_rn(__fmaf_rn(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f),
                         __fmul_rn(z, a), a);
    float cp = __fmaf_rn(__fmaf_rn(__fmaf_rn(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f),
                         __fmul_rn(z, z), __fmaf_rn(-0.5f, z, 1.0f));
    float c, s;
    switch (quad & 3u) {
        case 0:  c = cp;  s = sp;  break;
        case 1:  c = -sp; s = cp;  break;
        case 2:  c = -cp; s = -sp; break;
        default: c = sp;  s = -cp; break;
    }
    return make_float2(c, s);
}

// atan2 with octant reduction and the Cephes atanf polynomial; selects instead of branches.
__device__ __forceinline__ float atan2_spec(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    if (mx == 0.0f) return 0.0f;
    float r = __fdiv_rn(mn, mx);
    bool big = r > 0.4142135679721832f;
    float t = big ? __fdiv_rn(__fadd_rn(r, -1.0f), __fadd_rn(r, 1.0f)) : r;
    float z = __fmul_rn(t, t);
    float p = __fmaf_rn(__fmaf_rn(__fmaf_rn(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z, -3.33329491539e-1f);
    float a = __fmaf_rn(__fmul_rn(p, z), t, t);
    if (big) a = __fadd_rn(a, 0.785398163397448309f);
    if (ay > ax) a = __fadd_rn(1.57079632679489662f, -a);
    if (x < 0.0f) a = __fadd_rn(3.14159265358979324f, -a);
    if (y < 0.0f) a = -a;
    return a;
}

// ---- mbarrier / 1-D TMA (cp.async.bulk) primitives ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    uint32_t a = smem_u32(bar);
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    }
}
// global -> shared bulk copy (TMA, no tensor map needed for a linear run); bytes % 16 == 0
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace amps
