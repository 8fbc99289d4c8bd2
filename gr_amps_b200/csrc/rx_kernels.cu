// rx_kernels.cu -- fused RECC receive path for sm_100a.
//
//   rx_front_kernel  : IQ @10 MS/s --TMA--> smem -> NCO rotate -> CIC^3 /25 -> 299-tap channel
//                      filter /2 (the reference's lpf_taps @400 kS/s) -> quadrature demod -> d[] @200 kS/s.
//                      Replaces freq_xlating_fir_filter_ccc + quadrature_demod_cf
//                      (grc/ampsbs.grc:1814-1872, 774-816) and the 25x front-end decimation the
//                      10 MS/s configs need.
//   rx_detect_kernel : exact 74/74 trigger match on sign(d) at every sampling phase + soft
//                      correlation of the matches (the recc_impl.cc:118 memmem, at 10 phases).
//   rx_select_kernel : run grouping, sampling-phase choice, 3374-symbol capture
//                      (recc_impl.cc:124-126), Manchester decode + BCH validity + field parse
//                      (recc_decode_impl.cc:81-169).
//
// No tensor cores: there is no dense contraction on this path.  The front kernel is a persistent
// streaming kernel: each CTA owns a contiguous run of passes, keeps the filter history in shared
// memory and re-reads only one warm-up pass at the start of its run.
#include "rx_kernels.cuh"
#include "recc_compat.cuh"

namespace amps {

static_assert(sizeof(amps_burst) % 8 == 0, "burst records are streamed to the host ring in 8-byte words");

// ============================================================================================
// front end
// ============================================================================================
// 400 kS/s samples are kept as PAIRS (v[2P], v[2P+1]) = one float4, de-interleaved over kR rows by
// P mod kR: a thread that produces outputs R*c .. R*c+R-1 then walks pairs whose row is a
// compile-time constant and whose column is c + const, so consecutive lanes read consecutive
// 16-byte slots (conflict-free LDS.128) and every load feeds up to 2*kR FFMA2.
struct PassSmem {                         // what stage 2 + demod work on (shared by the 10 MS/s and the 400 kS/s front ends)
    float4   v[kR][kRowLen];              // row r, column c (c >= -kPorchCols) at v[r][c + kPorchCols]
    float2   ylast[kTB];                  // each thread's last output of the current pass
    float2   ycarry[2];                   // last output of a pass, by pass parity
};
// input sample formats: fc32 (gr_complex, what the reference's flowgraph carries) and sc16 (interleaved int16 I/Q, what
// the USRP puts on the wire before UHD's host-side conversion, grc/ampsbs.grc:3750): x = (float)int16 * in_scale
// kUnit: the scale is a power of two and has been folded into the NCO tables on the host -- (I s) w and I (s w) are the
// same real number when s is a power of two, so the result is bit-identical and the two multiplies per sample go away.
template <bool kUnit> __device__ __forceinline__ float2 to_c32(float2 v, float) { return v; }
template <bool kUnit> __device__ __forceinline__ float2 to_c32(short2 v, float s) {
    if (kUnit) return make_float2((float)v.x, (float)v.y);
    return make_float2(__fmul_rn((float)v.x, s), __fmul_rn((float)v.y, s));
}

template <typename In>
struct FrontSmem {
    In       in[kStages][kTile];          // TMA landing ring
    PassSmem ps;
    float2   pb[3][2][kTB];               // [tile % 3][P1|P2][block] rotated CIC partial sums (3 buffers: a tile reads its
                                          // own and the previous tile's, the next tile may already be writing)
    uint64_t full[kStages];
};

size_t rx_front_smem_bytes() { return sizeof(FrontSmem<float2>); }

template <typename In>
__device__ __forceinline__ void issue_tile(const RxFrontParams &p, FrontSmem<In> *sm, long tile, int stage) {
    // tiles with a negative index come from the history buffer (kWarmTiles tiles long)
    const In *src = tile < 0 ? static_cast<const In *>(p.tail) + (long)kHist + tile * (long)kTile
                             : static_cast<const In *>(p.chunk) + tile * (long)kTile;
    mbar_expect_tx(&sm->full[stage], kTile * (uint32_t)sizeof(In));
    tma_load_1d(sm->in[stage], src, kTile * (uint32_t)sizeof(In), &sm->full[stage]);
}

__host__ __device__ constexpr int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// y[q] = sum_k h2[k] v[2q-k] for the kR outputs q = kR*c0 + r, each as two FFMA2 chains (even taps,
// odd taps), k ascending -- the order the oracle uses.  vcol points at column c0 of row 0.
__device__ __forceinline__ void channel_filter(const RxFrontParams &p, const float4 *vcol, float2 (&y)[kR]) {
    float2 E[kR], O[kR];
#pragma unroll
    for (int r = 0; r < kR; ++r) { E[r] = make_float2(0.f, 0.f); O[r] = make_float2(0.f, 0.f); }
#pragma unroll
    for (int s = 0; s < 150 + kR - 1; ++s) {
        // pair P = kR*c0 + (kR-1) - s : row and column offset are compile-time
        const int row = (kR - 1 - s) & (kR - 1);
        const int cs  = floor_div(kR - 1 - s, kR);
        const float4 pr = vcol[row * kRowLen + cs];
#pragma unroll
        for (int r = 0; r < kR; ++r) {
            const int j = r - (kR - 1) + s;           // tap pair index for output r
            if (j >= 0 && j <= 149) E[r] = fma2(splat(p.h2[2 * j]), make_float2(pr.x, pr.y), E[r]);
            if (j >= 1 && j <= 149) O[r] = fma2(splat(p.h2[2 * j - 1]), make_float2(pr.z, pr.w), O[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < kR; ++r) y[r] = add2(E[r], O[r]);
}

// store one 400 kS/s sample (index relative to the start of the current pass, negative = history) into the pair/row layout
__device__ __forceinline__ void store_v(PassSmem *ps, int mrel, float2 v) {
    const int P   = mrel >> 1;                                   // pair index (floor)
    const int col = P >> kLogR;
    if (col >= -kPorchCols) reinterpret_cast<float2 *>(&ps->v[P & (kR - 1)][col + kPorchCols])[mrel & 1] = v;
}

// End of a pass (all threads call it): stage 2 (299-tap channel filter /2, kR outputs per thread), quadrature demod,
// hard decisions, and the history shuffle for the next pass.  With warm == true it only produces y[q0-1], the
// predecessor the demod of the CTA's first real output needs; nothing is written for it.
__device__ __forceinline__ void finish_pass(const RxFrontParams &p, PassSmem *ps, bool warm, int pc, uint32_t pa, int t) {
    __syncthreads();                                   // the pass's new v samples are in place
    float2 y[kR];
#pragma unroll
    for (int r = 0; r < kR; ++r) y[r] = make_float2(0.f, 0.f);
    if (!warm || t == 0) {
        const int c0 = warm ? -1 : t;
        channel_filter(p, &ps->v[0][c0 + kPorchCols], y);
        if (!warm) ps->ylast[t] = y[kR - 1];
        if (warm || t == kTB - 1) ps->ycarry[pc & 1] = y[kR - 1];
    }
    __syncthreads();
    if (!warm) {
        // the last kPorchCols columns become the history of the next pass
        if (t < kR * kPorchCols) {
            const int r = t / kPorchCols, c = t % kPorchCols;
            ps->v[r][c] = ps->v[r][kTB + c];
        }
        // ---- quadrature demod: arg(y[q] * conj(y[q-1]))
        float2 yp = t == 0 ? ps->ycarry[(pc + 1) & 1] : ps->ylast[t - 1];
        float d[kR];
#pragma unroll
        for (int r = 0; r < kR; ++r) {
            const float zr = __fmaf_rn(y[r].y, yp.y, __fmul_rn(y[r].x, yp.x));
            const float zi = __fmaf_rn(y[r].y, yp.x, -__fmul_rn(y[r].x, yp.y));
            d[r] = atan2_spec(zi, zr);
            yp = y[r];
        }
        const unsigned long long ql = (unsigned long long)(pa + (uint32_t)pc) * kPassOut + (unsigned long long)kR * t;
        const unsigned long long qabs = p.q_base + ql;
        float4 *dst = reinterpret_cast<float4 *>(&p.dring[qabs & p.dmask]);
        *dst = make_float4(d[0], d[1], d[2], d[3]);
        // hard decisions (binary_slicer_fb: x >= 0 -> 1), 32 per word: 8 lanes x 4 outputs
        unsigned int hb = 0;
#pragma unroll
        for (int r = 0; r < kR; ++r) hb |= (d[r] >= 0.0f ? 1u : 0u) << r;
        hb <<= 4 * (t & 7);
        hb |= __shfl_xor_sync(0xffffffffu, hb, 1);
        hb |= __shfl_xor_sync(0xffffffffu, hb, 2);
        hb |= __shfl_xor_sync(0xffffffffu, hb, 4);
        if ((t & 7) == 0) p.hring[(qabs & p.dmask) >> 5] = hb;
        if (p.ydump) {
#pragma unroll
            for (int r = 0; r < kR; ++r) p.ydump[ql + r] = y[r];
        }
    }
}

template <typename In, int kMinCtas, bool kUnit>
__global__ void __launch_bounds__(kTB, kMinCtas) rx_front_kernel(const __grid_constant__ RxFrontParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FrontSmem<In> *sm = reinterpret_cast<FrontSmem<In> *>(smem_raw);
    const int t = threadIdx.x;

    // history for the next call = the last kHist samples of this chunk; every CTA copies a slice (saves a memcpy node
    // between consecutive front kernels)
    if (p.tail_out) {
        const uint32_t per = ((uint32_t)kHist + gridDim.x - 1u) / gridDim.x;
        const uint32_t lo = blockIdx.x * per, hi = lo + per < (uint32_t)kHist ? lo + per : (uint32_t)kHist;
        const In *src = static_cast<const In *>(p.chunk) + ((size_t)p.npass * kPass - kHist);
        for (uint32_t i = lo + (uint32_t)t; i < hi; i += kTB) static_cast<In *>(p.tail_out)[i] = src[i];
    }

    uint32_t pa = blockIdx.x * p.pass_per_cta;
    if (pa >= p.npass) return;
    uint32_t pb = pa + p.pass_per_cta;
    if (pb > p.npass) pb = p.npass;
    const int  ntiles = kPassTiles * (int)(pb - pa) + kWarmTiles;     // warm-up tiles + the CTA's own passes
    const long tile0  = (long)kPassTiles * pa - kWarmTiles;

    if (t == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&sm->full[s], 1);
        mbar_fence_init();
    }
    // partial sums "before the first tile": they only reach samples the warm-up never uses
    sm->pb[2][0][t] = make_float2(0.f, 0.f);
    sm->pb[2][1][t] = make_float2(0.f, 0.f);
    __syncthreads();
    if (t == 0) {
        for (int s = 0; s < kStages && s < ntiles; ++s) issue_tile(p, sm, tile0 + s, s);
    }

    for (int i = 0; i < ntiles; ++i) {
        const int s = i % kStages;
        mbar_wait(&sm->full[s], (uint32_t)(i / kStages) & 1u);

        // ---- stage 1: NCO rotate + CIC^3 polyphase partial sums over this thread's 25 samples
        const In *xin = sm->in[s] + kD1 * t;
        float2 P0 = make_float2(0.f, 0.f), P1 = P0, P2 = P0;
#pragma unroll
        for (int k = 0; k < kD1; ++k) {
            const float2 x = to_c32<kUnit>(xin[k], p.in_scale);
            const float2 u = fma2(splat(x.y), p.wj[k], mul2(splat(x.x), p.w[k]));      // = cmul(x, w[k]), same operations
            P0 = fma2(splat(p.g[24 - k]), u, P0);
            P1 = fma2(splat(p.g[49 - k]), u, P1);
            if (74 - k < kNCic) P2 = fma2(splat(p.g[74 - k]), u, P2);
        }
        const long     blk  = (tile0 + i) * (long)kTB + t;
        const uint32_t babs = p.blk_base + (uint32_t)blk;
        const float2   W    = sincos_phase(babs * p.fcw25);
        P0 = cmul(P0, W);
        P1 = cmul(P1, W);
        P2 = cmul(P2, W);
        const int par = i % 3, prv = (i + 2) % 3;
        sm->pb[par][0][t] = P1;
        sm->pb[par][1][t] = P2;
        __syncthreads();                                   // partials visible; in[s] fully consumed
        if (t == 0 && i + kStages < ntiles) issue_tile(p, sm, tile0 + i + kStages, s);

        // ---- v[m] = (P0[m] + P1[m-1]) + P2[m-2], stored into the pair/row layout
        const float2 q1 = t >= 1 ? sm->pb[par][0][t - 1] : sm->pb[prv][0][kTB - 1];
        const float2 q2 = t >= 2 ? sm->pb[par][1][t - 2] : sm->pb[prv][1][kTB - 2 + t];
        const float2 v  = add2(add2(P0, q1), q2);
        const bool warm = i < kWarmTiles;
        const int  u    = warm ? 0 : (i - kWarmTiles) % kPassTiles;          // tile index inside the pass
        store_v(&sm->ps, warm ? (i - kWarmTiles) * kTB + t : u * kTB + t, v);

        const bool pass_end = !warm && u == kPassTiles - 1;
        if (pass_end || i == kWarmTiles - 1)
            finish_pass(p, &sm->ps, warm, warm ? -1 : (i - kWarmTiles) / kPassTiles, pa, t);
    }
}

// ---------------------------------------------------------------------------------------------
// native-rate front end: complex IQ @400 kS/s (the reference's own operating point, grc/ampsbs.grc:263).
// No decimating first stage: v[m] = x[m] e^{-j theta m}, then exactly freq_xlating_fir_filter_ccc's filter /2 and
// quadrature_demod_cf.  At 0.4 MS/s real time this kernel is never a bottleneck (it is FFMA-bound: 150 FFMA2 per
// 8-byte sample); it exists so that recc_iq drops into the reference flowgraph at its native rate.
// ---------------------------------------------------------------------------------------------
template <typename In, bool kUnit>
__global__ void __launch_bounds__(kTB, 4) rx_front400_kernel(const __grid_constant__ RxFrontParams p) {
    __shared__ PassSmem ps;
    const int t = threadIdx.x;
    uint32_t pa = blockIdx.x * p.pass_per_cta;
    if (pa >= p.npass) return;
    uint32_t pb = pa + p.pass_per_cta;
    if (pb > p.npass) pb = p.npass;
    auto load = [&](long L) -> float2 {                   // logical sample L of this call; L < 0 = history
        const float2 x = to_c32<kUnit>(L < 0 ? static_cast<const In *>(p.tail)[(long)kPass400 + L] : static_cast<const In *>(p.chunk)[L], p.in_scale);
        const unsigned long long nabs = p.n_base + (unsigned long long)(long long)L;    // (x is 0 where this wraps: stream start)
        const uint32_t b = (uint32_t)(nabs / kD1), k = (uint32_t)(nabs % kD1);
        return cmul(fma2(splat(x.y), p.wj[k], mul2(splat(x.x), p.w[k])), sincos_phase(b * p.fcw25));
    };
    const long first = (long)pa * kPass400;
    // warm-up: the kPorchCols columns of history in front of the first pass
    for (int idx = t; idx < 2 * kR * kPorchCols; idx += kTB) {
        const int mrel = idx - 2 * kR * kPorchCols;
        store_v(&ps, mrel, load(first + mrel));
    }
    finish_pass(p, &ps, true, -1, pa, t);
    for (uint32_t pc = 0; pc < pb - pa; ++pc) {
        const long base = first + (long)pc * kPass400;
        __syncthreads();                                   // the history shuffle of the previous pass is done
#pragma unroll
        for (int u = 0; u < kPassTiles; ++u) store_v(&ps, u * kTB + t, load(base + u * kTB + t));
        finish_pass(p, &ps, false, (int)pc, pa, t);
    }
}

cudaError_t launch_rx_front400(const RxFrontParams &p, int grid, cudaStream_t st, bool sc16, bool unit) {
    if (sc16 && unit) rx_front400_kernel<short2, true><<<grid, kTB, 0, st>>>(p);
    else if (sc16) rx_front400_kernel<short2, false><<<grid, kTB, 0, st>>>(p);
    else rx_front400_kernel<float2, false><<<grid, kTB, 0, st>>>(p);
    return cudaGetLastError();
}

// sc16 tiles are half the size, so three CTAs fit an SM (the kernel is no longer HBM-bound at 4 B/sample)
constexpr int kSc16Ctas = 3;
cudaError_t launch_rx_front(const RxFrontParams &p, int grid, cudaStream_t st, bool sc16, bool unit) {
    if (sc16 && unit) rx_front_kernel<short2, kSc16Ctas, true><<<grid, kTB, sizeof(FrontSmem<short2>), st>>>(p);
    else if (sc16) rx_front_kernel<short2, kSc16Ctas, false><<<grid, kTB, sizeof(FrontSmem<short2>), st>>>(p);
    else rx_front_kernel<float2, 2, false><<<grid, kTB, sizeof(FrontSmem<float2>), st>>>(p);
    return cudaGetLastError();
}
int rx_front_ctas_per_sm(bool sc16) { return sc16 ? kSc16Ctas : 2; }

// ============================================================================================
// trigger detection
// ============================================================================================
// 74 half-symbols: Manchester("10" x 13 + "11100010010"), bit 0 -> (1,0), bit 1 -> (0,1)
// (lib/recc_impl.cc:51-65,76).
__device__ __constant__ uint8_t c_trig[kTrig] = {
    0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,
    0,1,0,1,0,1,1,0,1,0,1,0,0,1,1,0,1,0,0,1,1,0};

// 74-symbol trigger packed LSB-first (symbol k = bit k)
__device__ __constant__ uint32_t c_trig_bits[3] = {0x66666666u, 0x56A66666u, 0x00000196u};

// Exact 74/74 hard match for the 32 adjacent sampling positions i0 .. i0+31 (i0 a multiple of 32):
// for half-symbol k the 32 hard decisions at positions i0+10k .. i0+10k+31 are one 32-bit window of
// the bit ring, so one AND per symbol tests all 32 positions; a random group dies after ~6 symbols.
__device__ __forceinline__ uint32_t group_match(const uint32_t *__restrict__ hring, uint32_t wmask, unsigned long long i0) {
    uint32_t m = 0xffffffffu;
#pragma unroll 1
    for (int k = 0; k < kTrig && m; ++k) {
        const unsigned long long b = i0 + (unsigned long long)(kOS * k);
        const uint32_t wi = (uint32_t)(b >> 5);
        const uint32_t w0 = hring[wi & wmask], w1 = hring[(wi + 1) & wmask];
        const uint32_t win = __funnelshift_r(w0, w1, (uint32_t)b & 31u);
        m &= ((c_trig_bits[k >> 5] >> (k & 31)) & 1u) ? win : ~win;
    }
    return m;
}

// soft correlation of the trigger at sampling position i: sum_k (+/-) d[i + 10k], k ascending
__device__ __forceinline__ float trig_corr(const float *__restrict__ dring, uint32_t dmask, unsigned long long i) {
    float v[kTrig];
#pragma unroll
    for (int k = 0; k < kTrig; ++k) v[k] = dring[(i + (unsigned long long)(kOS * k)) & dmask];
    float c = 0.0f;
#pragma unroll
    for (int k = 0; k < kTrig; ++k) c = __fadd_rn(c, c_trig[k] ? v[k] : -v[k]);
    return c;
}

// One thread per group of 32 sampling positions.  The thread owning the FIRST position of a run of
// matches (at most 10 long: the pattern cannot match one half-symbol later) emits one candidate for
// the whole run: its soft-correlation peak (first maximum) is the sampling phase.
// (Grid-stride loop; the launcher may cap the grid.  Measured on B200: the grid size does not matter for the pipeline, the
// shared-memory carve-out preference set in rx_configure_device() does.)
__device__ void detect_group(const float *__restrict__ dring, const uint32_t *__restrict__ hring, uint32_t dmask, RxState *state,
                             Candidate *cand, unsigned long long scan_lo, unsigned long long scan_hi, unsigned long long i0) {
    const uint32_t wmask = dmask >> 5;
    const uint32_t m0 = group_match(hring, wmask, i0);
    if (!m0) return;
    // rare path: neighbours' match bits decide where runs start and end
    const uint32_t mprev = i0 >= 32 ? group_match(hring, wmask, i0 - 32) : 0u;
    const uint32_t mnext = group_match(hring, wmask, i0 + 32);
    const unsigned long long M = (unsigned long long)m0 | ((unsigned long long)mnext << 32);
    uint32_t starts = m0 & ~((m0 << 1) | (mprev >> 31));
    const unsigned long long lo = state->lo > scan_lo ? state->lo : scan_lo;
    while (starts) {
        const int bit = __ffs(starts) - 1;
        starts &= starts - 1;
        const unsigned long long i = i0 + (unsigned long long)bit;
        if (i < lo || i >= scan_hi) continue;
        unsigned long long best = i;
        float bestc = trig_corr(dring, dmask, i);
        unsigned int run = 1;
        bool open = false;
        for (;;) {
            const unsigned long long j = i + run;
            if (j >= scan_hi) { open = true; break; }                 // the run may go on in data not searched yet
            if (!((M >> (bit + run)) & 1ull)) break;
            const float c = trig_corr(dring, dmask, j);
            if (c > bestc) { bestc = c; best = j; }
            ++run;
        }
        const unsigned int slot = atomicAdd(&state->ncand, 1u);
        if (slot < (unsigned)kMaxCand) {
            cand[slot].start = i;
            cand[slot].best = best;
            cand[slot].corr = bestc;
            cand[slot].run = run | (open ? 0x80000000u : 0u);
        } else {
            atomicAdd(&state->cand_overflow, 1u);
        }
    }
}

__global__ void __launch_bounds__(256) rx_detect_kernel(const float *__restrict__ dring, const uint32_t *__restrict__ hring,
                                                       uint32_t dmask, RxState *state, Candidate *cand,
                                                       unsigned long long scan_lo, unsigned long long scan_hi) {
    const unsigned long long base = scan_lo & ~31ull;
    const unsigned long long stride = 32ull * (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i0 = base + 32ull * ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x); i0 < scan_hi; i0 += stride)
        detect_group(dring, hring, dmask, state, cand, scan_lo, scan_hi, i0);
}

cudaError_t launch_rx_detect(const float *dring, const uint32_t *hring, uint32_t dmask, RxState *state, Candidate *cand,
                             unsigned long long scan_lo, unsigned long long scan_hi, int max_ctas, cudaStream_t st) {
    if (scan_hi <= scan_lo) return cudaSuccess;
    const unsigned long long groups = (scan_hi - (scan_lo & ~31ull) + 31ull) / 32ull;
    unsigned int grid = (unsigned int)((groups + 255) / 256);
    if (max_ctas > 0 && grid > (unsigned int)max_ctas) grid = (unsigned int)max_ctas;
    rx_detect_kernel<<<grid, 256, 0, st>>>(dring, hring, dmask, state, cand, scan_lo, scan_hi);
    return cudaGetLastError();
}

// ============================================================================================
// BCH(63,51) validity + RECC word parsing on the device
// ============================================================================================
struct Gf64 {
    uint8_t exp[128];
    uint8_t log[64];
};
static constexpr Gf64 make_gf() {
    Gf64 g{};
    unsigned x = 1;
    for (int i = 0; i < 63; ++i) {
        g.exp[i] = (uint8_t)x;
        g.exp[i + 63] = (uint8_t)x;
        g.log[x] = (uint8_t)i;
        x <<= 1;
        if (x & 0x40) x ^= 0x43;          // primitive polynomial x^6 + x + 1
    }
    g.exp[126] = g.exp[0];
    g.exp[127] = g.exp[1];
    g.log[0] = 0;
    return g;
}
__device__ __constant__ Gf64 c_gf = make_gf();

// The tables are copied to shared memory by the decode routine: lanes index them with different
// values, which the constant cache would serialise.
__device__ __forceinline__ unsigned gf_mul(const Gf64 &gf, unsigned a, unsigned b) {
    return (a && b) ? gf.exp[gf.log[a] + gf.log[b]] : 0u;
}

// Validity of one 48-bit RECC word repeat as itpp::BCH(63,2,true)::decode reports it for
// (15 zeros || 48 bits) (lib/recc_decode_impl.cc:53-79): <=2 errors anywhere in the 63 bits, plus
// the S1 == 0, S3 a non-zero cube case that Berlekamp's 2-step iteration turns into a degree-3
// locator with three roots.
__device__ bool bch48_valid(const Gf64 &gf, const uint8_t *bits48) {
    unsigned s1 = 0, s3 = 0;
    for (int b = 0; b < 48; ++b) {
        if (bits48[b] & 1u) {
            const int e = 47 - b;                       // exponent of x carried by this bit
            s1 ^= gf.exp[e];
            s3 ^= gf.exp[(3 * e) % 63];
        }
    }
    if (s1 == 0) return s3 == 0 || (gf.log[s3] % 3u) == 0;
    const unsigned s1c = gf_mul(gf, gf_mul(gf, s1, s1), s1);
    if (s3 == s1c) return true;                         // single error
    // Lambda(x) = 1 + s1 x + ((s3 + s1^3)/s1) x^2 must have two roots among alpha^0..alpha^62
    const unsigned num = s3 ^ s1c;
    const unsigned c2  = gf.exp[gf.log[num] + 63 - gf.log[s1]];
    int roots = 0;
    for (int j = 0; j < 63; ++j) {
        const unsigned x  = gf.exp[j];
        const unsigned x2 = gf.exp[(2 * j) % 63];
        if ((1u ^ gf_mul(gf, s1, x) ^ gf_mul(gf, c2, x2)) == 0u) ++roots;
    }
    return roots == 2;
}

__device__ __forceinline__ unsigned getbits(const uint8_t *b, int n) {
    unsigned v = 0;
    for (int i = 0; i < n; ++i) v = (v << 1) | (b[i] & 1u);
    return v;
}

// three MIN digits from a 10-bit group, lib/amps_packet.h:277-302
__device__ void extract_min_3(unsigned val, char *out3) {
    unsigned m2 = val + 111u;
    unsigned dig = m2 % 10u;
    out3[2] = (char)('0' + dig);
    m2 -= dig == 0 ? 10u : dig;
    dig = (m2 % 100u) / 10u;
    out3[1] = (char)('0' + dig);
    if (dig == 0) m2 -= 100u; else m2 -= m2 % 100u;
    dig = m2 / 100u;
    if (dig > 9) dig = 0;
    out3[0] = (char)('0' + dig);
}

// Whole-block routine: symbols (3374 x 0/1, any address space) -> *out (global).
// `scratch` is shared memory: 7*5 validity bytes.
__device__ void decode_burst_block(const uint8_t *symbols, amps_recc_words *out, uint8_t *scratch_valid,
                                   unsigned int *scratch_errs) {
    __shared__ Gf64 gf;
    const int t = threadIdx.x, nt = blockDim.x;
    for (int i = t; i < (int)sizeof(Gf64); i += nt) reinterpret_cast<uint8_t *>(&gf)[i] = reinterpret_cast<const uint8_t *>(&c_gf)[i];
    if (t < 8) scratch_errs[t] = 0;
    __syncthreads();
    // Manchester pairs: (1,0)->0 (0,1)->1 (1,1)->0+err (0,0)->1+err   (lib/utils.cc:36-50)
    for (int o = t; o < 7 + 7 * 240; o += nt) {
        int src, w;
        uint8_t *dst;
        if (o < 7) { src = 2 * o; w = 7; dst = &out->dcc[o]; }
        else { const int oo = o - 7; w = oo / 240; src = 14 + 480 * w + 2 * (oo % 240); dst = &out->words[w][oo % 240]; }
        const unsigned a = symbols[src] & 1u, b = symbols[src + 1] & 1u;
        *dst = (uint8_t)(a ? 0 : 1);
        if (a == b) atomicAdd(&scratch_errs[w], 1u);
    }
    __syncthreads();
    if (t < 35) scratch_valid[t] = bch48_valid(gf, &out->words[t / 5][48 * (t % 5)]) ? 1 : 0;
    __syncthreads();
    if (t == 0) {
        out->dcc_errs = (uint8_t)scratch_errs[7];
        for (int w = 0; w < 7; ++w) {
            out->errs[w] = (uint16_t)scratch_errs[w];
            out->valid[w] = 0;
            out->valid_repeat[w] = 5;
            for (int r = 0; r < 5; ++r)
                if (scratch_valid[5 * w + r]) { out->valid[w] = 1; out->valid_repeat[w] = (uint8_t)r; break; }
        }
        const uint8_t *a = out->words[0], *b = out->words[1];
        out->F = a[0] & 1u; out->NAWC = (uint8_t)getbits(a + 1, 3);
        out->T = a[4] & 1u; out->S = a[5] & 1u; out->E = a[6] & 1u; out->ER = a[7] & 1u;
        out->SCM = (uint8_t)getbits(a + 8, 4);
        out->pad0 = 0;
        out->MIN1 = getbits(a + 12, 24);
        out->B_F = b[0] & 1u; out->B_NAWC = (uint8_t)getbits(b + 1, 3);
        out->MSG_TYPE = (uint8_t)getbits(b + 4, 5); out->ORDQ = (uint8_t)getbits(b + 9, 3);
        out->ORDER = (uint8_t)getbits(b + 12, 5);
        out->LT = b[17] & 1u; out->EP = b[18] & 1u; out->SCM4 = b[19];
        out->MPCI = (uint8_t)getbits(b + 20, 2); out->SDCC1 = (uint8_t)getbits(b + 22, 2); out->SDCC2 = (uint8_t)getbits(b + 24, 2);
        out->pad1 = 0; out->pad2 = 0;
        out->MIN2 = (uint16_t)getbits(b + 26, 10);
        out->word_c_serial = getbits(out->words[2] + 4, 32);
        out->esn = 0;
        for (int i = 0; i < 12; ++i) out->min[i] = 0;
        for (int i = 0; i < 36; ++i) out->dialed[i] = 0;
        // MIN string, lib/amps_packet.h:354-363
        extract_min_3(out->MIN2, out->min);
        extract_min_3((out->MIN1 >> 14) & 0x3ffu, out->min + 3);
        unsigned thous = (out->MIN1 >> 10) & 0xfu;
        if (thous > 9) thous = 0;
        out->min[6] = (char)('0' + thous);
        extract_min_3(out->MIN1 & 0x3ffu, out->min + 7);
        // message class, lib/recc_decode_impl.cc:108-168
        int kind;
        const bool order_zero = out->ORDER == 0 && out->ORDQ == 0 && out->MSG_TYPE == 0;
        if (!out->valid[0]) kind = AMPS_MSG_INVALID_A;
        else if (!out->E) kind = AMPS_MSG_E0_DROPPED;
        else if (out->T == 0 && order_zero) kind = AMPS_MSG_PAGE_RESPONSE;
        else if (out->T == 1 && out->ORDER == 0xd) {
            kind = AMPS_MSG_REGISTRATION;
            if (out->S && out->NAWC > 1) out->esn = out->word_c_serial;
        } else if (out->T == 1 && (out->NAWC > 2 || order_zero)) {
            unsigned nawc = out->NAWC, next = 2;
            if (out->S) { out->esn = out->word_c_serial; next++; nawc = (unsigned)(uint8_t)(out->NAWC - 2); }
            if (nawc < 1 || nawc > 4) kind = AMPS_MSG_BAD_NAWC;
            else {
                kind = AMPS_MSG_ORIGINATION;
                int len = 0;
                for (; nawc > 0; --nawc) {
                    unsigned digs = getbits(out->words[next] + 4, 32);
                    ++next;
                    for (int k = 0; k < 8; ++k) {          // lib/amps_packet.h:207-273
                        const unsigned v = (digs >> 28) & 0xfu;
                        if (v == 0 || v >= 13) break;
                        out->dialed[len++] = v <= 9 ? (char)('0' + v) : (v == 10 ? '0' : (v == 11 ? '*' : '#'));
                        digs <<= 4;
                    }
                }
            }
        } else kind = AMPS_MSG_UNKNOWN;
        out->kind = kind;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) decode_blobs_kernel(const uint8_t *blobs, amps_recc_words *out) {
    __shared__ uint8_t s_valid[40];
    __shared__ unsigned int s_errs[8];
    decode_burst_block(blobs + (size_t)blockIdx.x * kCapture, out + blockIdx.x, s_valid, s_errs);
}

cudaError_t launch_decode_blobs(const uint8_t *blobs, int nbursts, amps_recc_words *out, cudaStream_t st) {
    if (nbursts <= 0) return cudaSuccess;
    decode_blobs_kernel<<<nbursts, 256, 0, st>>>(blobs, out);
    return cudaGetLastError();
}

// ============================================================================================
// candidate selection (single CTA; candidates are rare) and burst capture (one CTA per burst)
// ============================================================================================
// `sorted` is global scratch (kMaxCand entries) rather than 196 KB of shared memory, so that this CTA can become resident
// next to the front-kernel CTAs of the following call.
__global__ void __launch_bounds__(256) rx_select_kernel(RxState *state, Candidate *cand, Candidate *sorted, Accepted *acc,
                                                       unsigned long long scan_hi, RxPublished *host_pub) {
    const int t = threadIdx.x, nt = blockDim.x;
    unsigned int n = state->ncand;
    if (n > (unsigned)kMaxCand) n = kMaxCand;
    // rank sort by run start (the atomics made the order arbitrary; run starts are distinct)
    for (unsigned int i = t; i < n; i += nt) {
        const Candidate c = cand[i];
        unsigned int rank = 0;
        for (unsigned int j = 0; j < n; ++j) rank += cand[j].start < c.start ? 1u : 0u;
        sorted[rank] = c;
    }
    __syncthreads();
    if (t == 0) {
        const unsigned long long lo = state->lo;
        unsigned long long resume = state->resume_at;
        unsigned long long new_lo = scan_hi > lo ? scan_hi : lo;
        unsigned int na = 0;
        for (unsigned int i = 0; i < n; ++i) {
            const Candidate c = sorted[i];
            if (c.start < lo) continue;                    // handled by an earlier call
            if (c.run & 0x80000000u) {                     // run reaches the end of the searched range: decide next call
                new_lo = c.start;
                break;
            }
            if (c.start < resume) continue;                // begins inside a burst that was already captured
            if (na < (unsigned)kMaxAccept) {
                acc[na].pos = c.best;
                acc[na].corr = c.corr;
                acc[na].run = c.run;
                ++na;
            }
            resume = c.best + (unsigned long long)kBurstLen;
        }
        state->lo = new_lo;
        state->resume_at = resume;
        state->ncand = 0;
        state->n_acc = na;
        state->done = 0;
        state->rec_base = state->nrec_total;
        state->nrec_total += na;
        if (na == 0) {                                    // nothing for the capture kernel to publish
            __threadfence_system();
            host_pub->cand_overflow = state->cand_overflow;
        }
    }
}

cudaError_t launch_rx_select(RxState *state, Candidate *cand, Accepted *acc, unsigned long long scan_hi,
                             RxPublished *host_pub, cudaStream_t st) {
    rx_select_kernel<<<1, 256, 0, st>>>(state, cand, cand + kMaxCand, acc, scan_hi, host_pub);   // cand holds 2 x kMaxCand entries
    return cudaGetLastError();
}

// blobs == nullptr: feed-forward timing, the symbols are sliced out of the demod ring at the accepted phase.
// blobs != nullptr: M&M timing mode, amps.recc already cut the 3374-byte blobs (rx_mm_recc_kernel).
__global__ void __launch_bounds__(256) rx_capture_kernel(const float *__restrict__ dring, uint32_t dmask, RxState *state,
                                                        const Accepted *acc, amps_burst *host_ring, unsigned int ring_len,
                                                        RxPublished *host_pub, unsigned int decim,
                                                        const uint8_t *__restrict__ blobs,
                                                        const unsigned long long *__restrict__ blob_sym_index) {
    __shared__ __align__(16) unsigned char rec_raw[sizeof(amps_burst)];
    __shared__ uint8_t s_valid[40];
    __shared__ unsigned int s_errs[8];
    const unsigned int n_acc = state->n_acc;
    if (blockIdx.x >= n_acc) return;
    const int t = threadIdx.x, nt = blockDim.x;
    amps_burst *rec = reinterpret_cast<amps_burst *>(rec_raw);
    if (blobs) {
        for (int s = t; s < kCapture; s += nt) rec->symbols[s] = blobs[(size_t)blockIdx.x * kCapture + s];
        if (t == 0) {
            // position bookkeeping is nominal here: the recovered half-symbol index, 10 demod samples per half-symbol
            const unsigned long long first_sym = blob_sym_index[blockIdx.x];
            const unsigned long long trig_sym = first_sym >= (unsigned long long)kTrig ? first_sym - kTrig : 0ull;
            rec->demod_index = trig_sym * (unsigned long long)kOS;
            rec->sample_index = rec->demod_index * (unsigned long long)decim;
            rec->corr = 0.0f;
            rec->run_length = 0;
            rec->pad[0] = 0; rec->pad[1] = 0;
        }
    } else {
        const Accepted a = acc[blockIdx.x];
        // the 3374 half-symbols after the trigger, sliced at the chosen sampling phase (recc_impl.cc:124-126)
        for (int s = t; s < kCapture; s += nt) {
            const float v = dring[(a.pos + (unsigned long long)(kOS * (kTrig + s))) & dmask];
            rec->symbols[s] = v >= 0.0f ? 1 : 0;
        }
        if (t == 0) {
            rec->demod_index = a.pos;
            rec->sample_index = a.pos * (unsigned long long)decim;
            rec->corr = a.corr;
            rec->run_length = a.run;
            rec->pad[0] = 0; rec->pad[1] = 0;
        }
    }
    __syncthreads();
    decode_burst_block(rec->symbols, &rec->decoded, s_valid, s_errs);
    // publish: stream the finished record into the host-visible ring (posted PCIe writes)
    const unsigned long long rec_base = state->rec_base;
    // (when one call accepts more bursts than the ring holds, only the newest ring_len are written: two CTAs must
    // never race for the same slot)
    if (n_acc - blockIdx.x <= ring_len) {
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(rec);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(&host_ring[(rec_base + blockIdx.x) % ring_len]);
        for (int w = t; w < (int)(sizeof(amps_burst) / 8); w += nt) dst[w] = src[w];
    }
    __threadfence_system();
    __syncthreads();
    if (t == 0) {
        const unsigned int prev = atomicAdd(&state->done, 1u);
        if (prev + 1 == n_acc) {                          // last CTA: every record is on its way, publish the count
            __threadfence_system();
            host_pub->cand_overflow = state->cand_overflow;
            host_pub->nrec_total = rec_base + n_acc;
        }
    }
}

cudaError_t launch_rx_capture(const float *dring, uint32_t dmask, RxState *state, const Accepted *acc, int grid,
                              amps_burst *host_ring, unsigned int ring_len, RxPublished *host_pub, unsigned int decim,
                              cudaStream_t st, const uint8_t *blobs, const unsigned long long *blob_sym_index) {
    if (grid <= 0) return cudaSuccess;
    if (grid > kMaxAccept) grid = kMaxAccept;
    rx_capture_kernel<<<grid, 256, 0, st>>>(dring, dmask, state, acc, host_ring, ring_len, host_pub, decim, blobs, blob_sym_index);
    return cudaGetLastError();
}

// ============================================================================================
// M&M timing mode (AMPS_RX_TIMING_MM): the reference graph's own serial tail
//   clock_recovery_mm_ff(10, 0.02296875, 0, 0.05, 0.005) -> binary_slicer_fb -> amps.recc
// (grc/ampsbs.grc:1751-1813, 1712-1750, 4602) on the demodulated stream.  The loop is a recurrence
// over symbols (mu, omega and the sample position all feed the next step), so ONE thread walks it;
// the other threads of the CTA only stage the demod ring into shared memory and write the symbols
// out.  The exact fp32 operation order is the one oracle/mm_timing.c states.
// ============================================================================================
constexpr int kMmStage = 4096;     // demod samples staged per round
constexpr int kMmSymStage = 512;   // symbols produced per round at most

__global__ void __launch_bounds__(128) rx_mm_kernel(const float *__restrict__ dring, uint32_t dmask, unsigned long long total_d,
                                                   MmState *st, const float *__restrict__ table, uint8_t *__restrict__ sym_out,
                                                   unsigned int sym_cap) {
    __shared__ float s_tab[kMmPhases * 8];
    __shared__ float s_d[kMmStage + 8];
    __shared__ uint8_t s_sym[kMmSymStage];
    __shared__ unsigned long long s_pos;
    __shared__ int s_n;
    const int t = threadIdx.x, nt = blockDim.x;
    for (int i = t; i < kMmPhases * 8; i += nt) s_tab[i] = table[i];
    if (t == 0) s_pos = st->pos;
    float mu = st->mu, omega = st->omega, last = st->last;          // only thread 0's copies matter
    const float omega_mid = 10.0f, gain_omega = 0.02296875f, gain_mu = 0.05f;
    const float omega_lim = __fmul_rn(omega_mid, 0.005f);
    unsigned int out_n = 0;
    __syncthreads();
    for (;;) {
        const unsigned long long base = s_pos;
        if (base + 8ull > total_d || out_n >= sym_cap) break;
        const unsigned long long avail = total_d - base;
        const int n_ld = avail < (unsigned long long)(kMmStage + 8) ? (int)avail : kMmStage + 8;
        for (int i = t; i < n_ld; i += nt) s_d[i] = dring[(base + (unsigned long long)i) & dmask];
        __syncthreads();
        if (t == 0) {
            int p = 0, n = 0;
            const unsigned int room = sym_cap - out_n;
            const int n_max = room < (unsigned)kMmSymStage ? (int)room : kMmSymStage;
            while (p + 8 <= n_ld && p < kMmStage && n < n_max) {
                const int imu = __float2int_rn(__fmul_rn(mu, 128.0f));
                const float *tp = s_tab + 8 * imu, *x = s_d + p;
                float s = __fmul_rn(tp[0], x[0]);
#pragma unroll
                for (int k = 1; k < 8; ++k) s = __fmaf_rn(tp[k], x[k], s);
                const float a = last < 0.0f ? -s : s;                 // sgn(last) * s
                const float b = s < 0.0f ? -last : last;              // sgn(s) * last
                const float mm = __fsub_rn(a, b);
                last = s;
                omega = __fadd_rn(omega, __fmul_rn(gain_omega, mm));
                const float dev = __fsub_rn(omega, omega_mid);
                omega = __fadd_rn(omega_mid, __fmul_rn(0.5f, __fsub_rn(fabsf(__fadd_rn(dev, omega_lim)), fabsf(__fsub_rn(dev, omega_lim)))));
                mu = __fadd_rn(__fadd_rn(mu, omega), __fmul_rn(gain_mu, mm));
                float f = floorf(mu);
                if (f >= 1.0f && f <= 64.0f) mu = __fsub_rn(mu, f);
                else { f = f > 64.0f ? 64.0f : 1.0f; mu = 0.0f; }
                p += (int)f;
                s_sym[n++] = s >= 0.0f ? 1 : 0;                       // binary_slicer_fb
            }
            s_pos = base + (unsigned long long)p;
            s_n = n;
        }
        __syncthreads();
        const int n = s_n;
        for (int i = t; i < n; i += nt) sym_out[out_n + i] = s_sym[i];
        out_n += (unsigned int)n;
        __syncthreads();
        if (n == 0) break;
    }
    if (t == 0) {
        st->mu = mu; st->omega = omega; st->last = last;
        st->pos = s_pos;
        st->n_new = out_n;
        st->nsym_total += out_n;
    }
}

// amps.recc on the symbols the M&M kernel just produced, in work() calls of kMmQuantum bytes (the reference sees the
// stream in scheduler-sized pieces and searches / publishes at most once per call, lib/recc_impl.cc:115-126), then the
// bookkeeping rx_select_kernel does in the feed-forward mode.
__global__ void __launch_bounds__(256) rx_mm_recc_kernel(ReccCompatState *cs, const MmState *mm, const uint8_t *__restrict__ sym,
                                                        uint8_t *blobs, unsigned long long *blob_sym_index, int max_blobs,
                                                        RxState *state, RxPublished *host_pub) {
    const unsigned int n = mm->n_new;
    const int nchunks = (int)((n + (unsigned)kMmQuantum - 1u) / (unsigned)kMmQuantum);
    int nb = recc_compat_run(cs, sym, nchunks,
                             [n](int c) { const unsigned int rem = n - (unsigned)c * (unsigned)kMmQuantum; return rem < (unsigned)kMmQuantum ? rem : (unsigned)kMmQuantum; },
                             blobs, max_blobs, blob_sym_index);
    if (threadIdx.x == 0) {
        if (nb > max_blobs) { state->cand_overflow = 1; nb = max_blobs; }
        state->n_acc = (unsigned int)nb;
        state->done = 0;
        state->rec_base = state->nrec_total;
        state->nrec_total += (unsigned long long)nb;
        if (nb == 0) {
            __threadfence_system();
            host_pub->cand_overflow = state->cand_overflow;
        }
    }
}

cudaError_t launch_rx_mm(const float *dring, uint32_t dmask, unsigned long long total_d, MmState *mm, const float *table,
                         uint8_t *sym, unsigned int sym_cap, ReccCompatState *cs, uint8_t *blobs,
                         unsigned long long *blob_sym_index, int max_blobs, RxState *state, RxPublished *host_pub,
                         cudaStream_t st) {
    rx_mm_kernel<<<1, 128, 0, st>>>(dring, dmask, total_d, mm, table, sym, sym_cap);
    rx_mm_recc_kernel<<<1, 256, 0, st>>>(cs, mm, sym, blobs, blob_sym_index, max_blobs, state, host_pub);
    return cudaGetLastError();
}

// per-device opt-in to large dynamic shared memory (call once per device after cudaSetDevice)
cudaError_t rx_configure_device() {
    cudaError_t e = cudaFuncSetAttribute(rx_front_kernel<float2, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrontSmem<float2>));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(rx_front_kernel<short2, kSc16Ctas, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrontSmem<short2>));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(rx_front_kernel<short2, kSc16Ctas, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrontSmem<short2>));
    if (e != cudaSuccess) return e;
    // The side-stream kernels share SMs with the NEXT call's front kernel, whose two CTAs need 199 KB of shared memory per
    // SM.  An SM's L1/shared split is fixed while CTAs are resident: if a kernel that wants a big L1 gets there first, the
    // front CTAs wait until it has left.  Ask for the front kernel's split everywhere.
    const void *side[] = {(const void *)rx_detect_kernel, (const void *)rx_select_kernel, (const void *)rx_capture_kernel,
                          (const void *)rx_mm_kernel, (const void *)rx_mm_recc_kernel, (const void *)rx_front_kernel<float2, 2, false>,
                          (const void *)rx_front_kernel<short2, kSc16Ctas, false>, (const void *)rx_front_kernel<short2, kSc16Ctas, true>};
    for (const void *f : side) {
        e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace amps
