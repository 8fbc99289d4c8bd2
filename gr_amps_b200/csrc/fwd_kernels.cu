// fwd_kernels.cu -- fused forward (base-station transmit) path for sm_100a, 10 MS/s out.
//
// Per carrier (FOCC @0 Hz, FVC legs @+60/+90 kHz, grc/ampsbs.grc:4650,841,904):
//   half-symbol bytes (+1/-1/0) @100 kS/s        char_to_float                 (:1159-1252)
//   -> S[i] = running sum, fm[i] = e^{j 2 pi S[i] fcw_fm / 2^32}   frequency_modulator_fc  (:574-659)
//   -> x4 polyphase interpolation, the reference's own firdes taps (193 / 321) @400 kS/s  (:2120-2229)
//   -> x5 CIC^3 to 2 MS/s with the mixer folded in, carriers summed          (:817-942,1006-1056; the x25 is ours:
//   -> x5 CIC^3 to 10 MS/s (shared), x0.5                                     the reference stops at 400 kS/s) (:1355-1405)
// One fused kernel writes 8 bytes per output sample and reads ~0.05: it is HBM-WRITE bound.  Output tiles
// are assembled in shared memory and leave through TMA bulk stores (cp.async.bulk.global.shared::cta).
//
// The FM phase is an integer prefix sum of the +-1 symbols times a 32-bit phase step, so it is exact,
// never drifts and any CTA can start anywhere; two tiny scan kernels produce it.
#include "fwd_kernels.cuh"

namespace amps {

// ---------------------------------------------------------------------------------------------
// prefix sums of the symbol streams
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fwd_scan_local_kernel(FwdScanParams p) {
    const int c = blockIdx.y;
    const uint32_t b = blockIdx.x;
    const int t = threadIdx.x;
    __shared__ int warp_tot[8];
    const uint8_t *sym = p.sym[c];
    const uint32_t base = b * (uint32_t)kFwdScanBlock + (uint32_t)t * 16u;
    int v[16];
    int run = 0;
    // 16 symbols per thread: one 128-bit load when the run is fully inside the stream (base is 16-byte aligned)
    uint4 raw = make_uint4(0u, 0u, 0u, 0u);
    if (base + 16u <= p.nsym && (reinterpret_cast<uintptr_t>(sym) & 15u) == 0) {
        raw = *reinterpret_cast<const uint4 *>(sym + base);
    } else {
        uint8_t *rb = reinterpret_cast<uint8_t *>(&raw);
        for (int k = 0; k < 16; ++k) rb[k] = base + k < p.nsym ? sym[base + k] : 0;
    }
    const uint32_t words[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int s = (int)(int8_t)((words[k >> 2] >> (8 * (k & 3))) & 0xffu);      // bytes are signed: 0x01 = +1, 0xFF = -1
        run += s;
        v[k] = run;
    }
    int incl = run;                                               // inclusive scan of the per-thread totals
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if ((t & 31) >= d) incl += o;
    }
    if ((t & 31) == 31) warp_tot[t >> 5] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < (t >> 5); ++w) woff += warp_tot[w];
    const int excl = woff + incl - run;
    if (base + 16u <= p.nsym) {
        int4 *dst = reinterpret_cast<int4 *>(p.sloc[c] + base);      // cudaMalloc'ed, base multiple of 16: aligned
#pragma unroll
        for (int k = 0; k < 4; ++k) dst[k] = make_int4(excl + v[4 * k], excl + v[4 * k + 1], excl + v[4 * k + 2], excl + v[4 * k + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t i = base + k;
            if (i < p.nsym) p.sloc[c][i] = excl + v[k];
        }
    }
    if (t == 255) p.btot[c][b] = woff + incl;
}

// one CTA per carrier: exclusive scan of the block totals (+ carry from the previous call), the carry for
// the next call, and the next call's history (the last kFwdHistLen symbols with their absolute phase sums)
__global__ void __launch_bounds__(1024) fwd_scan_blocks_kernel(FwdScanParams p) {
    const int c = blockIdx.x;
    const int t = threadIdx.x;
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const uint32_t nblk = (p.nsym + kFwdScanBlock - 1) / kFwdScanBlock;
    if (t == 0) s_carry = p.carry[c];
    __syncthreads();
    for (uint32_t base = 0; base < nblk; base += 1024) {
        const uint32_t b = base + t;
        const int v = b < nblk ? p.btot[c][b] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if ((t & 31) >= d) incl += o;
        }
        if ((t & 31) == 31) s_warp[t >> 5] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < (t >> 5); ++w) woff += s_warp[w];
        const int carry = s_carry;
        if (b < nblk) p.boff[c][b] = carry + woff + incl - v;
        __syncthreads();
        if (t == 1023) s_carry = carry + woff + incl;
        __syncthreads();
    }
    // history for the next call: last kFwdHistLen of (old history ++ this call's symbols)
    if (t < kFwdHistLen) {
        const long src = (long)p.nsym - kFwdHistLen + t;
        uint8_t s;
        int S;
        if (src >= 0) {
            s = p.sym[c][src];
            S = p.boff[c][src / kFwdScanBlock] + p.sloc[c][src];
        } else {
            s = p.hsym_old[c][kFwdHistLen + src];
            S = p.hS_old[c][kFwdHistLen + src];
        }
        p.hsym_new[c][t] = s;
        p.hS_new[c][t] = S;
    }
    __syncthreads();
    if (t == 0) p.carry[c] = s_carry;
}

cudaError_t launch_fwd_scan(const FwdScanParams &p, int ncar, cudaStream_t st) {
    if (p.nsym == 0) return cudaSuccess;
    const uint32_t nblk = (p.nsym + kFwdScanBlock - 1) / kFwdScanBlock;
    fwd_scan_local_kernel<<<dim3(nblk, (unsigned)ncar), 256, 0, st>>>(p);
    fwd_scan_blocks_kernel<<<ncar, 1024, 0, st>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// fused modulator
// ---------------------------------------------------------------------------------------------
constexpr int kFmHalf = (kFwdTileSym + 1 + kFwdMaxTap4 + 1) / 2 + 1;   // 74 (odd/even halves land in different banks)
__device__ __forceinline__ int fm_idx(int k) { return (k & 1) * kFmHalf + (k >> 1); }
struct FwdSmem {
    float2 out[kFwdTileSym * 100];             // output tile (TMA store source); its store overlaps phases 1-3a of the next tile
    float2 B[5 * (4 * kFwdTileSym + 1) + 3 + 12];   // 2 MS/s samples (carriers mixed and summed) for m = m0-1 .. m0+251;
                                               // before that, the polyphase partial sums of the helper warps (5 x 8 x 32)
    float2 fm[kFwdMaxCar][2 * kFmHalf];        // FM samples for symbols i0-1-81 .. i0+62 (145), de-interleaved: sample k sits at
                                               // fm_idx(k) = (k & 1) * kFmHalf + (k >> 1), so the polyphase warps, whose lanes are two
                                               // symbols apart, read consecutive slots (no bank conflicts)
    alignas(16) float taps[kFwdMaxCar][4 * kFwdMaxTap4];       // polyphase taps (broadcast LDS.128 instead of constant loads)
    float2 a[kFwdMaxCar][4 * (kFwdTileSym + 1) + 16 + (4 * (kFwdTileSym + 1) + 16) / 8 + 1];   // 400 kS/s samples for symbols i0-1 .. i0+62
                                               // (+ slack for unrolled reads), sample m at a_idx(m): one slot skipped every 8, so the
                                               // polyphase warps' stride-8 stores spread over all banks
};

size_t fwd_smem_bytes() { return sizeof(FwdSmem); }
static size_t fwd_voice_smem_bytes();

__device__ __forceinline__ void tma_store_1d(void *gdst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ int a_idx(int m) { return m + (m >> 3); }
__device__ __forceinline__ int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// voice legs (kVoice): the phasors a tile needs, the x25 resampler taps and the legs' rotated 400 kS/s samples
constexpr int kVoiceStage = 48;                // audio samples staged per tile: 11 new + kFwdVoicePer - 1 of history, rounded up
constexpr int kVoiceERow = kFwdVoicePer + 1;   // odd row stride: the 25 rows start in different banks
struct FwdVoiceSmem {                           // overlays FwdSmem::B, which is dead during phases 1-2 (refilled every tile)
    float  E[25 * kVoiceERow];
    float2 vph[kFwdVoiceLegs][kVoiceStage];
};
static_assert(sizeof(FwdSmem::B) >= 5 * 8 * 32 * sizeof(float2), "room for the partial sums of five helper warps");
static_assert(sizeof(FwdVoiceSmem) <= sizeof(FwdSmem::B), "voice staging must fit into the 2 MS/s buffer it overlays");
constexpr int kVoiceThreads = kFwdThreads - 96;                                   // warps 3..7
constexpr int kVoiceOutPerThread = (kFwdVoiceLegs * 4 * (kFwdTileSym + 1) + kVoiceThreads - 1) / kVoiceThreads;   // 4

static size_t fwd_voice_smem_bytes() { return sizeof(FwdSmem); }

template <bool kVoice>
__global__ void __launch_bounds__(kFwdThreads, 3) fwd_fused_kernel(const __grid_constant__ FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FwdSmem *sm = reinterpret_cast<FwdSmem *>(smem_raw);
    FwdVoiceSmem *sv = reinterpret_cast<FwdVoiceSmem *>(&sm->B[0]);
    const int t = threadIdx.x;
    const uint32_t ntiles = (p.nsym + kFwdTileSym - 1) / kFwdTileSym;
    constexpr int kFmLen = kFwdTileSym + 1 + kFwdMaxTap4;          // 145
    for (int i = t; i < kFwdMaxCar * 4 * kFwdMaxTap4; i += kFwdThreads) (&sm->taps[0][0])[i] = (&p.taps[0][0])[i];

    // The symbols and phase sums a tile needs (145 per carrier: i0-82 .. i0+62) are fetched ONE TILE AHEAD into registers,
    // two items per thread, so their global-load latency hides behind a tile of work instead of opening every tile.
    static_assert(kFwdMaxCar * kFmLen <= 2 * kFwdThreads, "two FM items per thread");
    int fc_[2], fk_[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int idx = t + r * kFwdThreads;
        fc_[r] = idx < p.ncar * kFmLen ? idx / kFmLen : -1;
        fk_[r] = idx - (fc_[r] < 0 ? 0 : fc_[r]) * kFmLen;
    }
    // (the two addends of a phase sum stay in separate registers until they are used: adding them here would wait for the loads)
    auto fetch = [&](uint32_t tile_, int r, uint8_t &s_, int &Sa_, int &Sb_) {
        s_ = 0; Sa_ = 0; Sb_ = 0;
        const int c = fc_[r];
        if (c < 0 || tile_ >= ntiles) return;
        const long i = (long)tile_ * kFwdTileSym - (kFwdMaxTap4 + 1) + fk_[r];
        if (i >= (long)p.nsym) return;
        if (i >= 0) { s_ = p.sym[c][i]; Sa_ = p.boff[c][i / kFwdScanBlock]; Sb_ = p.sloc[c][i]; }
        else if (i >= -(long)kFwdHistLen) { s_ = p.hsym[c][kFwdHistLen + i]; Sa_ = p.hS[c][kFwdHistLen + i]; }
    };
    uint8_t ns[2];
    int nSa[2], nSb[2];
    fetch(blockIdx.x, 0, ns[0], nSa[0], nSb[0]);
    fetch(blockIdx.x, 1, ns[1], nSa[1], nSb[1]);

    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long i0 = (long)tile * kFwdTileSym;                  // first new symbol of the tile
        const int nvalid = (int)((long)p.nsym - i0 < kFwdTileSym ? (long)p.nsym - i0 : kFwdTileSym);

        // ---- phase 1: FM samples fm[c][k] for symbol i0 - 82 + k, k < 145
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (fc_[r] >= 0) {
                float2 v = make_float2(0.f, 0.f);
                if (ns[r] != 0) v = sincos_phase((uint32_t)(nSa[r] + nSb[r]) * p.fcw_fm);
                sm->fm[fc_[r]][fm_idx(fk_[r])] = v;
            }
        }
        fetch(tile + gridDim.x, 0, ns[0], nSa[0], nSb[0]);
        fetch(tile + gridDim.x, 1, ns[1], nSa[1], nSb[1]);
        // voice: phasors for audio samples ia0 .. ia0 + 47, ia0 = floor(4 (i0 - 1) / 25) - (kFwdVoicePer - 1)
        const int ia0 = floor_div(4 * ((int)i0 - 1), 25) - (kFwdVoicePer - 1);
        if (kVoice) {
            for (int i = t; i < 25 * kFwdVoicePer; i += kFwdThreads) sv->E[(i / kFwdVoicePer) * kVoiceERow + (i % kFwdVoicePer)] = p.Eg[i];
            for (int idx = t; idx < kFwdVoiceLegs * kVoiceStage; idx += kFwdThreads) {
                const int l = idx / kVoiceStage, k = idx - l * kVoiceStage;
                const int ia = ia0 + k;
                float2 v = make_float2(0.f, 0.f);
                if (p.vcar[l] >= 0) {
                    if (ia >= 0) { if (ia < (int)p.n_audio) v = p.vph[l][ia]; }
                    else if (ia >= -kFwdVoiceHist) v = p.vhist[l][kFwdVoiceHist + ia];
                }
                sv->vph[l][k] = v;
            }
        }
        __syncthreads();

        // ---- phase 2: x4 polyphase interpolation.  One thread = two adjacent symbols (8 outputs) of one carrier: the
        //      FM sample loaded for symbol i at tap k is symbol i+1's sample at tap k+1, so each shared-memory load and
        //      each uniform 4-tap load feeds 8 FFMA2.  The 400 kS/s samples are stored already rotated by the carrier's
        //      NCO: at[m] = a[m] e^{j phi_c(25 m)}.
        // Without voice legs all eight warps share the taps (p.seg: e.g. 49 + 81 + 81 tap groups -> 2 + 3 + 3 warps of
        // 25..27 groups each); with voice legs the five spare warps resample the audio and each carrier keeps one warp.
        const FwdParams::Seg sg = kVoice ? FwdParams::Seg{(int8_t)((t >> 5) < p.ncar ? (t >> 5) : -1), 0, 1, 0, 0, (int16_t)((t >> 5) < p.ncar ? p.ntap4[(t >> 5) < p.ncar ? (t >> 5) : 0] : 0)}
                                         : p.seg[t >> 5];
        float2 lo0 = make_float2(0.f, 0.f), lo1 = lo0, lo2 = lo0, lo3 = lo0, hi0 = lo0, hi1 = lo0, hi2 = lo0, hi3 = lo0;
        const int ip = t & 31;                                      // symbols i0 - 1 + 2 ip, i0 + 2 ip
        if (sg.c >= 0) {
            const int c = sg.c;                                     // warp-uniform
            // fm of the pair's first symbol at tap k is sample 2 ip + 81 - k: odd samples for even k, even samples for odd k
            const float2 *F = sm->fm[c];
            const float *T = sm->taps[c];
            const int base = 2 * ip + kFwdMaxTap4;
            float2 xh = F[fm_idx(base + 1 - sg.k0)];                // fm[i+1 - k0]
            auto tap = [&](int k, float2 xl) {                      // xl = fm[i - k] == fm[(i+1) - (k+1)]
                const float4 tk = *reinterpret_cast<const float4 *>(T + 4 * k);     // same address in every lane: broadcast LDS.128
                lo0 = fma2(splat(tk.x), xl, lo0); hi0 = fma2(splat(tk.x), xh, hi0);
                lo1 = fma2(splat(tk.y), xl, lo1); hi1 = fma2(splat(tk.y), xh, hi1);
                lo2 = fma2(splat(tk.z), xl, lo2); hi2 = fma2(splat(tk.z), xh, hi2);
                lo3 = fma2(splat(tk.w), xl, lo3); hi3 = fma2(splat(tk.w), xh, hi3);
                xh = xl;
            };
            int k = sg.k0;
            if ((k & 1) && k < sg.k1) { tap(k, F[fm_idx(base - k)]); ++k; }
            const float2 *pO = F + kFmHalf + ((base - k) >> 1);      // k even: odd sample (base - k), then even sample (base - k - 1)
            const float2 *pE = F + ((base - k - 1) >> 1);
#pragma unroll 2
            for (; k + 1 < sg.k1; k += 2) {
                tap(k, *pO);
                tap(k + 1, *pE);
                --pO; --pE;
            }
            if (k < sg.k1) tap(k, *pO);
            if (!kVoice && sg.slot > 0) {                           // helper warp: hand the partial sums over
                float2 *ps = &sm->B[(sg.hidx * 8) * 32 + ip];
                ps[0 * 32] = lo0; ps[1 * 32] = lo1; ps[2 * 32] = lo2; ps[3 * 32] = lo3;
                ps[4 * 32] = hi0; ps[5 * 32] = hi1; ps[6 * 32] = hi2; ps[7 * 32] = hi3;
            }
        }
        if (!kVoice) __syncthreads();
        if (sg.c >= 0 && sg.slot == 0) {
            const int c = sg.c;
            if (!kVoice) {
                for (int s2 = 1; s2 < sg.nslot; ++s2) {
                    const float2 *ps = &sm->B[((sg.hidx + s2 - 1) * 8) * 32 + ip];
                    lo0 = add2(lo0, ps[0 * 32]); lo1 = add2(lo1, ps[1 * 32]); lo2 = add2(lo2, ps[2 * 32]); lo3 = add2(lo3, ps[3 * 32]);
                    hi0 = add2(hi0, ps[4 * 32]); hi1 = add2(hi1, ps[5 * 32]); hi2 = add2(hi2, ps[6 * 32]); hi3 = add2(hi3, ps[7 * 32]);
                }
            }
            const uint32_t m8 = p.m_base + (uint32_t)(4 * (i0 - 1 + 2 * ip));
            const float2 w25 = p.w25[c];
            // each symbol starts its own phasor recurrence, so a sample's value does not depend on how symbols pair up
            float2 W = sincos_phase(m8 * p.fcw_mix25[c]);
            float2 *a = &sm->a[c][9 * ip];                          // a_idx(8 ip + j) = 9 ip + j, j < 8
            a[0] = cmul(lo0, W); W = cmul(W, w25);
            a[1] = cmul(lo1, W); W = cmul(W, w25);
            a[2] = cmul(lo2, W); W = cmul(W, w25);
            a[3] = cmul(lo3, W);
            W = sincos_phase((m8 + 4u) * p.fcw_mix25[c]);
            a[4] = cmul(hi0, W); W = cmul(W, w25);
            a[5] = cmul(hi1, W); W = cmul(W, w25);
            a[6] = cmul(hi2, W); W = cmul(W, w25);
            a[7] = cmul(hi3, W);
        }
        // the warps the symbol carriers leave idle: x25 polyphase resampling of the voice phasors, rotated by the carrier
        // NCO like the symbol legs.  One output = 400 kS/s sample m = 4 (i0 - 1) + idx of one leg; kept in registers
        // until the symbol warps have stored theirs, then added in place (add_xx in front of the mixer).
        float2 vout[kVoiceOutPerThread];
        if (kVoice && t >= 96) {
#pragma unroll
            for (int j = 0; j < kVoiceOutPerThread; ++j) {
                const int o = t - 96 + j * kVoiceThreads;
                vout[j] = make_float2(0.f, 0.f);
                if (o >= kFwdVoiceLegs * 4 * (kFwdTileSym + 1)) continue;
                const int l = o / (4 * (kFwdTileSym + 1)), idx = o - l * 4 * (kFwdTileSym + 1);
                const int c = p.vcar[l];
                if (c < 0) continue;
                const int m = 4 * ((int)i0 - 1) + idx;
                const int ia = floor_div(m, 25);
                const int r = m - 25 * ia;
                const float *Er = &sv->E[r * kVoiceERow];
                const float2 *x = &sv->vph[l][ia - ia0];              // x[-k] = phasor of audio sample ia - k
                float2 acc = make_float2(0.f, 0.f);
                for (int k = 0; k < p.vper; ++k) acc = fma2(splat(Er[k]), x[-k], acc);
                const float2 W = sincos_phase((p.m_base + (uint32_t)m) * p.fcw_mix25[c]);
                vout[j] = cmul(acc, W);
            }
        }
        __syncthreads();
        if (kVoice) {
            if (t >= 96) {
#pragma unroll
                for (int j = 0; j < kVoiceOutPerThread; ++j) {
                    const int o = t - 96 + j * kVoiceThreads;
                    if (o >= kFwdVoiceLegs * 4 * (kFwdTileSym + 1)) continue;
                    const int l = o / (4 * (kFwdTileSym + 1)), idx = o - l * 4 * (kFwdTileSym + 1);
                    const int c = p.vcar[l];
                    if (c >= 0) sm->a[c][a_idx(idx)] = add2(sm->a[c][a_idx(idx)], vout[j]);
                }
            }
            __syncthreads();
        }

        // ---- phase 3a: x5 CIC^3 interpolation to 2 MS/s with the mixer folded into complex taps, carriers summed.
        //      B[5 m + r] = sum_c sum_j C1_c[r + 5 j] at_c[m - j];  thread = one 400 kS/s sample -> 5 outputs
        if (t < 4 * kFwdTileSym + 1) {
            float2 acc[5];
#pragma unroll
            for (int r = 0; r < 5; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < kFwdMaxCar; ++c) {
                if (c < p.ncar) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const float2 x = sm->a[c][a_idx(t + 3 - j)];
                        const float2 xr = splat(x.x), xi = splat(x.y);
#pragma unroll
                        for (int r = 0; r < 5; ++r) {
                            if (r + 5 * j < 13) {
                                const float2 C = p.C1[c][r + 5 * j];
                                acc[r] = fma2(xi, p.C1j[c][r + 5 * j], fma2(xr, C, acc[r]));
                            }
                        }
                    }
                }
            }
            float2 *Bo = &sm->B[5 * t];
#pragma unroll
            for (int r = 0; r < 5; ++r) Bo[r] = acc[r];
        }
        // the output tile we are about to refill must have left the SM (its bulk store has read it)
        if (t == 0) tma_store_wait_read0();
        __syncthreads();

        // ---- phase 3b: shared x5 CIC^3 interpolation to 10 MS/s (real taps, x out_scale folded in): 25 outputs per thread
        if (t < 4 * nvalid) {
            const float2 *Bi = &sm->B[5 * (t + 1)];                 // B[q], q = 5 t (tile-local)
            float2 b[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) b[k] = Bi[k - 2];
            float2 *o = &sm->out[25 * t];
#pragma unroll
            for (int qq = 0; qq < 5; ++qq)
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    float2 v = mul2(splat(p.G2[r]), b[qq + 2]);
                    v = fma2(splat(p.G2[r + 5]), b[qq + 1], v);
                    if (r + 10 < 13) v = fma2(splat(p.G2[r + 10]), b[qq], v);
                    o[5 * qq + r] = v;
                }
        }
        fence_async_smem();
        __syncthreads();
        if (t == 0) tma_store_1d(p.out + (size_t)i0 * 100, sm->out, (uint32_t)nvalid * 100u * (uint32_t)sizeof(float2));
    }
    if (t == 0) tma_store_wait_all();
}

// ---------------------------------------------------------------------------------------------
// Manchester-bit fast path.  Every FOCC / FVC data bit is the half-symbol pair (-1 x5, +1 x5) or (+1 x5, -1 x5), so
// the FM phase returns to zero at every bit boundary and the modulator output over one bit is one of two fixed
// 10-sample waveforms.  The x4 polyphase interpolator is linear, hence its 400 kS/s output is a sum of per-bit
// responses: a[m] = sum_{d<9} R[bit(q-d)][(m - 40 q) + 40 d], q = m div 40 -- table lookups instead of 49..81
// taps, no prefix sum, no sincos per sample ("closed-form Manchester symbol lookup").  And since a 1 is
// the mirror image of a 0, R1 = conj(R0): the real part of a[m] does not depend on the data and the imaginary part is a
// signed sum, looked up three bits at a time (RW / JW / I3 tables, fwd_kernels.cuh: kFbFastLen).
// The rest (x5 CIC^3 with folded mixers, carriers summed at 2 MS/s, shared x5 CIC^3, TMA bulk store) is the same.
// ---------------------------------------------------------------------------------------------
struct FwdBitsSmem {
    float2  out[kFbTileM * 25];                               // 5000 output samples (TMA store source)
    float2  B[5 * (kFbTileM + 1) + 3];
    float2  a[kFwdMaxCar][kFbTileM + 4];                      // rotated 400 kS/s samples m0-4 .. m0+199
    float2  RW[kFwdMaxCar][kFbMPerBit];                       // grouped response tables (kFbFastLen)
    float2  JW[kFwdMaxCar][kFbMPerBit];
    float   I3[kFwdMaxCar][3][8][kFbMPerBit];
    float2  Wq[kFwdMaxCar][kFbTileBits + 1];                  // mixer phasor at the start of bits q0-1 .. q0+4
    uint32_t bval[kFwdMaxCar];                                // bit k = value of bit q0-10+k (0 where muted)
    uint32_t bmute[kFwdMaxCar];                               // bit k = bit q0-10+k is muted (0xFF)
};

size_t fwd_bits_smem_bytes() { return sizeof(FwdBitsSmem); }

__global__ void __launch_bounds__(kFwdThreads, 3) fwd_bits_kernel(const __grid_constant__ FwdBitsParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FwdBitsSmem *sm = reinterpret_cast<FwdBitsSmem *>(smem_raw);
    const int t = threadIdx.x;
    const uint32_t ntiles = (p.nbits + kFbTileBits - 1) / kFbTileBits;
    for (int i = t; i < p.ncar * kFbFastLen; i += kFwdThreads) {
        const int c = i / kFbFastLen, k = i - c * kFbFastLen;
        const float v = p.fast[i];
        if (k < 2 * kFbMPerBit) (&sm->RW[c][0].x)[k] = v;
        else if (k < 4 * kFbMPerBit) (&sm->JW[c][0].x)[k - 2 * kFbMPerBit] = v;
        else (&sm->I3[c][0][0][0])[k - 4 * kFbMPerBit] = v;
    }

    // the bits a tile needs (15 per carrier: q0-10 .. q0+4) are fetched one tile ahead by lanes 0..14 of warp c, so their
    // global-load latency hides behind a tile of work; a ballot turns them into a value mask and a mute mask per carrier
    constexpr int kNB = kFbTileBits + kFbHistBits + 1;             // 15
    const int fc = t >> 5, fk = t & 31;
    auto fetch_bit = [&](uint32_t tile_) -> uint8_t {
        if (fc >= p.ncar || fk >= kNB || tile_ >= ntiles) return 0xFF;
        const long q = (long)tile_ * kFbTileBits - (kFbHistBits + 1) + fk;
        if (q >= (long)p.nbits) return 0xFF;
        if (q >= 0) return p.bits[fc][q];
        if (q >= -(long)kFbHistBits) return p.hbits[fc][kFbHistBits + q];
        return 0xFF;
    };
    uint8_t bit_next = fetch_bit(blockIdx.x);

    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long q0 = (long)tile * kFbTileBits;                  // first bit of the tile (call-local)
        const int nvalid = (int)((long)p.nbits - q0 < kFbTileBits ? (long)p.nbits - q0 : kFbTileBits);
        // ---- bit masks and the per-bit mixer phasors
        if (fc < kFwdMaxCar) {                                      // warps 0..2, whole warps: ballots are convergent
            const uint32_t val = __ballot_sync(0xffffffffu, bit_next == 1);
            const uint32_t mut = __ballot_sync(0xffffffffu, bit_next > 1);
            if (fk == 0 && fc < p.ncar) { sm->bval[fc] = val & 0x7FFFu; sm->bmute[fc] = mut & 0x7FFFu; }
        } else if (t >= 96 && t < 96 + p.ncar * (kFbTileBits + 1)) {
            const int idx = t - 96, c = idx / (kFbTileBits + 1), k = idx - c * (kFbTileBits + 1);
            const uint32_t qabs = (uint32_t)(p.bit_base + (unsigned long long)(q0 - 1 + k));
            sm->Wq[c][k] = sincos_phase(qabs * p.fcw_mix1000[c]);
        }
        bit_next = fetch_bit(tile + gridDim.x);
        __syncthreads();

        // ---- phase A: 400 kS/s samples by table lookup, rotated by the carrier NCO.  thread = one sample m0-4+t
        if (t < kFbTileM + 4) {
            const int mrel = t - 4;                                 // relative to the tile's first sample
            const int qrel = mrel >= 0 ? mrel / kFbMPerBit : -1;     // bit containing it, relative to q0
            const int u0 = mrel - qrel * kFbMPerBit;                // 0..39
            const int k0 = qrel + kFbHistBits + 1;                  // its index in the tile's bit window (9..14)
#pragma unroll
            for (int c = 0; c < kFwdMaxCar; ++c) {
                if (c < p.ncar) {
                    // the path is chosen from the sample's OWN nine bits, so a sample's value does not depend on where the
                    // tile (i.e. the call) boundaries fall: bit-identical under any chunking
                    const uint32_t mute = (sm->bmute[c] >> (k0 - 8)) & 0x1FFu;
                    float2 v = make_float2(0.f, 0.f);
                    if (mute == 0u) {
                        // nothing muted: real part is data-independent, imaginary part three bits per lookup.
                        // window bit i <-> response slot d = 8 - i
                        const uint32_t w = (sm->bval[c] >> (k0 - 8)) & 0x1FFu;
                        float im = sm->I3[c][0][w & 7u][u0];
                        im = __fadd_rn(im, sm->I3[c][1][(w >> 3) & 7u][u0]);
                        im = __fadd_rn(im, sm->I3[c][2][(w >> 6) & 7u][u0]);
                        v = cmul(fma2(splat(im), sm->JW[c][u0], sm->RW[c][u0]), sm->Wq[c][qrel + 1]);
                    } else if (mute != 0x1FFu) {
                        // a mute transition inside the window (stream start, fvc/audio switch-over): per-bit walk
                        const uint32_t val = sm->bval[c], mu = sm->bmute[c];
                        float2 acc = make_float2(0.f, 0.f);
                        const float2 *R = p.resp + (size_t)c * 2 * kFbRespLen;
#pragma unroll 1
                        for (int d = 0; d < kFbRespBits; ++d) {
                            const int k = k0 - d;
                            if (!((mu >> k) & 1u)) acc = add2(acc, __ldg(&R[((val >> k) & 1u) * kFbRespLen + u0 + kFbMPerBit * d]));
                        }
                        v = cmul(acc, cmul(sm->Wq[c][qrel + 1], p.w40[c][u0]));
                    }
                    sm->a[c][t] = v;
                }
            }
        }
        __syncthreads();

        // ---- phase 3a: x5 CIC^3 to 2 MS/s with folded mixers, carriers summed (as in fwd_fused_kernel)
        if (t < kFbTileM + 1) {
            float2 acc[5];
#pragma unroll
            for (int r = 0; r < 5; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < kFwdMaxCar; ++c) {
                if (c < p.ncar) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const float2 x = sm->a[c][t + 3 - j];
                        const float2 xr = splat(x.x), xi = splat(x.y);
#pragma unroll
                        for (int r = 0; r < 5; ++r) {
                            if (r + 5 * j < 13) {
                                const float2 C = p.C1[c][r + 5 * j];
                                acc[r] = fma2(xi, make_float2(-C.y, C.x), fma2(xr, C, acc[r]));   // (the ready-made j C1 table of the general kernel
                                                                                               //  measured 1.6 % slower here)
                            }
                        }
                    }
                }
            }
            float2 *Bo = &sm->B[5 * t];
#pragma unroll
            for (int r = 0; r < 5; ++r) Bo[r] = acc[r];
        }
        if (t == 0) tma_store_wait_read0();
        __syncthreads();

        // ---- phase 3b: shared x5 CIC^3 to 10 MS/s
        if (t < kFbMPerBit * nvalid) {
            const float2 *Bi = &sm->B[5 * (t + 1)];
            float2 b[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) b[k] = Bi[k - 2];
            float2 *o = &sm->out[25 * t];
#pragma unroll
            for (int qq = 0; qq < 5; ++qq)
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    float2 v = mul2(splat(p.G2[r]), b[qq + 2]);
                    v = fma2(splat(p.G2[r + 5]), b[qq + 1], v);
                    if (r + 10 < 13) v = fma2(splat(p.G2[r + 10]), b[qq], v);
                    o[5 * qq + r] = v;
                }
        }
        fence_async_smem();
        __syncthreads();
        if (t == 0) tma_store_1d(p.out + (size_t)q0 * kFbOutPerBit, sm->out, (uint32_t)nvalid * kFbOutPerBit * (uint32_t)sizeof(float2));
    }
    if (t == 0) tma_store_wait_all();
}

cudaError_t launch_fwd_bits(const FwdBitsParams &p, int grid, cudaStream_t st) {
    if (p.nbits == 0) return cudaSuccess;
    fwd_bits_kernel<<<grid, kFwdThreads, sizeof(FwdBitsSmem), st>>>(p);
    return cudaGetLastError();
}

cudaError_t fwd_configure_device() {
    cudaError_t e = cudaFuncSetAttribute(fwd_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdBitsSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fwd_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(fwd_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_voice_smem_bytes());
}

cudaError_t launch_fwd_fused(const FwdParams &p, int grid, cudaStream_t st) {
    if (p.nsym == 0) return cudaSuccess;
    fwd_fused_kernel<false><<<grid, kFwdThreads, sizeof(FwdSmem), st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fwd_fused_voice(const FwdParams &p, int grid, cudaStream_t st) {
    if (p.nsym == 0) return cudaSuccess;
    fwd_fused_kernel<true><<<grid, kFwdThreads, fwd_voice_smem_bytes(), st>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// voice pre-pass (16 kS/s: 1/625 of the output rate, so these kernels are off the roofline).
//   x[n] = audio[n] + SAT;  y[n] = sum_k g[k] x[n - k]   (fm_preemph's IIR as its truncated impulse response: no
//   recurrence, any sample computable on its own, float64);  delta[n] = frac(y[n] max_dev / fs) * 2^64
//   Phi[n] = Phi[n-1] + delta[n] (mod 2^64, an exact integer scan in two levels);  phasor[n] = e^{j 2 pi Phi[n] / 2^64}
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) voice_delta_kernel(const __grid_constant__ VoicePrepParams p) {
    __shared__ double s_x[256 + kVoiceImp];
    __shared__ double s_g[kVoiceImp];
    __shared__ unsigned long long s_warp[8];
    const int t = threadIdx.x;
    const long n0 = (long)blockIdx.x * 256;
    for (int i = t; i < kVoiceImp; i += 256) s_g[i] = p.g[i];
    for (int i = t; i < 256 + kVoiceImp; i += 256) {
        const long n = n0 - kVoiceImp + i;
        double x = 0.0;
        if (n >= (long)p.n_audio) x = 0.0;
        else if (n >= 0) x = (double)p.audio[n] + p.sat[(p.a_base + (unsigned long long)n) & 7ull];
        else x = p.hx_old[kVoiceImp + n];
        s_x[i] = x;
    }
    __syncthreads();
    const long n = n0 + t;
    unsigned long long d = 0;
    if (n < (long)p.n_audio) {
        double y = 0.0;
        for (int k = 0; k < kVoiceImp; ++k) y = fma(s_g[k], s_x[kVoiceImp + t - k], y);
        double cyc = y * p.cycles_per_unit;
        cyc -= floor(cyc);
        d = __double2ull_rn(cyc * 18446744073709551616.0);
    }
    // inclusive scan of the phase steps inside the block (mod 2^64: exact), block total for the second level
    unsigned long long incl = d;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, s);
        if ((t & 31) >= s) incl += o;
    }
    if ((t & 31) == 31) s_warp[t >> 5] = incl;
    __syncthreads();
    unsigned long long woff = 0;
    for (int w = 0; w < (t >> 5); ++w) woff += s_warp[w];
    if (n < (long)p.n_audio) p.delta[n] = woff + incl;
    if (t == 255) p.btot[blockIdx.x] = woff + incl;
    // history of inputs for the next call: the last kVoiceImp of (old history ++ this call's inputs)
    if (blockIdx.x == 0 && t < kVoiceImp) {
        const long src = (long)p.n_audio - kVoiceImp + t;
        double x;
        if (src >= 0) x = (double)p.audio[src] + p.sat[(p.a_base + (unsigned long long)src) & 7ull];
        else x = p.hx_old[kVoiceImp + src];
        p.hx_new[t] = x;
    }
}

// one CTA: exclusive scan of the block totals on top of the running phase; the running phase for the next call
__global__ void __launch_bounds__(1024) voice_totals_kernel(const __grid_constant__ VoicePrepParams p) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const int t = threadIdx.x;
    const uint32_t nblk = (p.n_audio + 255u) / 256u;
    if (t == 0) s_carry = *p.phase;
    __syncthreads();
    for (uint32_t base = 0; base < nblk; base += 1024) {
        const uint32_t b = base + t;
        const unsigned long long v = b < nblk ? p.btot[b] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, s);
            if ((t & 31) >= s) incl += o;
        }
        if ((t & 31) == 31) s_warp[t >> 5] = incl;
        __syncthreads();
        unsigned long long woff = 0;
        for (int w = 0; w < (t >> 5); ++w) woff += s_warp[w];
        const unsigned long long carry = s_carry;
        if (b < nblk) p.boff[b] = carry + woff + incl - v;
        __syncthreads();
        if (t == 1023) s_carry = carry + woff + incl;
        __syncthreads();
    }
    if (t == 0) *p.phase = s_carry;
}

__global__ void __launch_bounds__(256) voice_phasor_kernel(const __grid_constant__ VoicePrepParams p) {
    const int t = threadIdx.x;
    const long n = (long)blockIdx.x * 256 + t;
    const long h0 = (long)p.n_audio - kFwdVoiceHist;               // samples >= h0 are the next call's history
    if (n < (long)p.n_audio) {
        const unsigned long long phi = p.boff[blockIdx.x] + p.delta[n];
        const float2 v = sincos_phase((uint32_t)(phi >> 32));
#pragma unroll
        for (int l = 0; l < kFwdVoiceLegs; ++l) {
            if (!p.vph[l]) continue;
            const float2 w = p.leg_muted[l] ? make_float2(0.f, 0.f) : v;
            p.vph[l][n] = w;
            if (n >= h0) p.vhist_new[l][n - h0] = w;
        }
    }
    if (blockIdx.x == 0 && t < kFwdVoiceHist && h0 + t < 0) {      // fewer than 32 new samples: the rest comes from the old history
#pragma unroll
        for (int l = 0; l < kFwdVoiceLegs; ++l)
            if (p.vph[l]) p.vhist_new[l][t] = p.vhist_old[l][kFwdVoiceHist + h0 + t];
    }
}

cudaError_t launch_voice_prep(const VoicePrepParams &p, cudaStream_t st) {
    if (p.n_audio == 0) return cudaSuccess;
    const unsigned int nblk = (p.n_audio + 255u) / 256u;
    voice_delta_kernel<<<nblk, 256, 0, st>>>(p);
    voice_totals_kernel<<<1, 1024, 0, st>>>(p);
    voice_phasor_kernel<<<nblk, 256, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace amps
