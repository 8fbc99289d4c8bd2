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

namespace amps {

static_assert(sizeof(amps_burst) % 8 == 0, "burst records are streamed to the host ring in 8-byte words");

// ============================================================================================
// front end
// ============================================================================================
struct FrontSmem {
    float2   in[kStages][kTile];          // TMA landing ring
    float2   v[kPorch + kVRing];          // 400 kS/s samples: porch mirrors the ring tail
    float2   pb[2][2][kTB];               // [tile parity][P1|P2][block] rotated CIC partial sums
    float2   y[2][kTB + 1];               // [pass parity][1 + thread] 200 kS/s baseband
    uint64_t full[kStages];
};

size_t rx_front_smem_bytes() { return sizeof(FrontSmem); }

__device__ __forceinline__ void issue_tile(const RxFrontParams &p, FrontSmem *sm, long tile, int stage) {
    // tiles with a negative index come from the history buffer (one pass = 2 tiles long)
    const float2 *src = tile < 0 ? p.tail + (long)kHist + tile * (long)kTile : p.chunk + tile * (long)kTile;
    mbar_expect_tx(&sm->full[stage], kTile * (uint32_t)sizeof(float2));
    tma_load_1d(sm->in[stage], src, kTile * (uint32_t)sizeof(float2), &sm->full[stage]);
}

// y[q] = sum_k h2[k] v[2q-k] as two interleaved FFMA2 chains (even taps, odd taps), k ascending.
// vq points at v[2q]; pairs (v[2q-2j], v[2q-2j+1]) are fetched with one 128-bit shared load.
__device__ __forceinline__ float2 channel_filter(const RxFrontParams &p, const float2 *vq) {
    float2 E = make_float2(0.f, 0.f), O = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 150; ++j) {
        float4 pr = *reinterpret_cast<const float4 *>(vq - 2 * j);
        E = fma2(splat(p.h2[2 * j]), make_float2(pr.x, pr.y), E);
        if (j >= 1) O = fma2(splat(p.h2[2 * j - 1]), make_float2(pr.z, pr.w), O);
    }
    return add2(E, O);
}

__global__ void __launch_bounds__(kTB, 2) rx_front_kernel(const __grid_constant__ RxFrontParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FrontSmem *sm = reinterpret_cast<FrontSmem *>(smem_raw);
    const int t = threadIdx.x;

    const uint32_t pa = (uint32_t)(((uint64_t)p.npass * blockIdx.x) / gridDim.x);
    const uint32_t pb = (uint32_t)(((uint64_t)p.npass * (blockIdx.x + 1)) / gridDim.x);
    if (pa == pb) return;
    const int  ntiles = 2 * (int)(pb - pa) + 2;          // one warm-up pass + the CTA's own passes
    const long tile0  = 2 * (long)pa - 2;

    if (t == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&sm->full[s], 1);
        mbar_fence_init();
    }
    // partial sums "before the first tile": only ever feed warm-up outputs that are discarded
    sm->pb[1][0][t] = make_float2(0.f, 0.f);
    sm->pb[1][1][t] = make_float2(0.f, 0.f);
    __syncthreads();
    if (t == 0) {
        for (int s = 0; s < kStages && s < ntiles; ++s) issue_tile(p, sm, tile0 + s, s);
    }

    float2 *vorg = sm->v + kPorch;

    for (int i = 0; i < ntiles; ++i) {
        const int s = i % kStages;
        mbar_wait(&sm->full[s], (uint32_t)(i / kStages) & 1u);

        // ---- stage 1: NCO rotate + CIC^3 polyphase partial sums over this thread's 25 samples
        const float2 *xin = sm->in[s] + kD1 * t;
        float2 P0 = make_float2(0.f, 0.f), P1 = P0, P2 = P0;
#pragma unroll
        for (int k = 0; k < kD1; ++k) {
            float2 u = cmul(xin[k], p.w[k]);
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
        const int par = i & 1;
        sm->pb[par][0][t] = P1;
        sm->pb[par][1][t] = P2;
        __syncthreads();                                   // partials visible; in[s] fully consumed
        if (t == 0 && i + kStages < ntiles) issue_tile(p, sm, tile0 + i + kStages, s);

        // ---- v[m] = (P0[m] + P1[m-1]) + P2[m-2]
        const float2 q1 = t >= 1 ? sm->pb[par][0][t - 1] : sm->pb[par ^ 1][0][kTB - 1];
        const float2 q2 = t >= 2 ? sm->pb[par][1][t - 2] : sm->pb[par ^ 1][1][kTB - 2 + t];
        const float2 v  = add2(add2(P0, q1), q2);
        const int base = ((i >> 1) & 1) * 2 * kTB;
        const int r    = base + (i & 1) * kTB + t;
        vorg[r] = v;
        if (r >= kVRing - kPorch) vorg[r - kVRing] = v;

        if (i & 1) {
            // End of a pass.  The warm-up pass (i == 1) only produces y[q0-1], the predecessor the
            // quadrature demod of the CTA's first real output needs; nothing is written for it.
            const bool warm = (i == 1);
            __syncthreads();                               // the pass's 2*TB new v samples are in place
            // ---- stage 2: 299-tap channel filter /2, one output per thread
            const int yb = (i >> 1) & 1;
            float2 y = make_float2(0.f, 0.f);
            if (!warm || t == kTB - 1) {
                y = channel_filter(p, vorg + base + 2 * t);
                sm->y[yb][t + 1] = y;
            }
            __syncthreads();
            if (!warm) {
                // ---- quadrature demod: arg(y[q] * conj(y[q-1]))
                const float2 yp = t == 0 ? sm->y[yb ^ 1][kTB] : sm->y[yb][t];
                const float zr = __fmaf_rn(y.y, yp.y, __fmul_rn(y.x, yp.x));
                const float zi = __fmaf_rn(y.y, yp.x, -__fmul_rn(y.x, yp.y));
                const float d  = atan2_spec(zi, zr);
                const uint32_t    lp = pa + (uint32_t)(i >> 1) - 1u;          // pass index within this call
                const unsigned long long ql = (unsigned long long)lp * kTB + t;
                p.dring[(p.q_base + ql) & p.dmask] = d;
                if (p.ydump) p.ydump[ql] = y;
            }
        }
    }
}

cudaError_t launch_rx_front(const RxFrontParams &p, int grid, cudaStream_t st) {
    rx_front_kernel<<<grid, kTB, rx_front_smem_bytes(), st>>>(p);
    return cudaGetLastError();
}

// ============================================================================================
// trigger detection
// ============================================================================================
// 74 half-symbols: Manchester("10" x 13 + "11100010010"), bit 0 -> (1,0), bit 1 -> (0,1)
// (lib/recc_impl.cc:51-65,76).
__device__ __constant__ uint8_t c_trig[kTrig] = {
    0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,
    0,1,0,1,0,1,1,0,1,0,1,0,0,1,1,0,1,0,0,1,1,0};

__global__ void __launch_bounds__(256) rx_detect_kernel(const float *__restrict__ dring, uint32_t dmask, RxState *state,
                                                       Candidate *cand, unsigned long long scan_lo, unsigned long long scan_hi) {
    const unsigned long long i = scan_lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= scan_hi) return;
    if (i < state->lo || i < state->resume_at) return;
    // early-out hard match: after k compares a random position survives with probability 2^-k
#pragma unroll 1
    for (int k = 0; k < kTrig; ++k) {
        const float v = dring[(i + (unsigned long long)(kOS * k)) & dmask];
        if ((v >= 0.0f) != (c_trig[k] != 0)) return;
    }
    float c = 0.0f;
#pragma unroll 1
    for (int k = 0; k < kTrig; ++k) {
        const float v = dring[(i + (unsigned long long)(kOS * k)) & dmask];
        c = __fadd_rn(c, c_trig[k] ? v : -v);
    }
    const unsigned int slot = atomicAdd(&state->ncand, 1u);
    if (slot < (unsigned)kMaxCand) {
        cand[slot].pos = i;
        cand[slot].corr = c;
    } else {
        atomicAdd(&state->cand_overflow, 1u);
    }
}

cudaError_t launch_rx_detect(const float *dring, uint32_t dmask, RxState *state, Candidate *cand,
                             unsigned long long scan_lo, unsigned long long scan_hi, cudaStream_t st) {
    if (scan_hi <= scan_lo) return cudaSuccess;
    unsigned long long n = scan_hi - scan_lo;
    unsigned int grid = (unsigned int)((n + 255) / 256);
    rx_detect_kernel<<<grid, 256, 0, st>>>(dring, dmask, state, cand, scan_lo, scan_hi);
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

__device__ __forceinline__ unsigned gf_mul(unsigned a, unsigned b) {
    return (a && b) ? c_gf.exp[c_gf.log[a] + c_gf.log[b]] : 0u;
}

// Validity of one 48-bit RECC word repeat as itpp::BCH(63,2,true)::decode reports it for
// (15 zeros || 48 bits) (lib/recc_decode_impl.cc:53-79): <=2 errors anywhere in the 63 bits, plus
// the S1 == 0, S3 a non-zero cube case that Berlekamp's 2-step iteration turns into a degree-3
// locator with three roots.
__device__ bool bch48_valid(const uint8_t *bits48) {
    unsigned s1 = 0, s3 = 0;
    for (int b = 0; b < 48; ++b) {
        if (bits48[b] & 1u) {
            const int e = 47 - b;                       // exponent of x carried by this bit
            s1 ^= c_gf.exp[e];
            s3 ^= c_gf.exp[(3 * e) % 63];
        }
    }
    if (s1 == 0) return s3 == 0 || (c_gf.log[s3] % 3u) == 0;
    const unsigned s1c = gf_mul(gf_mul(s1, s1), s1);
    if (s3 == s1c) return true;                         // single error
    // Lambda(x) = 1 + s1 x + ((s3 + s1^3)/s1) x^2 must have two roots among alpha^0..alpha^62
    const unsigned num = s3 ^ s1c;
    const unsigned c2  = c_gf.exp[c_gf.log[num] + 63 - c_gf.log[s1]];
    int roots = 0;
    for (int j = 0; j < 63; ++j) {
        const unsigned x  = c_gf.exp[j];
        const unsigned x2 = c_gf.exp[(2 * j) % 63];
        if ((1u ^ gf_mul(s1, x) ^ gf_mul(c2, x2)) == 0u) ++roots;
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
    const int t = threadIdx.x, nt = blockDim.x;
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
    if (t < 35) scratch_valid[t] = bch48_valid(&out->words[t / 5][48 * (t % 5)]) ? 1 : 0;
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
// candidate selection + capture + decode (single CTA; candidates are rare)
// ============================================================================================
__global__ void __launch_bounds__(256) rx_select_kernel(const float *__restrict__ dring, uint32_t dmask, RxState *state,
                                                       Candidate *cand, unsigned long long scan_hi, amps_burst *scratch,
                                                       amps_burst *host_ring, unsigned int ring_len, RxPublished *host_pub) {
    extern __shared__ unsigned char sel_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(sel_raw);          // kMaxCand
    float *corr = reinterpret_cast<float *>(keys + kMaxCand);                            // kMaxCand
    __shared__ unsigned long long acc_pos[kMaxAccept];
    __shared__ float acc_corr[kMaxAccept];
    __shared__ unsigned int acc_run[kMaxAccept];
    __shared__ unsigned int n_acc;
    __shared__ uint8_t s_valid[40];
    __shared__ unsigned int s_errs[8];
    __shared__ unsigned long long rec_base;

    const int t = threadIdx.x, nt = blockDim.x;
    unsigned int n = state->ncand;
    if (n > (unsigned)kMaxCand) n = kMaxCand;
    // pad to a power of two and bitonic-sort by position (atomics made the order arbitrary)
    unsigned int np = 1;
    while (np < n) np <<= 1;
    for (unsigned int i = t; i < np; i += nt) {
        keys[i] = i < n ? cand[i].pos : ~0ull;
        corr[i] = i < n ? cand[i].corr : 0.0f;
    }
    __syncthreads();
    for (unsigned int k = 2; k <= np; k <<= 1) {
        for (unsigned int j = k >> 1; j > 0; j >>= 1) {
            for (unsigned int i = t; i < np; i += nt) {
                const unsigned int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long a = keys[i], b = keys[ixj];
                    if ((a > b) == up) {
                        keys[i] = b; keys[ixj] = a;
                        const float ca = corr[i]; corr[i] = corr[ixj]; corr[ixj] = ca;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (t == 0) {
        unsigned long long lo = state->lo, resume = state->resume_at;
        unsigned long long new_lo = scan_hi > lo ? scan_hi : lo;
        unsigned int na = 0;
        unsigned int i = 0;
        while (i < n) {
            if (keys[i] < lo || keys[i] < resume) { ++i; continue; }
            // run of adjacent sampling phases that all matched
            unsigned int j = i;
            unsigned int best = i;
            while (j + 1 < n && keys[j + 1] == keys[j] + 1) {
                ++j;
                if (corr[j] > corr[best]) best = j;
            }
            if (keys[j] + 1 >= scan_hi) {                  // run may continue past the searched range: decide next call
                new_lo = keys[i];
                break;
            }
            if (na < (unsigned)kMaxAccept) {
                acc_pos[na] = keys[best];
                acc_corr[na] = corr[best];
                acc_run[na] = j - i + 1;
                ++na;
            }
            resume = keys[best] + (unsigned long long)kBurstLen;
            i = j + 1;
        }
        state->lo = new_lo;
        state->resume_at = resume;
        state->ncand = 0;
        n_acc = na;
        rec_base = state->nrec_total;
    }
    __syncthreads();
    for (unsigned int a = 0; a < n_acc; ++a) {
        amps_burst *rec = &scratch[a];
        const unsigned long long pos = acc_pos[a];
        for (int s = t; s < kCapture; s += nt) {
            const float v = dring[(pos + (unsigned long long)(kOS * (kTrig + s))) & dmask];
            rec->symbols[s] = v >= 0.0f ? 1 : 0;
        }
        if (t == 0) {
            rec->demod_index = pos;
            rec->sample_index = pos * (unsigned long long)(kD1 * kD2);
            rec->corr = acc_corr[a];
            rec->run_length = acc_run[a];
            rec->pad[0] = 0; rec->pad[1] = 0;
        }
        __syncthreads();
        decode_burst_block(rec->symbols, &rec->decoded, s_valid, s_errs);
        // publish: stream the finished record into the host-visible ring (posted PCIe writes)
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(rec);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(&host_ring[(rec_base + a) % ring_len]);
        for (int w = t; w < (int)(sizeof(amps_burst) / 8); w += nt) dst[w] = src[w];
    }
    __syncthreads();
    if (t == 0) {
        state->nrec_total = rec_base + n_acc;
        __threadfence_system();
        host_pub->cand_overflow = state->cand_overflow;
        host_pub->nrec_total = rec_base + n_acc;
    }
}

cudaError_t launch_rx_select(const float *dring, uint32_t dmask, RxState *state, Candidate *cand,
                             unsigned long long scan_hi, amps_burst *scratch, amps_burst *host_ring, unsigned int ring_len,
                             RxPublished *host_pub, cudaStream_t st) {
    const size_t smem = (size_t)kMaxCand * (sizeof(unsigned long long) + sizeof(float));
    rx_select_kernel<<<1, 256, smem, st>>>(dring, dmask, state, cand, scan_hi, scratch, host_ring, ring_len, host_pub);
    return cudaGetLastError();
}

// per-device opt-in to large dynamic shared memory (call once per device after cudaSetDevice)
cudaError_t rx_configure_device() {
    cudaError_t e = cudaFuncSetAttribute(rx_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rx_front_smem_bytes());
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(rx_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)((size_t)kMaxCand * (sizeof(unsigned long long) + sizeof(float))));
}

}  // namespace amps
