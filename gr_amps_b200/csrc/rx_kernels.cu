// rx_kernels.cu -- fused RECC receive path for sm_100a.
//
//   rx_front_kernel   : IQ @10 MS/s --TMA--> smem -> NCO rotate -> CIC^3 /25 -> 299-tap channel filter /2 (the reference's
//                       lpf_taps @400 kS/s) -> quadrature demod -> d[] @200 kS/s and its hard decisions, THEN, in the same
//                       launch, the trigger search on what it just wrote (exact 74/74 match of sign(d) at every sampling
//                       phase + soft correlation of the matches: the recc_impl.cc:118 memmem, at 10 phases) and, by the last
//                       CTA of each channel, the choice of the bursts to capture.  Replaces freq_xlating_fir_filter_ccc +
//                       quadrature_demod_cf (grc/ampsbs.grc:1814-1872, 774-816), the 25x front-end decimation the 10 MS/s
//                       configs need and the trigger half of amps.recc (lib/recc_impl.cc:115-119).
//                       One launch serves one channel or a batch of independent channels.
//   rx_capture_kernel : 3374-symbol capture at the chosen phase (recc_impl.cc:124-126), Manchester decode + BCH validity +
//                       field parse (recc_decode_impl.cc:81-169), record streamed into the pinned host ring.
//   rx_front400_kernel / rx_search_kernel : the same chain at the reference's own 400 kS/s (no CIC stage).
//
// No tensor cores: there is no dense contraction on this path.  The front kernel is a persistent streaming kernel: the
// launch's tiles (3 units of 1600 samples) are dealt evenly to the CTAs, each CTA keeps the filter history of its segment
// in shared memory and re-reads only two warm-up tiles in front of it.
#include "rx_kernels.cuh"
#include "recc_compat.cuh"

namespace amps {

static_assert(sizeof(amps_burst) % 8 == 0, "burst records are streamed to the host ring in 8-byte words");

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
    return v;
}
// measurement aid: time stamp `idx` of this CTA (thread 0 only)
#define RX_PROF(p, idx) do { if ((p).prof && threadIdx.x == 0) (p).prof[blockIdx.x * 16 + (idx)] = global_ns(); } while (0)

// ============================================================================================
// trigger search (device routines shared by rx_front_kernel and rx_search_kernel)
// ============================================================================================
// 74 half-symbols: Manchester("10" x 13 + "11100010010"), bit 0 -> (1,0), bit 1 -> (0,1)
// (lib/recc_impl.cc:51-65,76).
//   0,1,1,0 x 13 | 0,1,0,1,0,1,1,0,1,0,1,0,0,1,1,0,1,0,0,1,1,0
// 74-symbol trigger packed LSB-first (symbol k = bit k)
__device__ __constant__ uint32_t c_trig_bits[3] = {0x66666666u, 0x56A66666u, 0x00000196u};
__host__ __device__ constexpr uint32_t trig_bit(int k) {                             // (the same, folded where k is a constant)
    return ((k < 32 ? 0x66666666u : k < 64 ? 0x56A66666u : 0x00000196u) >> (k & 31)) & 1u;
}

// The rings are read with ld.global.cg: inside rx_front_kernel the words were written during this very launch, some of them
// by other SMs, and an L1 line fetched earlier by a co-resident CTA may predate them.
__device__ __forceinline__ uint32_t hard_window(const uint32_t *__restrict__ hring, uint32_t wmask, unsigned long long b) {
    const uint32_t wi = (uint32_t)(b >> 5);
    const uint32_t w0 = __ldcg(&hring[wi & wmask]), w1 = __ldcg(&hring[(wi + 1) & wmask]);
    return __funnelshift_r(w0, w1, (uint32_t)b & 31u);
}

// Exact 74/74 hard match for the 32 adjacent sampling positions i0 .. i0+31 (i0 a multiple of 32):
// for half-symbol k the 32 hard decisions at positions i0+10k .. i0+10k+31 are one 32-bit window of
// the bit ring, so one AND per symbol tests all 32 positions; a random group dies after ~6 symbols.
// The first 8 windows are fetched together (one round trip to L2 decides 7 groups out of 8), the others 22 at a time.
__device__ __forceinline__ uint32_t trig_and(uint32_t m, int k, uint32_t win) {
    return m & (((c_trig_bits[k >> 5] >> (k & 31)) & 1u) ? win : ~win);
}
__device__ __forceinline__ uint32_t group_match(const uint32_t *__restrict__ hring, uint32_t wmask, unsigned long long i0) {
    uint32_t win[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) win[k] = hard_window(hring, wmask, i0 + (unsigned long long)(kOS * k));
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 8; ++k) m &= ((0x66u >> k) & 1u) ? win[k] : ~win[k];      // c_trig_bits[0] & 0xff
#pragma unroll 1
    for (int k0 = 8; k0 < kTrig && m; k0 += 22) {         // 8 + 3 x 22 = 74
        uint32_t w[22];
#pragma unroll
        for (int j = 0; j < 22; ++j) w[j] = hard_window(hring, wmask, i0 + (unsigned long long)(kOS * (k0 + j)));
#pragma unroll
        for (int j = 0; j < 22; ++j) m = trig_and(m, k0 + j, w[j]);
    }
    return m;
}

// The same test on the CTA's own recent decisions, kept in a shared-memory ring of kHwRing words (word index = absolute
// demod index / 32): what rx_front_kernel runs after every pass, so that in a long segment the search costs nothing.
constexpr int kHwRing = 128;
__device__ __forceinline__ uint32_t group_match_smem(const uint32_t *hw, unsigned long long i0) {
    uint32_t m = 0xffffffffu;
#pragma unroll 1
    for (int k = 0; k < kTrig && m; ++k) {
        const unsigned long long b = i0 + (unsigned long long)(kOS * k);
        const uint32_t wi = (uint32_t)(b >> 5);
        m = trig_and(m, k, __funnelshift_r(hw[wi & (kHwRing - 1)], hw[(wi + 1) & (kHwRing - 1)], (uint32_t)b & 31u));
    }
    return m;
}

// Shared-memory scratch of the search (one per CTA).
constexpr int kSearchSpan = 32 + 10 + kOS * (kTrig - 1);      // demod samples a group's correlations can touch: 772
constexpr int kLocalCand = 16;
struct SearchScratch {
    float     d[kSearchSpan + 4];
    float     corr[48];
    uint32_t  m0s[256];
    uint32_t  wb[8];
    uint32_t  mprev, mnext;
    uint32_t  nlc;                        // candidates found by this CTA and not yet appended to the channel's list
    uint32_t  pad;
    Candidate lc[kLocalCand];
};

// append the CTA's local candidates to the channel's list (thread 0; one atomic for all of them)
__device__ __forceinline__ void flush_candidates(RxState *state, Candidate *cand, SearchScratch *sc) {
    const uint32_t n = sc->nlc;
    if (n == 0u) return;
    const unsigned int slot = atomicAdd(&state->ncand, n);
    for (uint32_t i = 0; i < n; ++i) {
        if (slot + i < (unsigned)kMaxCand) cand[slot + i] = sc->lc[i];
        else atomicAdd(&state->cand_overflow, 1u);
    }
    sc->nlc = 0u;
}

// One group with at least one match (rare: about one per burst), all threads of the CTA: the neighbours' masks delimit
// the runs, the demod samples under the group are staged once, one thread per sampling position sums its soft correlation
// (sum_k +/- d[i + 10k], k ascending), and the first position of every run yields one candidate for the whole run: its
// soft-correlation peak (first maximum) is the sampling phase (a run is at most 10 long: the pattern cannot match one
// half-symbol later).  Reads the stream up to i0 + 793.  hw != nullptr: the neighbours are in the CTA's own decision ring.
__device__ __noinline__ void resolve_group(const float *__restrict__ dring, const uint32_t *__restrict__ hring, uint32_t dmask, RxState *state,
                                           Candidate *cand, unsigned long long i0, uint32_t m0, SearchScratch *sc, const uint32_t *hw) {
    const int t = threadIdx.x, nt = blockDim.x;
    const uint32_t wmask = dmask >> 5;
    if (t == 0) sc->mprev = i0 >= 32 ? (hw ? group_match_smem(hw, i0 - 32) : group_match(hring, wmask, i0 - 32)) : 0u;
    if (t == 32) sc->mnext = hw ? group_match_smem(hw, i0 + 32) : group_match(hring, wmask, i0 + 32);
    for (int j = t; j < kSearchSpan; j += nt) sc->d[j] = __ldcg(&dring[(i0 + (unsigned long long)j) & dmask]);
    __syncthreads();
    const unsigned long long M = (unsigned long long)m0 | ((unsigned long long)sc->mnext << 32);
    if (t < 42 && ((M >> t) & 1ull)) {
        float c = 0.0f;
#pragma unroll
        for (int k = 0; k < kTrig; ++k) { const float v = sc->d[t + kOS * k]; c = __fadd_rn(c, trig_bit(k) ? v : -v); }
        sc->corr[t] = c;
    }
    __syncthreads();
    if (t == 0) {
        uint32_t starts = m0 & ~((m0 << 1) | (sc->mprev >> 31));
        while (starts) {
            const int bit = __ffs(starts) - 1;
            starts &= starts - 1;
            int best = bit;
            float bestc = sc->corr[bit];
            unsigned int run = 1;
            while (run < 11u && ((M >> (bit + run)) & 1ull)) {
                const float c = sc->corr[bit + run];
                if (c > bestc) { bestc = c; best = bit + (int)run; }
                ++run;
            }
            if (sc->nlc == (uint32_t)kLocalCand) flush_candidates(state, cand, sc);
            Candidate &o = sc->lc[sc->nlc++];
            o.start = i0 + (unsigned long long)bit;
            o.best = i0 + (unsigned long long)best;
            o.corr = bestc;
            o.run = run;
        }
    }
    __syncthreads();
}

// Second half of a search round, all threads of the CTA: thread t found the match mask m0 (mostly 0) for group base + t;
// the CTA resolves the groups that matched.
__device__ __noinline__ void resolve_round(const float *__restrict__ dring, const uint32_t *__restrict__ hring, uint32_t dmask, RxState *state,
                                           Candidate *cand, unsigned long long base, uint32_t m0, SearchScratch *sc, const uint32_t *hw) {
    const int t = threadIdx.x, nt = blockDim.x;
    sc->m0s[t] = m0;
    const uint32_t bal = __ballot_sync(0xffffffffu, m0 != 0u);
    if ((t & 31) == 0) sc->wb[t >> 5] = bal;
    __syncthreads();
    for (int w = 0; w < nt / 32; ++w) {
        uint32_t bits = sc->wb[w];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            resolve_group(dring, hring, dmask, state, cand, 32ull * (base + (unsigned long long)(32 * w + b)), sc->m0s[32 * w + b], sc, hw);
        }
    }
    __syncthreads();
}

// Groups [g_lo, g_hi) from the rings in global memory, all threads of the CTA (uniform arguments): one group per thread and
// round for the hard match, then the CTA resolves the groups that matched (candidates go to the CTA's local list: the caller
// flushes it).  Every group is searched exactly once over the life of a stream.
__device__ __noinline__ void search_groups(const float *__restrict__ dring, const uint32_t *__restrict__ hring, uint32_t dmask, RxState *state,
                              Candidate *cand, unsigned long long g_lo, unsigned long long g_hi, SearchScratch *sc) {
    const int t = threadIdx.x, nt = blockDim.x;
    for (unsigned long long base = g_lo; base < g_hi; base += (unsigned long long)nt) {
        const unsigned long long g = base + (unsigned long long)t;
        const uint32_t m0 = g < g_hi ? group_match(hring, dmask >> 5, 32ull * g) : 0u;
        if (__syncthreads_or(m0 != 0u)) resolve_round(dring, hring, dmask, state, cand, base, m0, sc, nullptr);
    }
}

// Candidate selection, all threads of the CTA (candidates are rare: about one per burst).  Sort the list by run start,
// then walk it in stream order: a run is DECIDED once the stream reaches kSpan beyond its end (the whole capture is in the
// ring) -- the first undecided run and everything after it stay on the list for a later call; a decided run that starts
// inside an already captured burst is dropped; the others are accepted and the search resumes (74+3374)*10 positions after
// the sampling position (oracle/dsp_chain.c orc_rx_detect).  The outcome does not depend on how the stream was cut into calls.
__device__ __noinline__ void select_channel(RxState *state, Candidate *cand, Candidate *sorted, Accepted *acc, uint32_t par,
                                            unsigned long long total_d, RxPublished *host_pub, Candidate *scratch, unsigned int scratch_cap) {
    const int t = threadIdx.x, nt = blockDim.x;
    unsigned int n = __ldcg(&state->ncand);
    if (n > (unsigned)kMaxCand) n = kMaxCand;
    if (2u * n <= scratch_cap) {
        // the usual case: the whole list fits in shared memory (unsorted copy | sorted copy)
        for (unsigned int i = t; i < n; i += nt) {
            Candidate c;
            c.start = __ldcg(&cand[i].start); c.best = __ldcg(&cand[i].best); c.corr = __ldcg(&cand[i].corr); c.run = __ldcg(&cand[i].run);
            scratch[i] = c;
        }
        __syncthreads();
        for (unsigned int i = t; i < n; i += nt) {
            const Candidate c = scratch[i];
            unsigned int rank = 0;
            for (unsigned int j = 0; j < n; ++j) rank += scratch[j].start < c.start ? 1u : 0u;
            scratch[n + rank] = c;
        }
        sorted = scratch + n;
    } else {
        for (unsigned int i = t; i < n; i += nt) {
            Candidate c;
            c.start = __ldcg(&cand[i].start); c.best = __ldcg(&cand[i].best); c.corr = __ldcg(&cand[i].corr); c.run = __ldcg(&cand[i].run);
            unsigned int rank = 0;
            for (unsigned int j = 0; j < n; ++j) rank += __ldcg(&cand[j].start) < c.start ? 1u : 0u;
            sorted[rank] = c;
        }
    }
    __syncthreads();
    if (t == 0) {
        unsigned long long resume = state->resume_at;
        unsigned int na = 0, keep = 0;
        for (unsigned int i = 0; i < n; ++i) {
            const Candidate c = sorted[i];
            const bool decided = c.start + c.run + (unsigned long long)kSpan < total_d;
            if (!decided || na == (unsigned)kMaxAccept) {             // (or no room left in this call's list: next call)
                for (unsigned int j = i; j < n; ++j) cand[keep++] = sorted[j];
                break;
            }
            if (c.start < resume) continue;                // begins inside a burst that was already captured
            acc[na].pos = c.best;
            acc[na].corr = c.corr;
            acc[na].run = c.run;
            ++na;
            resume = c.best + (unsigned long long)kBurstLen;
        }
        state->resume_at = resume;
        state->ncand = keep;
        state->n_acc[par] = na;
        state->rec_base[par] = state->nrec_total;
        state->nrec_total += na;
        if (state->cand_overflow != state->pub_overflow) {     // (rare; the capture of an earlier call may be running on its own
            state->pub_overflow = state->cand_overflow;        //  stream right now: in this mode only the selection publishes this)
            __threadfence_system();
            host_pub->cand_overflow = state->cand_overflow;
        }
    }
    __syncthreads();
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
struct PassOut {                          // where a pass's outputs go
    float    *dring;
    uint32_t *hring;
    float2   *ydump;
    uint32_t *hw;                         // shared-memory ring of the newest decisions (kHwRing words) or nullptr
    uint32_t  dmask;
};
// input sample formats: fc32 (gr_complex, what the reference's flowgraph carries) and sc16 (interleaved int16 I/Q, what
// the USRP puts on the wire before UHD's host-side conversion, grc/ampsbs.grc:3750): x = (float)int16 * in_scale
// kUnitScale: the scale is a power of two and has been folded into the NCO tables on the host -- (I s) w and I (s w) are the
// same real number when s is a power of two, so the result is bit-identical and the two multiplies per sample go away.
template <bool kUnitScale> __device__ __forceinline__ float2 to_c32(float2 v, float) { return v; }
// (float)int16 without the conversion unit (I2F runs on the XU pipe at 1/8 of the FMA rate: 50 of them per thread and tile were
// a third of the sc16 kernel's time): 0x4B000000 | (x + 32768) is the float 2^23 + (x + 32768), exactly; subtracting 2^23 + 32768
// is exact as well.  Two integer ops and one packed add per sample, the same value bit for bit.
__device__ __forceinline__ float2 s16x2_to_f32x2(short2 v) {
    const uint32_t w = *reinterpret_cast<const uint32_t *>(&v) ^ 0x80008000u;           // x + 32768 in each half, as unsigned
    const float2 m = make_float2(__uint_as_float(0x4B000000u | (w & 0xFFFFu)), __uint_as_float(0x4B000000u | (w >> 16)));
    return __fadd2_rn(m, make_float2(-8421376.0f, -8421376.0f));
}
template <bool kUnitScale> __device__ __forceinline__ float2 to_c32(short2 v, float s) {
    const float2 f = s16x2_to_f32x2(v);
    if (kUnitScale) return f;
    return make_float2(__fmul_rn(f.x, s), __fmul_rn(f.y, s));
}

constexpr int kMaxBound = 16;           // a CTA's output is read by at most ceil(26 / 3) + 1 = 10 boundaries
template <typename In>
struct FrontSmem {
    In       in[kStages][kTile];          // TMA landing ring
    PassSmem ps;
    float2   pb[3][2][kTB];               // [tile % 3][P1|P2][block] rotated CIC partial sums (3 buffers: a tile reads its
                                          // own and the previous tile's, the next tile may already be writing)
    uint64_t full[kStages];
    uint32_t is_last;                     // this CTA is the last one of the channel to finish
    uint32_t pad;
};
// what only the instance with the trigger search inside needs: it follows FrontSmem in the dynamic shared memory of that
// instance alone, so that next to two CTAs of the default instance an SM has room for a search AND a capture CTA
struct FusedSmem {
    uint32_t hw[kHwRing];                 // the segment's newest hard decisions (trigger search without a trip to L2)
    uint32_t nb;                          // boundaries this CTA contributes to: whose (CTA index), the counter value that
    uint32_t bj[kMaxBound];               // means "everybody else has been here", and whether this CTA was the last to arrive
    uint32_t bneed[kMaxBound];
    uint32_t bmine[kMaxBound];
    uint32_t pad;
    SearchScratch sc;
};
template <typename In, bool kFused>
constexpr size_t front_smem_bytes() { return sizeof(FrontSmem<In>) + (kFused ? sizeof(FusedSmem) : 0); }

size_t rx_front_smem_bytes() { return front_smem_bytes<float2, false>(); }

__host__ __device__ constexpr int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// y[q] = sum_k h2[k] v[2q-k] for the kR outputs q = kR*c0 + r, each as two FFMA2 chains (even taps,
// odd taps), k ascending -- the order the oracle uses.  vcol points at column c0 of row 0.
__device__ __forceinline__ void channel_filter(const float (&h2)[300], const float4 *vcol, float2 (&y)[kR]) {
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
            if (j >= 0 && j <= 149) E[r] = fma2(splat(h2[2 * j]), make_float2(pr.x, pr.y), E[r]);
            if (j >= 1 && j <= 149) O[r] = fma2(splat(h2[2 * j - 1]), make_float2(pr.z, pr.w), O[r]);
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
// hard decisions, and the history shuffle for the next pass.  Threads t < nact own valid outputs (nact = kTB for a full
// pass, 8 per unit for the partial pass that ends a segment).  With warm == true it only produces y[q0-1], the predecessor
// the demod of the segment's first real output needs; nothing is written for it.
//   q_pass: absolute demod index of the pass's first output; ql_pass: the same relative to the call's first output.
__device__ __forceinline__ void finish_pass(const float (&h2)[300], const PassOut &o, PassSmem *ps, bool warm, int pc,
                                            unsigned long long q_pass, unsigned long long ql_pass, int nact, int t) {
    __syncthreads();                                   // the pass's new v samples are in place
    float2 y[kR];
#pragma unroll
    for (int r = 0; r < kR; ++r) y[r] = make_float2(0.f, 0.f);
    // (warm-up: one thread computes the kR outputs in front of the segment and keeps the last.  A one-output variant of the
    // filter would be 4x shorter, but as a second unrolled block it more than doubles the kernel's registers -- a CTA of the
    // search / capture kernels then no longer fits next to two of these -- and as a real call it reads the taps through
    // generic loads: measured 4 us slower per segment.)
    if (warm ? t == 0 : t < nact) {
        const int c0 = warm ? -1 : t;
        channel_filter(h2, &ps->v[0][c0 + kPorchCols], y);
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
        const unsigned long long qabs = q_pass + (unsigned long long)kR * t;
        const bool act = t < nact;
        if (act) *reinterpret_cast<float4 *>(&o.dring[qabs & o.dmask]) = make_float4(d[0], d[1], d[2], d[3]);
        // hard decisions (binary_slicer_fb: x >= 0 -> 1), 32 per word: 8 lanes x 4 outputs
        unsigned int hb = 0;
#pragma unroll
        for (int r = 0; r < kR; ++r) hb |= (d[r] >= 0.0f ? 1u : 0u) << r;
        hb <<= 4 * (t & 7);
        hb |= __shfl_xor_sync(0xffffffffu, hb, 1);
        hb |= __shfl_xor_sync(0xffffffffu, hb, 2);
        hb |= __shfl_xor_sync(0xffffffffu, hb, 4);
        if (act && (t & 7) == 0) {
            o.hring[(qabs & o.dmask) >> 5] = hb;
            if (o.hw) o.hw[(uint32_t)(qabs >> 5) & (kHwRing - 1)] = hb;
        }
        if (o.ydump && act) {
            const unsigned long long ql = ql_pass + (unsigned long long)kR * t;
#pragma unroll
            for (int r = 0; r < kR; ++r) o.ydump[ql + r] = y[r];
        }
    }
}

// ---- how a launch's tiles are dealt: the Tt tiles are split evenly, CTA s owns tiles [Tt*s/grid, Tt*(s+1)/grid) -- its
// SEGMENTS (one per channel the range touches).  (A pool of late segments for CTAs that finish early was tried on the
// 2^28-sample batch: the spread between CTAs is not removed by it and the extra warm-up tiles cost 2 %.)
// (deal_lo() / deal_owner(): rx_kernels.cuh -- the host-side tests walk them too)

// samples [L0, L0 + n) of a channel's logical stream -> shared memory; logical samples below `carry` live in the tail
// buffer (history + what the previous call could not use), the rest in this call's chunk; a tile may straddle the seam
template <typename In>
__device__ __forceinline__ void issue_tile(const RxChan &ch, In *dst, uint64_t *bar, long L0, uint32_t n) {
    mbar_expect_tx(bar, n * (uint32_t)sizeof(In));
    const long carry = (long)ch.carry;
    const In *tail = static_cast<const In *>(ch.tail) + (long)kHist;      // logical sample 0 of the tail buffer
    const In *chunk = static_cast<const In *>(ch.chunk) - carry;          // logical sample 0 of the chunk (virtual)
    if (L0 + (long)n <= carry) tma_load_1d(dst, tail + L0, n * (uint32_t)sizeof(In), bar);
    else if (L0 >= carry) tma_load_1d(dst, chunk + L0, n * (uint32_t)sizeof(In), bar);
    else {
        const uint32_t n1 = (uint32_t)(carry - L0);
        tma_load_1d(dst, tail + L0, n1 * (uint32_t)sizeof(In), bar);
        tma_load_1d(dst + n1, chunk + carry, (n - n1) * (uint32_t)sizeof(In), bar);
    }
}

// the groups a segment [qs, qe) of demod samples has to search, and the first one that needs no other CTA's output
struct SegGroups { unsigned long long lo, mid, hi; };
__device__ __forceinline__ SegGroups seg_groups(unsigned long long qs, unsigned long long qe, bool first_of_call) {
    SegGroups g;
    g.lo = qs / kUnitOut >= (unsigned)kGroupLag ? qs / kUnitOut - kGroupLag : 0ull;
    g.hi = qe / kUnitOut >= (unsigned)kGroupLag ? qe / kUnitOut - kGroupLag : 0ull;
    g.mid = g.lo;
    if (!first_of_call) {                       // groups below qs/32 + 1 look at samples in front of qs (written in this launch)
        g.mid = qs / kUnitOut + 1;
        if (g.mid > g.hi) g.mid = g.hi;
        if (g.mid < g.lo) g.mid = g.lo;
    }
    return g;
}

// next call's tail = logical samples [units*kUnit - kHist, carry + nchunk): the CTAs that touch the channel copy a slice each
// (saves a memcpy node between consecutive front kernels)
template <typename In>
__device__ __forceinline__ void copy_tail(const RxChan &ch, const RxDeal &d, uint32_t cta, uint32_t cbase, uint32_t cend, int t) {
    const uint32_t k0 = deal_owner(d, cbase), nsl = deal_owner(d, cend - 1) - k0 + 1u;
    const long first = (long)ch.units * kUnit - kHist;
    const uint32_t len = (uint32_t)((long)ch.carry + (long)ch.nchunk - first);
    const uint32_t per = (len + nsl - 1u) / nsl;
    const uint32_t lo = (cta - k0) * per, hi = lo + per < len ? lo + per : len;
    const In *tl = static_cast<const In *>(ch.tail) + (long)kHist;
    const In *ck = static_cast<const In *>(ch.chunk) - (long)ch.carry;
    In *dst = static_cast<In *>(ch.tail_out);
    // (eight loads in flight per thread: with few CTAs on a channel a slice is a dozen rounds of DRAM latency otherwise)
#ifndef AMPS_TAIL_BATCH
#define AMPS_TAIL_BATCH 8
#endif
    constexpr int kB = AMPS_TAIL_BATCH;
    for (uint32_t i0 = lo + (uint32_t)t; i0 < hi; i0 += kB * kTB) {
        In v[kB];
#pragma unroll
        for (int j = 0; j < kB; ++j) {
            const uint32_t i = i0 + (uint32_t)(j * kTB);
            const long L = first + (long)i;
            if (i < hi) v[j] = L < (long)ch.carry ? __ldg(tl + L) : __ldg(ck + L);
        }
#pragma unroll
        for (int j = 0; j < kB; ++j) {
            const uint32_t i = i0 + (uint32_t)(j * kTB);
            if (i < hi) dst[i] = v[j];
        }
    }
}

// One segment = tiles [ta, tb) of channel c (tile j = units 3j .. 3j+2, the channel's last tile may be shorter), preceded by
// kWarmTiles tiles of history.  `it` counts the tiles this CTA has consumed (TMA ring position / mbarrier phase).
template <typename In, bool kUnitScale, int kMaxChan, bool kFused>
__device__ __forceinline__ void run_segment(const RxFrontParamsT<kMaxChan> &p, const RxChan &ch, FrontSmem<In> *sm,
                                            uint32_t ta, uint32_t tb, uint32_t &it, int t, const RxDeal &dl, uint32_t cta,
                                            uint32_t cbase, uint32_t cend) {
    FusedSmem *fs = reinterpret_cast<FusedSmem *>(sm + 1);               // (exists in the kFused instance only)
    const int  ntiles = (int)(tb - ta) + kWarmTiles;
    const long tj0    = (long)ta - kWarmTiles;                          // channel tile index of tile i = 0
    const uint32_t U  = ch.units;
    auto tile_blocks = [&](int i) -> uint32_t {                         // valid 25-sample blocks of tile i
        const long tj = tj0 + i;
        if (tj < 0) return (uint32_t)kTB;
        const uint32_t left = (U - (uint32_t)kTileUnits * (uint32_t)tj) * (uint32_t)kUnitBlk;
        return left < (uint32_t)kTB ? left : (uint32_t)kTB;
    };
    const In *chunk0 = static_cast<const In *>(ch.chunk) - (long)ch.carry;      // logical sample 0 of the chunk (virtual)
    PassOut o;
    o.dring = ch.dring; o.hring = ch.hring; o.ydump = ch.ydump; o.dmask = ch.dmask; o.hw = (kFused && ch.search) ? fs->hw : nullptr;
    const uint32_t useg = ((uint32_t)kTileUnits * tb < U ? (uint32_t)kTileUnits * tb : U) - (uint32_t)kTileUnits * ta;   // units of the segment
    const unsigned long long ql0 = (unsigned long long)kUnitOut * kTileUnits * ta;     // first output, relative to the call
    // first group that reads nothing in front of the segment (the groups below it are searched in finish_segment)
    unsigned long long g_next = (ch.q_base + ql0) / kUnitOut + 1ull;

    // partial sums "before the first tile": they only reach samples the warm-up never uses
    sm->pb[2][0][t] = make_float2(0.f, 0.f);
    sm->pb[2][1][t] = make_float2(0.f, 0.f);
    __syncthreads();                                     // (also: the previous segment is done with ps / pb / in)
    if (t == 0) {
        for (int s = 0; s < kStages && s < ntiles; ++s)
            issue_tile<In>(ch, sm->in[(it + s) % kStages], &sm->full[(it + s) % kStages], (tj0 + s) * (long)kTile, tile_blocks(s) * kD1);
    }
    // (with the first tiles on their way) this CTA's slice of the next call's history
    copy_tail<In>(ch, dl, cta, cbase, cend, t);

    for (int i = 0; i < ntiles; ++i, ++it) {
        const int s = (int)(it % kStages);
        mbar_wait(&sm->full[s], (it / kStages) & 1u);
#ifdef AMPS_RX_PROF_TILES                                  // (two compares per tile: only in a build made for tools/front_phases.py)
        if (i == 0) RX_PROF(p, 1);
        if (i == kWarmTiles) RX_PROF(p, 2);
#endif

        // ---- stage 1: NCO rotate + CIC^3 polyphase partial sums over this thread's 25 samples
        const In *xin = sm->in[s] + kD1 * t;
        float2 P0 = make_float2(0.f, 0.f), P1 = P0, P2 = P0;
#pragma unroll
        for (int k = 0; k < kD1; ++k) {
            const float2 x = to_c32<kUnitScale>(xin[k], ch.in_scale);
            const float2 w = ch.w[k];
            float2 wj;
            if constexpr (kMaxChan == 1) wj = p.wj0[k]; else wj = make_float2(-w.y, w.x);
            const float2 u = fma2(splat(x.y), wj, mul2(splat(x.x), w));      // = cmul(x, w[k]), same operations
            P0 = fma2(splat(p.g[24 - k]), u, P0);
            P1 = fma2(splat(p.g[49 - k]), u, P1);
            if (74 - k < kNCic) P2 = fma2(splat(p.g[74 - k]), u, P2);
        }
        // absolute block index (mod 2^32): blocks of logical sample 0 + (tile index relative to it) * kTB + t
        const uint32_t babs = ch.blk_base + (uint32_t)((int)tj0 + i) * (uint32_t)kTB + (uint32_t)t;
        const float2   W    = sincos_phase(babs * ch.fcw25);
        P0 = cmul(P0, W);
        P1 = cmul(P1, W);
        P2 = cmul(P2, W);
        const int par = i % 3, prv = (i + 2) % 3;
        sm->pb[par][0][t] = P1;
        sm->pb[par][1][t] = P2;
        __syncthreads();                                   // partials visible; in[s] fully consumed
        if (t == 0 && i + kStages < ntiles) {
            // the common case by the shortest route (this thread's warp is the one every tile waits for): a whole tile that lies
            // in the chunk; the seam with the tail buffer and the segment's last (possibly short) tile take the general one
            const int j = i + kStages;
            const long L0 = (tj0 + j) * (long)kTile;
            if (L0 >= (long)ch.carry && j != ntiles - 1) {
                mbar_expect_tx(&sm->full[s], kTile * (uint32_t)sizeof(In));
                tma_load_1d(sm->in[s], chunk0 + L0, kTile * (uint32_t)sizeof(In), &sm->full[s]);
            } else {
                issue_tile<In>(ch, sm->in[s], &sm->full[s], L0, tile_blocks(j) * kD1);
            }
        }

        // ---- v[m] = (P0[m] + P1[m-1]) + P2[m-2], stored into the pair/row layout
        const float2 q1 = t >= 1 ? sm->pb[par][0][t - 1] : sm->pb[prv][0][kTB - 1];
        const float2 q2 = t >= 2 ? sm->pb[par][1][t - 2] : sm->pb[prv][1][kTB - 2 + t];
        const float2 v  = add2(add2(P0, q1), q2);
        const bool warm = i < kWarmTiles;
        const int  u    = warm ? 0 : (i - kWarmTiles) % kPassTiles;          // tile index inside the pass
        store_v(&sm->ps, warm ? (i - kWarmTiles) * kTB + t : u * kTB + t, v);

        const bool pass_end = !warm && (u == kPassTiles - 1 || i == ntiles - 1);
        if (pass_end || i == kWarmTiles - 1) {
            const int pc = warm ? -1 : (i - kWarmTiles) / kPassTiles;
            int nact = kTB;
            if (!warm) {
                const uint32_t left = useg - (uint32_t)(kPassTiles * kTileUnits) * (uint32_t)pc;          // units from this pass on
                if (left < (uint32_t)(kPassTiles * kTileUnits)) nact = (int)left * (kUnitOut / kR);
            }
            const unsigned long long ql = ql0 + (unsigned long long)(pc < 0 ? 0 : pc) * kPassOut;
            finish_pass(p.h2, o, &sm->ps, warm, pc, ch.q_base + ql, ql, nact, t);
            if (kFused && !warm && ch.search) {
                // ---- trigger search on the segment's own output: the groups whose lookahead this pass completed
                const unsigned long long q_done = ch.q_base + ql + (unsigned long long)(kR * nact);
                const unsigned long long g_end = q_done / kUnitOut >= (unsigned)kGroupLag ? q_done / kUnitOut - kGroupLag : 0ull;
                if (g_end > g_next) {                      // (at most 24 groups: one pass)
                    __syncthreads();                       // the pass's decisions are in the ring (and its demod samples in L2)
                    const unsigned long long g = g_next + (unsigned long long)t;
                    const uint32_t m0 = g < g_end ? group_match_smem(fs->hw, 32ull * g) : 0u;
                    if (__syncthreads_or(m0 != 0u)) resolve_round(ch.dring, ch.hring, ch.dmask, ch.state, ch.cand, g_next, m0, &fs->sc, fs->hw);
                    g_next = g_end;
                }
            }
        }
    }
}

// Segment k's part of a channel that owns global tiles [cbase, cend): tiles [lo, hi) relative to the channel
struct SegTiles { uint32_t lo, hi; };
__device__ __forceinline__ SegTiles seg_tiles(const RxDeal &d, uint32_t k, uint32_t cbase, uint32_t cend) {
    uint32_t lo = deal_lo(d, k), hi = deal_lo(d, k + 1);
    if (lo < cbase) lo = cbase;
    if (hi > cend) hi = cend;
    SegTiles s;
    s.lo = lo - cbase;
    s.hi = hi > lo ? hi - cbase : s.lo;
    return s;
}
// first segment whose output the boundary groups of a segment starting at channel tile ta (> 0) read: they look back
// kGroupLag + 1 units from the segment's first output
__device__ __forceinline__ uint32_t boundary_first(const RxDeal &d, uint32_t cbase, uint32_t ta) {
    const uint32_t u0 = (uint32_t)kTileUnits * ta;
    const uint32_t ulo = u0 > (uint32_t)(kGroupLag + 1) ? u0 - (uint32_t)(kGroupLag + 1) : 0u;
    return deal_owner(d, cbase + ulo / (uint32_t)kTileUnits);
}

// After a segment: search the groups whose lookahead ends inside it, and if this CTA is the last of the channel to get
// here, select the bursts.
//   Groups near the front of a segment (its BOUNDARY groups) look at demod samples that other CTAs of this launch produce:
// the segments from boundary_first() up to the segment itself.  Nobody waits for anybody: whoever finishes one of those
// segments bumps the boundary's counter once its outputs are globally visible, and whoever arrives last searches the boundary
// -- by then everything it reads is in place.  The counters reset themselves; the order in which segments run is irrelevant.
template <typename In, int kMaxChan, bool kFused>
__device__ __forceinline__ void finish_segment(const RxFrontParamsT<kMaxChan> &p, const RxChan &ch, FrontSmem<In> *sm, uint32_t c,
                                               uint32_t seg, uint32_t ta, uint32_t tb, int t) {
    const RxDeal &dl = p.deal;
    const uint32_t cbase = p.tile_cum[c], cend = p.tile_cum[c + 1];
    const uint32_t U = ch.units;
    const unsigned long long qs = ch.q_base + (unsigned long long)kUnitOut * kTileUnits * ta;
    const unsigned long long qe = ch.q_base + (unsigned long long)kUnitOut * ((uint32_t)kTileUnits * tb < U ? (uint32_t)kTileUnits * tb : U);
    const unsigned long long total_d = ch.q_base + (unsigned long long)kUnitOut * U;
    if constexpr (!kFused) { RX_PROF(p, 3); RX_PROF(p, 7); return; }      // search and selection are a launch of their own
    FusedSmem *fs = reinterpret_cast<FusedSmem *>(sm + 1);
    RxState *state = ch.state;
    const uint32_t k1 = deal_owner(dl, cend - 1);             // last segment of the channel
    __syncthreads();                                      // this CTA's demod samples and decisions are visible to all its threads
    RX_PROF(p, 3);
    if (ch.search) {
        // ---- which boundaries does this CTA contribute to?  Its own (if it has one) and those of the CTAs right behind it.
        if (t == 0) {
            __threadfence();                              // my outputs before my counter bumps
            uint32_t n = 0;
            for (uint32_t j = seg; j <= k1 && n < (uint32_t)kMaxBound; ++j) {
                const SegTiles sj = seg_tiles(dl, j, cbase, cend);
                if (sj.lo == 0u) continue;                // the channel's first segment of the call: nothing in front of it in this launch
                const uint32_t first = boundary_first(dl, cbase, sj.lo);
                if (first > seg) break;                   // segment j (and everything behind it) does not read my output
                fs->bj[n] = j;
                fs->bneed[n] = j - first;                 // counter value the last arrival sees
                ++n;
            }
            fs->nb = n;
        }
        __syncthreads();
        const uint32_t nb = fs->nb;
        if ((uint32_t)t < nb) {
            const uint32_t j = fs->bj[t];
            const uint32_t prev = atomicAdd(&ch.flags[j], 1u);
            const bool mine = prev == fs->bneed[t];
            if (mine) ch.flags[j] = 0u;                   // everybody has been here: ready for the next launch
            fs->bmine[t] = mine ? 1u : 0u;
        }
        // (the groups that read only this segment's own output were searched pass by pass in run_segment)
        RX_PROF(p, 4);
        if (ta == 0) {
            // ---- the channel's first segment of the call: its front groups read the previous calls' output, which is in place
            const SegGroups g = seg_groups(qs, qe, false);
            search_groups(ch.dring, ch.hring, ch.dmask, state, ch.cand, g.lo, g.mid, &fs->sc);
        }
        // ---- the boundaries this CTA was the last to reach
        __syncthreads();
        for (uint32_t i = 0; i < nb; ++i) {
            if (!fs->bmine[i]) continue;
            __threadfence();
            const SegTiles sj = seg_tiles(dl, fs->bj[i], cbase, cend);
            const unsigned long long js = ch.q_base + (unsigned long long)kUnitOut * kTileUnits * sj.lo;
            const uint32_t ju = (uint32_t)kTileUnits * sj.hi;
            const unsigned long long je = ch.q_base + (unsigned long long)kUnitOut * (ju < U ? ju : U);
            const SegGroups gj = seg_groups(js, je, false);
            search_groups(ch.dring, ch.hring, ch.dmask, state, ch.cand, gj.lo, gj.mid, &fs->sc);
        }
    }
    __syncthreads();
    RX_PROF(p, 5);
    if (t == 0) {
        if (ch.search) flush_candidates(state, ch.cand, &fs->sc);
        __threadfence();
        const uint32_t n_touch = k1 - deal_owner(dl, cbase) + 1u;
        const unsigned int prev = atomicAdd(&state->front_done, 1u);
        sm->is_last = prev + 1u == n_touch ? 1u : 0u;
    }
    __syncthreads();
    RX_PROF(p, 6);
    if (sm->is_last) {
        __threadfence();
        if (ch.search)
            select_channel(state, ch.cand, ch.cand + kMaxCand, ch.acc + (size_t)ch.par * kMaxAccept, ch.par, total_d, ch.host_pub,
                           reinterpret_cast<Candidate *>(sm->in), (unsigned int)(sizeof(sm->in) / sizeof(Candidate)));
        if (t == 0) state->front_done = 0u;
        RX_PROF(p, 8);
    }
    RX_PROF(p, 7);
}

template <typename In, int kMinCtas, bool kUnitScale, int kMaxChan, bool kFused>
__global__ void __launch_bounds__(kTB, kMinCtas) rx_front_kernel(const __grid_constant__ RxFrontParamsT<kMaxChan> p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FrontSmem<In> *sm = reinterpret_cast<FrontSmem<In> *>(smem_raw);
    const int t = threadIdx.x;
    const uint32_t cta = blockIdx.x;                     // == its first (static) segment; gridDim.x == p.deal.nstat
    const RxDeal &dl = p.deal;

    RX_PROF(p, 0);
    if (t == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&sm->full[s], 1);
        mbar_fence_init();
        if constexpr (kFused) reinterpret_cast<FusedSmem *>(sm + 1)->sc.nlc = 0u;
    }
    uint32_t it = 0;
    if constexpr (kMaxChan == 1) {
        // one channel: everything about the channel is a compile-time offset into the parameters
        const RxChan &ch = p.ch[0];
        const uint32_t lo = deal_lo(dl, cta), hi = deal_lo(dl, cta + 1);
        run_segment<In, kUnitScale, kMaxChan, kFused>(p, ch, sm, lo, hi, it, t, dl, cta, 0u, dl.Tt);
        finish_segment<In, kMaxChan, kFused>(p, ch, sm, 0u, cta, lo, hi, t);
    } else {
        const uint32_t gt_lo = deal_lo(dl, cta), gt_hi = deal_lo(dl, cta + 1);
        uint32_t c = 0;
        for (uint32_t g0 = gt_lo; g0 < gt_hi;) {
            while (p.tile_cum[c + 1] <= g0) ++c;
            const RxChan &ch = p.ch[c];
            const uint32_t cbase = p.tile_cum[c], cend = p.tile_cum[c + 1];
            const uint32_t g1 = gt_hi < cend ? gt_hi : cend;
            run_segment<In, kUnitScale, kMaxChan, kFused>(p, ch, sm, g0 - cbase, g1 - cbase, it, t, dl, cta, cbase, cend);
            finish_segment<In, kMaxChan, kFused>(p, ch, sm, c, cta, g0 - cbase, g1 - cbase, t);
            g0 = g1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// native-rate front end: complex IQ @400 kS/s (the reference's own operating point, grc/ampsbs.grc:263).
// No decimating first stage: v[m] = x[m] e^{-j theta m}, then exactly freq_xlating_fir_filter_ccc's filter /2 and
// quadrature_demod_cf.  At 0.4 MS/s real time this kernel is never a bottleneck (it is FFMA-bound: 150 FFMA2 per
// 8-byte sample); it exists so that recc_iq drops into the reference flowgraph at its native rate.
// ---------------------------------------------------------------------------------------------
template <typename In, bool kUnitScale>
__global__ void __launch_bounds__(kTB, 4) rx_front400_kernel(const __grid_constant__ RxFront400Params p) {
    __shared__ PassSmem ps;
    const int t = threadIdx.x;
    uint32_t pa = blockIdx.x * p.pass_per_cta;
    if (pa >= p.npass) return;
    uint32_t pb = pa + p.pass_per_cta;
    if (pb > p.npass) pb = p.npass;
    PassOut o;
    o.dring = p.dring; o.hring = p.hring; o.ydump = p.ydump; o.dmask = p.dmask; o.hw = nullptr;
    auto load = [&](long L) -> float2 {                   // logical sample L of this call; L < 0 = history
        const float2 x = to_c32<kUnitScale>(L < 0 ? static_cast<const In *>(p.tail)[(long)kPass400 + L] : static_cast<const In *>(p.chunk)[L], p.in_scale);
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
    finish_pass(p.h2, o, &ps, true, -1, 0ull, 0ull, kTB, t);
    for (uint32_t pc = 0; pc < pb - pa; ++pc) {
        const long base = first + (long)pc * kPass400;
        __syncthreads();                                   // the history shuffle of the previous pass is done
#pragma unroll
        for (int u = 0; u < kPassTiles; ++u) store_v(&ps, u * kTB + t, load(base + u * kTB + t));
        const unsigned long long ql = (unsigned long long)(pa + pc) * kPassOut;
        finish_pass(p.h2, o, &ps, false, (int)pc, p.q_base + ql, ql, kTB, t);
    }
}

cudaError_t launch_rx_front400(const RxFront400Params &p, int grid, cudaStream_t st, bool sc16, bool unit) {
    if (sc16 && unit) rx_front400_kernel<short2, true><<<grid, kTB, 0, st>>>(p);
    else if (sc16) rx_front400_kernel<short2, false><<<grid, kTB, 0, st>>>(p);
    else rx_front400_kernel<float2, false><<<grid, kTB, 0, st>>>(p);
    return cudaGetLastError();
}

// sc16 tiles are half the size, so three CTAs fit an SM (the kernel is no longer HBM-bound at 4 B/sample)
constexpr int kSc16Ctas = 3;
// fc32: two CTAs fit an SM (shared memory); the bound is 3 to hold the kernel to 112 registers, so that a CTA of the search /
// capture kernels of the previous call finds room (registers as well as shared memory) next to them
constexpr int kFc32Regs = 2;
template <int kMaxChan, bool kFused, typename P>
static cudaError_t launch_front_t(const P &p, int grid, cudaStream_t st, bool sc16, bool unit) {
    if (sc16 && unit) rx_front_kernel<short2, kSc16Ctas, true, kMaxChan, kFused><<<grid, kTB, front_smem_bytes<short2, kFused>(), st>>>(p);
    else if (sc16) rx_front_kernel<short2, kSc16Ctas, false, kMaxChan, kFused><<<grid, kTB, front_smem_bytes<short2, kFused>(), st>>>(p);
    else rx_front_kernel<float2, kFc32Regs, false, kMaxChan, kFused><<<grid, kTB, front_smem_bytes<float2, kFused>(), st>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_rx_front(const RxFrontParams1 &p, int grid, cudaStream_t st, bool sc16, bool unit, bool fused) {
    return fused ? launch_front_t<1, true>(p, grid, st, sc16, unit) : launch_front_t<1, false>(p, grid, st, sc16, unit);
}
cudaError_t launch_rx_front_batch(const RxFrontParamsB &p, int grid, cudaStream_t st, bool sc16, bool unit, bool fused) {
    return fused ? launch_front_t<kMaxBatch, true>(p, grid, st, sc16, unit) : launch_front_t<kMaxBatch, false>(p, grid, st, sc16, unit);
}
int rx_front_ctas_per_sm(bool sc16) { return sc16 ? kSc16Ctas : 2; }

uint32_t rx_make_deal(RxDeal &d, uint32_t tiles, uint32_t resident, uint32_t nchan, uint32_t equal_tiles) {
    uint32_t grid = resident < (uint32_t)kMaxGrid ? resident : (uint32_t)kMaxGrid;
    if (grid > tiles) grid = tiles;
    if (grid < 1u) grid = 1u;
    d.Tt = tiles; d.nstat = grid; d.P = 0; d.Tc = 0;
    if (nchan > 1u && equal_tiles > 0u && nchan <= grid) {
        uint32_t P = grid / nchan;                           // whole CTAs per channel
        if (P > equal_tiles) P = equal_tiles;
        if (4u * nchan * P >= 3u * grid) { d.P = P; d.Tc = equal_tiles; d.nstat = nchan * P; return d.nstat; }    // >= 3/4 of the CTAs busy
    }
    return grid;
}

// ---------------------------------------------------------------------------------------------
// One accepted burst, all threads of the CTA: gather the 3374 half-symbols at the chosen phase (or take the blob amps.recc
// cut in M&M mode), decode, and stream the finished record into the host-visible ring (posted PCIe writes).  Ends with the
// record on its way (__threadfence_system + barrier); publishing the count is the caller's business.
// ---------------------------------------------------------------------------------------------
__device__ void capture_burst(const float *__restrict__ dring, uint32_t dmask, const Accepted *acc, const uint8_t *blobs,
                              const unsigned long long *blob_sym_index, uint32_t decim, amps_burst *host_ring, uint32_t ring_len,
                              unsigned long long rec_base, unsigned int b, unsigned int n_acc, amps_burst *rec, uint8_t *s_valid,
                              unsigned int *s_errs) {
    const int t = threadIdx.x, nt = blockDim.x;
    if (blobs) {
        for (int s = t; s < kCapture; s += nt) rec->symbols[s] = blobs[(size_t)b * kCapture + s];
        if (t == 0) {
            // position bookkeeping is nominal here: the recovered half-symbol index, 10 demod samples per half-symbol
            const unsigned long long first_sym = blob_sym_index[b];
            const unsigned long long trig_sym = first_sym >= (unsigned long long)kTrig ? first_sym - kTrig : 0ull;
            rec->demod_index = trig_sym * (unsigned long long)kOS;
            rec->sample_index = rec->demod_index * (unsigned long long)decim;
            rec->corr = 0.0f;
            rec->run_length = 0;
            rec->pad[0] = 0; rec->pad[1] = 0;
        }
    } else {
        const Accepted a = acc[b];
        // the 3374 half-symbols after the trigger, sliced at the chosen sampling phase (recc_impl.cc:124-126)
        for (int s = t; s < kCapture; s += nt) {
            const float v = __ldcg(&dring[(a.pos + (unsigned long long)(kOS * (kTrig + s))) & dmask]);
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
    // (when one call accepts more bursts than the ring holds, only the newest ring_len are written: two CTAs must
    // never race for the same slot)
    if (n_acc - b <= ring_len) {
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(rec);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(&host_ring[(rec_base + b) % ring_len]);
        for (int w = t; w < (int)(sizeof(amps_burst) / 8); w += nt) dst[w] = src[w];
    }
    __threadfence_system();
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// stand-alone trigger search + selection on the demod rings, one or many channels per launch: CTAs [cta_first, cta_first +
// cta_count) of the launch search a contiguous slice each of channel c's groups [g_lo, g_hi); the last of them to finish
// selects.  Same rules (and the same device routines) as inside rx_front_kernel.
//   kCapture2 (calls so short that at most two bursts can become capturable): that last CTA also captures them, and the
// call needs no capture launch at all -- two launches per call, the side stream's share of a small pipelined call halves.
// ---------------------------------------------------------------------------------------------
template <bool kCapture2>
__global__ void __launch_bounds__(256) rx_search_kernel(const __grid_constant__ RxSearchParams p) {
    __shared__ uint32_t s_last;
    __shared__ SearchScratch sc;
    __shared__ Candidate s_cand[64];                   // (small: this CTA must fit on an SM next to two front-kernel CTAs)
#define S_PROF(k) do { if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) p.prof[9 + (k)] = global_ns(); } while (0)
    S_PROF(0);
    uint32_t c = 0;
    while (c + 1 < p.nchan && blockIdx.x >= p.ch[c + 1].cta_first) ++c;
    const RxSearchChan &ch = p.ch[c];
    if (threadIdx.x == 0) sc.nlc = 0u;
    __syncthreads();
    const unsigned long long n = ch.g_hi > ch.g_lo ? ch.g_hi - ch.g_lo : 0ull;
    const unsigned long long per = (n + ch.cta_count - 1) / ch.cta_count;
    unsigned long long lo = ch.g_lo + per * (blockIdx.x - ch.cta_first), hi = lo + per;
    if (lo > ch.g_hi) lo = ch.g_hi;
    if (hi > ch.g_hi) hi = ch.g_hi;
    search_groups(ch.dring, ch.hring, ch.dmask, ch.state, ch.cand, lo, hi, &sc);
    __syncthreads();
    S_PROF(1);
    if (threadIdx.x == 0) {
        flush_candidates(ch.state, ch.cand, &sc);
        __threadfence();
        s_last = atomicAdd(&ch.state->search_done, 1u) + 1u == ch.cta_count ? 1u : 0u;
    }
    __syncthreads();
    S_PROF(2);
    if (s_last) {
        __threadfence();
        select_channel(ch.state, ch.cand, ch.cand + kMaxCand, ch.acc + (size_t)ch.par * kMaxAccept, ch.par, ch.total_d, ch.host_pub, s_cand, 64u);
        S_PROF(3);
        if (threadIdx.x == 0) ch.state->search_done = 0u;
        if constexpr (kCapture2) {
            __shared__ __align__(16) unsigned char rec_raw[sizeof(amps_burst)];
            __shared__ uint8_t s_valid[40];
            __shared__ unsigned int s_errs[8];
            RxState *state = ch.state;
            const unsigned int n_acc = state->n_acc[ch.par];
            const unsigned long long rec_base = state->rec_base[ch.par];
            for (unsigned int b = 0; b < n_acc; ++b)
                capture_burst(ch.dring, ch.dmask, ch.acc + (size_t)ch.par * kMaxAccept, nullptr, nullptr, ch.decim, ch.host_ring, ch.ring_len,
                              rec_base, b, n_acc, reinterpret_cast<amps_burst *>(rec_raw), s_valid, s_errs);
            if (threadIdx.x == 0 && n_acc > 0) {
                __threadfence_system();
                ch.host_pub->nrec_total = rec_base + n_acc;
            }
        }
    }
    S_PROF(4);
#undef S_PROF
}

cudaError_t launch_rx_search(const RxSearchParams &p, int grid, bool capture_too, cudaStream_t st) {
    if (grid <= 0) return cudaSuccess;
    if (capture_too) rx_search_kernel<true><<<grid, 256, 0, st>>>(p);
    else rx_search_kernel<false><<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

// ============================================================================================
// burst capture (CTAs [cta_first, cta_first + cta_count) of the launch serve channel c, one burst at a time)
// ============================================================================================
// blobs == nullptr: feed-forward timing, the symbols are sliced out of the demod ring at the accepted phase.
// blobs != nullptr: M&M timing mode, amps.recc already cut the 3374-byte blobs (rx_mm_recc_kernel).
__global__ void __launch_bounds__(256) rx_capture_kernel(const __grid_constant__ RxCaptureParams p) {
    __shared__ __align__(16) unsigned char rec_raw[sizeof(amps_burst)];
    __shared__ uint8_t s_valid[40];
    __shared__ unsigned int s_errs[8];
    uint32_t c = 0;
    while (c + 1 < p.nchan && blockIdx.x >= p.ch[c + 1].cta_first) ++c;
    const RxCaptureChan &ch = p.ch[c];
    RxState *state = ch.state;
    const unsigned int n_acc = state->n_acc[ch.par];
    const unsigned long long rec_base = state->rec_base[ch.par];
    amps_burst *rec = reinterpret_cast<amps_burst *>(rec_raw);
    for (unsigned int b = blockIdx.x - ch.cta_first; b < n_acc; b += ch.cta_count) {
        capture_burst(ch.dring, ch.dmask, ch.acc, ch.blobs, ch.blob_sym_index, ch.decim, ch.host_ring, ch.ring_len, rec_base, b, n_acc, rec,
                      s_valid, s_errs);
        if (threadIdx.x == 0) {
            const unsigned int prev = atomicAdd(&state->done, 1u);
            if (prev + 1 == n_acc) {                          // last burst of the call: every record is on its way, publish the count
                state->done = 0;
                if (ch.blobs) state->pub_overflow = state->cand_overflow;      // M&M mode: no selection, amps.recc counts the surplus
                __threadfence_system();
                if (ch.blobs) ch.host_pub->cand_overflow = state->cand_overflow;
                ch.host_pub->nrec_total = rec_base + n_acc;
            }
        }
        __syncthreads();
    }
}

cudaError_t launch_rx_capture(const RxCaptureParams &p, int grid, cudaStream_t st) {
    if (grid <= 0) return cudaSuccess;
    rx_capture_kernel<<<grid, 256, 0, st>>>(p);
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

__global__ void __launch_bounds__(128) rx_mm_kernel(const __grid_constant__ RxMmParams p) {
    const RxMmChan &c = p.ch[blockIdx.x];                                 // one CTA per channel
    const float *__restrict__ dring = c.dring;
    const uint32_t dmask = c.dmask;
    const unsigned long long total_d = c.total_d;
    MmState *st = c.mm;
    const float *__restrict__ table = p.table;
    uint8_t *__restrict__ sym_out = c.sym;
    const unsigned int sym_cap = c.sym_cap;
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
// bookkeeping select_channel does in the feed-forward mode.
__global__ void __launch_bounds__(256) rx_mm_recc_kernel(const __grid_constant__ RxMmParams p) {
    const RxMmChan &c = p.ch[blockIdx.x];
    ReccCompatState *cs = c.cs;
    const MmState *mm = c.mm;
    const uint8_t *__restrict__ sym = c.sym;
    uint8_t *blobs = c.blobs;
    unsigned long long *blob_sym_index = c.blob_sym_index;
    const int max_blobs = p.max_blobs;
    RxState *state = c.state;
    RxPublished *host_pub = c.host_pub;
    const uint32_t par = c.par;
    const unsigned int n = mm->n_new;
    const int nchunks = (int)((n + (unsigned)kMmQuantum - 1u) / (unsigned)kMmQuantum);
    int nb = recc_compat_run(cs, sym, nchunks,
                             [n](int c) { const unsigned int rem = n - (unsigned)c * (unsigned)kMmQuantum; return rem < (unsigned)kMmQuantum ? rem : (unsigned)kMmQuantum; },
                             blobs, max_blobs, blob_sym_index);
    if (threadIdx.x == 0) {
        if (nb > max_blobs) { state->cand_overflow += (unsigned int)(nb - max_blobs); nb = max_blobs; }
        state->n_acc[par] = (unsigned int)nb;
        state->rec_base[par] = state->nrec_total;
        state->nrec_total += (unsigned long long)nb;
        if (nb == 0 && state->cand_overflow != state->pub_overflow) {
            state->pub_overflow = state->cand_overflow;
            __threadfence_system();
            host_pub->cand_overflow = state->cand_overflow;
        }
    }
}

cudaError_t launch_rx_mm(const RxMmParams &p, cudaStream_t st) {
    if (p.nchan == 0) return cudaSuccess;
    rx_mm_kernel<<<p.nchan, 128, 0, st>>>(p);
    rx_mm_recc_kernel<<<p.nchan, 256, 0, st>>>(p);
    return cudaGetLastError();
}

// per-device opt-in to large dynamic shared memory (call once per device after cudaSetDevice)
template <typename K>
static cudaError_t front_attrs(K kernel, size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
template <int kMaxChan, bool kFused>
static cudaError_t front_attrs_all() {
    cudaError_t e;
    if ((e = front_attrs(rx_front_kernel<float2, kFc32Regs, false, kMaxChan, kFused>, front_smem_bytes<float2, kFused>())) != cudaSuccess) return e;
    if ((e = front_attrs(rx_front_kernel<short2, kSc16Ctas, false, kMaxChan, kFused>, front_smem_bytes<short2, kFused>())) != cudaSuccess) return e;
    return front_attrs(rx_front_kernel<short2, kSc16Ctas, true, kMaxChan, kFused>, front_smem_bytes<short2, kFused>());
}
cudaError_t rx_configure_device() {
    cudaError_t e;
    if ((e = front_attrs_all<1, false>()) != cudaSuccess) return e;
    if ((e = front_attrs_all<1, true>()) != cudaSuccess) return e;
    if ((e = front_attrs_all<kMaxBatch, false>()) != cudaSuccess) return e;
    if ((e = front_attrs_all<kMaxBatch, true>()) != cudaSuccess) return e;
    // The side-stream kernels share SMs with the NEXT call's front kernel, whose two CTAs need 199 KB of shared memory per
    // SM.  An SM's L1/shared split is fixed while CTAs are resident: if a kernel that wants a big L1 gets there first, the
    // front CTAs wait until it has left.  Ask for the front kernel's split everywhere.
    const void *side[] = {(const void *)rx_search_kernel<false>, (const void *)rx_search_kernel<true>, (const void *)rx_capture_kernel, (const void *)rx_mm_kernel,
                          (const void *)rx_mm_recc_kernel};
    for (const void *f : side) {
        e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace amps
