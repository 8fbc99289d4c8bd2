// rx_kernels.cuh -- launch interface of the fused RECC receive kernels (implementation: rx_kernels.cu)
#pragma once
#include "spec.cuh"
#include "blocks_kernels.cuh"
#include "../../include/amps_b200.h"

namespace amps {

// Geometry of the 10 MS/s front-end kernel.
//   block = 25 input samples -> one 400 kS/s sample (stage 1: NCO + CIC^3 /25), one block per thread;
//   unit  = 64 blocks = 1600 input samples = 32 demodulated samples = ONE word of hard decisions: the API granularity
//           and the granularity at which work is dealt to CTAs;
//   tile  = kTB blocks = 3 units = one TMA landing buffer (the last tile of a segment may hold 1 or 2 units);
//   pass  = kPassTiles tiles -> 2*kR*kTB samples @400 kS/s -> kR*kTB demodulated outputs, kR per thread
//           (stage 2: channel filter /2 with kR-fold register reuse of every shared-memory load).  Passes are counted
//           from the start of a CTA's SEGMENT (a contiguous run of tiles of one channel), not from the start of the
//           stream: nothing in the arithmetic depends on where a pass begins, so the result is the same for any split.
constexpr int kTB        = 192;              // threads per CTA == 25-sample blocks per tile
constexpr int kStages    = 2;                // TMA ring depth
constexpr int kR         = 4;                // outputs per thread per pass (power of two)
constexpr int kLogR      = 2;
constexpr int kUnitBlk   = 64;               // blocks per unit
constexpr int kUnit      = kUnitBlk * kD1;   // 1600 input samples (API granularity at 10 MS/s)
constexpr int kUnitOut   = kUnitBlk / kD2;   // 32 demodulated samples
constexpr int kTileUnits = kTB / kUnitBlk;   // 3
constexpr int kTile      = kTB * kD1;        // 4800 input samples, 38400 bytes
constexpr int kPassTiles = 2 * kR;
constexpr int kPass      = kPassTiles * kTile;   // 38400 input samples per full pass
constexpr int kPassOut   = kR * kTB;         // 768 demodulated samples per full pass
constexpr int kWarmTiles = 2;                // history a CTA re-reads before its segment (>= 306 blocks)
constexpr int kHist      = kWarmTiles * kTile;   // input history carried between calls
constexpr int kPorchCols = 40;               // columns of v history kept in front of each row (>= (150 + 2R)/R)
constexpr int kRowLen    = kPorchCols + kTB + 2; // pairs per row (+2: rows land 32 B apart in bank space)
constexpr int kPass400   = kPassTiles * kTB;     // native 400 kS/s front end: 1536 input samples per pass (no CIC stage)
constexpr int kRxDepth   = 4;                // calls whose side-stream work (search, capture) may be in flight: the demod ring
                                             // holds kRxDepth + 1 calls, the accepted-burst lists are kept per call slot
constexpr int kMaxGrid   = 1024;             // most CTAs a front launch may have (per-channel completion flags)
constexpr int kMaxBatch  = 64;               // channels one batched launch can carry (kernel parameter space: 32 KB)

// Trigger search geometry.  Group g = the 32 sampling positions 32g .. 32g+31.  Deciding a group (its own matches, the
// neighbours' masks that delimit runs, and the soft correlations of a run of up to 10 positions) reads the demodulated
// stream up to 32g + 63 + 10*73 = 32g + 793: group g is searched as soon as the stream is longer than that, by whoever
// produced sample 32g + 793.
constexpr int kGroupLag  = 24;               // a segment [qs, qe) of demod samples searches groups [qs/32 - 24, qe/32 - 24)
static_assert(kUnitOut * (kGroupLag + 1) > 63 + kOS * (kTrig - 1), "group lag too short for the trigger lookahead");

struct RxState {             // device-resident stream state of one channel
    unsigned long long resume_at;   // positions below this are inside an already captured burst
    unsigned long long nrec_total;  // bursts published since stream start (monotonic)
    unsigned long long rec_base[kRxDepth]; // nrec_total before the bursts accepted by the select of call slot (call number mod kRxDepth)
    unsigned int       n_acc[kRxDepth];    // bursts accepted by that select, captured by the capture kernel
    unsigned int       ncand;       // candidates on the list (new ones appended by the search, undecided ones kept by select)
    unsigned int       cand_overflow;   // candidates dropped because the list was full (monotonic)
    unsigned int       pub_overflow;    // value of cand_overflow last mirrored to the host
    unsigned int       done;        // capture: bursts of the current call finished (reset by the last one)
    unsigned int       front_done;  // front kernel: segments of this channel finished in the current launch (reset by the last one)
    unsigned int       search_done; // stand-alone search kernel: CTAs of this channel finished (it may overlap the next front kernel)
};

struct Accepted { unsigned long long pos; float corr; unsigned int run; };

struct RxPublished {         // mirror of the counters in mapped pinned host memory, written by the kernels
    unsigned long long nrec_total;
    unsigned int       cand_overflow;
    unsigned int       pad;
};

// one run of adjacent sampling phases that all match the trigger 74/74
struct Candidate {
    unsigned long long start;   // first matching position of the run
    unsigned long long best;    // position of the soft-correlation peak inside the run (first maximum)
    float              corr;
    unsigned int       run;     // run length (1..10)
};

constexpr int kMaxCand = 8192;
constexpr int kMaxAccept = 512;    // bursts one call can publish

// Per-channel arguments of a front launch.  They live in kernel parameter space (constant bank): the NCO tables are
// FFMA2 operands straight from there.
struct RxChan {
    const void   *chunk;     // this call's new samples: logical samples [carry, carry + nchunk)
    const void   *tail;      // logical samples [-kHist, carry): history + the samples the previous call could not use
    void         *tail_out;  // where to leave the next call's tail: logical [units*kUnit - kHist, carry + nchunk)
    float        *dring;     // demod ring, indexed by (absolute demod index & dmask)
    uint32_t     *hring;     // hard decisions d >= 0, bit (i & 31) of word ((i & dmask) >> 5)
    float2       *ydump;     // optional: complex baseband of this call (units*kUnitOut entries) or nullptr
    RxState      *state;
    Candidate    *cand;      // 2 x kMaxCand: list + the select's sorted scratch
    Accepted     *acc;       // kRxDepth x kMaxAccept, by call slot
    RxPublished  *host_pub;
    uint32_t     *flags;     // kMaxGrid words, all zero between launches: flags[j] counts the CTAs whose output the boundary
                             // groups of CTA j's segment read and that have published it (the last one searches them)
    uint64_t      q_base;    // absolute demod index of this call's first output
    uint32_t      dmask;
    uint32_t      units;     // whole units this call processes (> 0)
    uint32_t      carry;     // samples at the end of `tail` beyond the history
    uint32_t      nchunk;    // samples in `chunk`
    uint32_t      blk_base;  // absolute 25-sample block index (mod 2^32) of logical sample 0
    uint32_t      fcw25;     // NCO phase step per block (25 * fcw mod 2^32)
    uint32_t      par;       // call slot (call number mod kRxDepth): which acc / n_acc / rec_base entry this call's select fills
    uint32_t      search;    // 0: no trigger search in this launch (M&M timing mode runs its own tail)
    float         in_scale;  // sc16 input: x = (float)int16 * in_scale (one fp32 multiply per component)
    float2        w[kD1];    // NCO phasors inside a block
};

// how a launch's tiles are dealt to CTAs (see deal_lo() in rx_kernels.cu)
struct RxDeal {
    uint32_t Tt;             // tiles of the launch
    uint32_t nstat;          // CTAs of the launch
    uint32_t P;              // 0: CTA s owns tiles [Tt*s/nstat, Tt*(s+1)/nstat) (a range may cross channels);
                             // > 0: the channels all have Tc tiles and P CTAs each: CTA s owns part s % P of channel s / P
    uint32_t Tc;
};
// fills a deal for `tiles` tiles on at most `resident` CTAs; returns the grid size.  nchan channels of equal_tiles tiles each
// (equal_tiles = 0: lengths differ) are given whole CTAs when that keeps most of the machine busy: a CTA that finishes one
// channel and starts the next pays a second warm-up and a cold restart of its copy ring.
uint32_t rx_make_deal(RxDeal &d, uint32_t tiles, uint32_t resident, uint32_t nchan = 1, uint32_t equal_tiles = 0);

// first tile of CTA s (s == nstat gives Tt) and the CTA that owns a tile
__host__ __device__ __forceinline__ uint32_t deal_lo(const RxDeal &d, uint32_t s) {
    if (d.P) return (s / d.P) * d.Tc + (uint32_t)((unsigned long long)d.Tc * (s % d.P) / d.P);
    return (uint32_t)((unsigned long long)d.Tt * s / d.nstat);
}
__host__ __device__ __forceinline__ uint32_t deal_owner(const RxDeal &d, uint32_t tile) {
    if (d.P) return (tile / d.Tc) * d.P + (uint32_t)((((unsigned long long)(tile % d.Tc) + 1ull) * d.P - 1ull) / d.Tc);
    return (uint32_t)((((unsigned long long)tile + 1ull) * d.nstat - 1ull) / d.Tt);
}

template <int kMaxChan>
struct RxFrontParamsT {
    float     g[75];         // CIC^3 taps (73 + 2 zeros)
    float     h2[300];       // channel filter (299 + pad)
    float2    wj0[kD1];      // single-channel launch: j * w = (-w.im, w.re) of channel 0, the second operand of the complex
                             // product ready-made, so that it is a uniform-register operand of FFMA2 instead of a negate +
                             // move per sample (a batched launch indexes ch[] at run time and negates in the kernel)
    uint32_t  nchan;
    RxDeal    deal;
    unsigned long long *prof; // measurement aid (AMPS_RX_PROF=1): 16 time stamps per CTA, else nullptr
    uint32_t  tile_cum[kMaxChan + 1];   // tiles of channels 0..c-1
    RxChan    ch[kMaxChan];
};
using RxFrontParams1 = RxFrontParamsT<1>;
using RxFrontParamsB = RxFrontParamsT<kMaxBatch>;
static_assert(sizeof(RxFrontParamsB) <= 32000, "batched launch parameters exceed the kernel parameter space");

// native 400 kS/s front end (the reference's own rate): whole 1536-sample passes, one channel, no carry
struct RxFront400Params {
    const void   *chunk;     // logical samples [0, npass*kPass400)
    const void   *tail;      // logical samples [-kPass400, 0)
    float        *dring;
    uint32_t     *hring;
    float2       *ydump;
    uint64_t      q_base;
    uint32_t      dmask;
    uint32_t      npass;
    uint32_t      pass_per_cta;
    unsigned long long n_base;  // absolute index of logical sample 0
    uint32_t      fcw25;
    float         in_scale;
    float2        w[kD1];
    float2        wj[kD1];
    float         h2[300];
};

// capture launch: one entry per channel
struct RxCaptureChan {
    const float  *dring;
    RxState      *state;
    const Accepted *acc;     // this call's parity slot
    amps_burst   *host_ring;
    RxPublished  *host_pub;
    const uint8_t *blobs;    // M&M timing mode: the blobs amps.recc cut (else nullptr)
    const unsigned long long *blob_sym_index;
    uint32_t      dmask;
    uint32_t      ring_len;
    uint32_t      decim;
    uint32_t      par;
    uint32_t      cta_first; // first CTA of the launch that serves this channel
    uint32_t      cta_count;
};
struct RxCaptureParams {
    uint32_t      nchan;
    RxCaptureChan ch[kMaxBatch];
};

// stand-alone search + select launch: one entry per channel
struct RxSearchChan {
    const float    *dring;
    const uint32_t *hring;
    RxState        *state;
    Candidate      *cand;
    Accepted       *acc;        // kRxDepth x kMaxAccept
    RxPublished    *host_pub;
    unsigned long long g_lo, g_hi;   // groups to search
    unsigned long long total_d;      // demod samples produced so far (what select decides against)
    uint32_t        dmask;
    uint32_t        par;
    uint32_t        cta_first, cta_count;
    // for the variant that also captures (calls that can make at most two bursts capturable)
    amps_burst     *host_ring;
    uint32_t        ring_len, decim;
};
struct RxSearchParams {
    uint32_t     nchan;
    unsigned long long *prof = nullptr;    // measurement aid (AMPS_RX_PROF=1): CTA 0 stamps slots 9..15 of the front kernel's first record
    RxSearchChan ch[kMaxBatch];
};
// CTAs a channel with `groups` groups to search gets
inline uint32_t rx_search_ctas(unsigned long long groups) {
    unsigned long long n = (groups + 255ull) / 256ull;
    return (uint32_t)(n < 1ull ? 1ull : (n > 592ull ? 592ull : n));
}

// M&M timing mode: state of the clock_recovery_mm_ff recurrence (device-resident, carried between calls)
struct MmState {
    float mu, omega, last;
    unsigned int       n_new;       // half-symbols produced by the last rx_mm_kernel launch
    unsigned long long pos;         // absolute demod index of the interpolator window's first sample
    unsigned long long nsym_total;
};
constexpr int kMmPhases = 129;     // interpolator table rows (mu = 0, 1/128 .. 1)
constexpr int kMmQuantum = 256;    // bytes per emulated amps.recc work() call

size_t rx_front_smem_bytes();
cudaError_t rx_configure_device();
int rx_front_ctas_per_sm(bool sc16);
// tiles a channel with `units` whole units contributes to a launch
inline uint32_t rx_tiles_of(uint32_t units) { return (units + kTileUnits - 1) / kTileUnits; }
// front end + trigger search + (by the last CTA of each channel) candidate selection, all channels of the launch
cudaError_t launch_rx_front(const RxFrontParams1 &p, int grid, cudaStream_t st, bool sc16, bool unit, bool fused);
cudaError_t launch_rx_front_batch(const RxFrontParamsB &p, int grid, cudaStream_t st, bool sc16, bool unit, bool fused);
cudaError_t launch_rx_front400(const RxFront400Params &p, int grid, cudaStream_t st, bool sc16 = false, bool unit = false);
// stand-alone trigger search + selection on the demod rings (same rules as inside rx_front_kernel): the tail of the 400 kS/s
// front end, and of the 10 MS/s one unless AMPS_RX_FUSED_SEARCH asks for the search inside the front kernel
cudaError_t launch_rx_search(const RxSearchParams &p, int grid, bool capture_too, cudaStream_t st);
// capture: CTAs [cta_first, cta_first + cta_count) of channel c walk its accepted bursts (stride cta_count): gather the
// 3374 half-symbols, decode, and stream the record into host_ring[(rec_base + b) % ring_len] (mapped pinned host memory)
cudaError_t launch_rx_capture(const RxCaptureParams &p, int grid, cudaStream_t st);
// M&M timing mode, one CTA per channel and kernel (the recurrences of different channels run side by side): serial clock
// recovery + slicer over the demod ring up to total_d, then amps.recc on the new half-symbols; leaves state->n_acc[par] blobs
// (<= max_blobs) for the capture launch
struct RxMmChan {
    const float *dring;
    MmState     *mm;
    uint8_t     *sym;
    ReccCompatState *cs;
    uint8_t     *blobs;
    unsigned long long *blob_sym_index;
    RxState     *state;
    RxPublished *host_pub;
    unsigned long long total_d;
    uint32_t     dmask, sym_cap, par, pad;
};
struct RxMmParams {
    uint32_t     nchan;
    int          max_blobs;
    const float *table;      // 129 x 8 MMSE interpolator (one per device)
    RxMmChan     ch[kMaxBatch];
};
cudaError_t launch_rx_mm(const RxMmParams &p, cudaStream_t st);
cudaError_t launch_decode_blobs(const uint8_t *blobs, int nbursts, amps_recc_words *out, cudaStream_t st);

}  // namespace amps
