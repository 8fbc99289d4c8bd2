// rx_kernels.cuh -- launch interface of the fused RECC receive kernels (implementation: rx_kernels.cu)
#pragma once
#include "spec.cuh"
#include "blocks_kernels.cuh"
#include "../../include/amps_b200.h"

namespace amps {

// Geometry of the front-end kernel.
//   tile  = kTB blocks of 25 samples, one block per thread (stage 1: NCO + CIC^3 /25);
//   pass  = kPassTiles tiles -> 2*kR*kTB samples @400 kS/s -> kR*kTB demodulated outputs, kR per thread
//           (stage 2: channel filter /2 with kR-fold register reuse of every shared-memory load).
constexpr int kTB        = 192;              // threads per CTA == 25-sample blocks per tile
constexpr int kStages    = 2;                // TMA ring depth
constexpr int kR         = 4;                // outputs per thread per pass (power of two)
constexpr int kLogR      = 2;
constexpr int kTile      = kTB * kD1;        // 4800 input samples, 38400 bytes
constexpr int kPassTiles = 2 * kR;
constexpr int kPass      = kPassTiles * kTile;   // 38400 input samples per pass (API granularity)
constexpr int kPassOut   = kR * kTB;         // 768 demodulated samples per pass
constexpr int kWarmTiles = 2;                // history a CTA re-reads before its first pass (>= 306 blocks)
constexpr int kHist      = kWarmTiles * kTile;   // input history carried between calls
constexpr int kPorchCols = 40;               // columns of v history kept in front of each row (>= (150 + 2R)/R)
constexpr int kRowLen    = kPorchCols + kTB + 2; // pairs per row (+2: rows land 32 B apart in bank space)
constexpr int kPass400   = kPassTiles * kTB;     // native 400 kS/s front end: 1536 input samples per pass (no CIC stage)

struct RxFrontParams {
    const void   *chunk;     // logical samples [0, npass*kPass): float2 (fc32) or short2 (sc16, the USRP's wire format)
    const void   *tail;      // logical samples [-kHist, 0), same format
    float        *dring;     // demod ring, indexed by (absolute demod index & dmask)
    void         *tail_out;  // 10 MS/s kernel: where to leave the next call's history (kHist samples), or nullptr
    uint32_t     *hring;     // hard decisions d >= 0, bit (i & 31) of word ((i & dmask) >> 5)
    float2       *ydump;     // optional: complex baseband of this call (npass*kPassOut entries) or nullptr
    uint64_t      q_base;    // absolute demod index of this call's first output
    uint32_t      dmask;
    uint32_t      npass;
    uint32_t      pass_per_cta;
    unsigned long long n_base;  // 400 kS/s front end: absolute index of logical sample 0
    uint32_t      blk_base;  // absolute 25-sample block index (mod 2^32) of logical sample 0
    uint32_t      fcw25;     // NCO phase step per block (25 * fcw mod 2^32)
    float         in_scale;  // sc16 input: x = (float)int16 * in_scale (one fp32 multiply per component)
    float2        w[kD1];    // NCO phasors inside a block
    float2        wj[kD1];   // j * w = (-w.im, w.re): the second operand of the complex product, ready-made so that it is a
                             // uniform-register operand of FFMA2 instead of a negate + move per sample
    float         g[75];     // CIC^3 taps (73 + 2 zeros)
    float         h2[300];   // channel filter (299 + pad)
};

struct RxState {             // device-resident stream state
    unsigned long long lo;          // next demod position to search
    unsigned long long resume_at;   // positions below this are inside an already captured burst
    unsigned long long nrec_total;  // bursts published since stream start (monotonic)
    unsigned long long rec_base;    // nrec_total before the bursts accepted by the last select
    unsigned int       ncand;       // candidates found by the detect kernel (reset by select)
    unsigned int       cand_overflow;
    unsigned int       n_acc;       // bursts accepted by the last select, captured by the capture kernel
    unsigned int       done;        // capture CTAs finished
};

struct Accepted { unsigned long long pos; float corr; unsigned int run; };

struct RxPublished {         // mirror of the counters in mapped pinned host memory, written by the select kernel
    unsigned long long nrec_total;
    unsigned int       cand_overflow;
    unsigned int       pad;
};

// one run of adjacent sampling phases that all match the trigger 74/74, found by the detect kernel
struct Candidate {
    unsigned long long start;   // first matching position of the run
    unsigned long long best;    // position of the soft-correlation peak inside the run (first maximum)
    float              corr;
    unsigned int       run;     // run length; bit 31 set = the run reaches the end of the searched range
};

// M&M timing mode: state of the clock_recovery_mm_ff recurrence (device-resident, carried between calls)
struct MmState {
    float mu, omega, last;
    unsigned int       n_new;       // half-symbols produced by the last rx_mm_kernel launch
    unsigned long long pos;         // absolute demod index of the interpolator window's first sample
    unsigned long long nsym_total;
};
constexpr int kMmPhases = 129;     // interpolator table rows (mu = 0, 1/128 .. 1)
constexpr int kMmQuantum = 256;    // bytes per emulated amps.recc work() call

constexpr int kMaxCand = 8192;
constexpr int kMaxAccept = 512;    // bursts one call can publish

size_t rx_front_smem_bytes();
cudaError_t rx_configure_device();
cudaError_t launch_rx_front(const RxFrontParams &p, int grid, cudaStream_t st, bool sc16 = false, bool unit = false);
int rx_front_ctas_per_sm(bool sc16);
cudaError_t launch_rx_front400(const RxFrontParams &p, int grid, cudaStream_t st, bool sc16 = false, bool unit = false);
cudaError_t launch_rx_detect(const float *dring, const uint32_t *hring, uint32_t dmask, RxState *state, Candidate *cand,
                             unsigned long long scan_lo, unsigned long long scan_hi, int max_ctas, cudaStream_t st);
// select: sorts the candidates (cand must hold 2 x kMaxCand entries: list + sorted scratch), groups runs, picks sampling
// phases -> acc[0 .. state->n_acc)
cudaError_t launch_rx_select(RxState *state, Candidate *cand, Accepted *acc, unsigned long long scan_hi,
                             RxPublished *host_pub, cudaStream_t st);
// capture: one CTA per accepted burst (grid = upper bound, surplus CTAs exit): gathers the 3374 half-symbols,
// decodes, and streams the record into host_ring[(rec_base + b) % ring_len] (mapped pinned host memory)
cudaError_t launch_rx_capture(const float *dring, uint32_t dmask, RxState *state, const Accepted *acc, int grid,
                              amps_burst *host_ring, unsigned int ring_len, RxPublished *host_pub, unsigned int decim,
                              cudaStream_t st, const uint8_t *blobs = nullptr, const unsigned long long *blob_sym_index = nullptr);
// M&M timing mode: serial clock recovery + slicer over the demod ring up to total_d, then amps.recc on the new
// half-symbols; leaves state->n_acc blobs (<= max_blobs) for launch_rx_capture(..., blobs, blob_sym_index)
cudaError_t launch_rx_mm(const float *dring, uint32_t dmask, unsigned long long total_d, MmState *mm, const float *table,
                         uint8_t *sym, unsigned int sym_cap, ReccCompatState *cs, uint8_t *blobs,
                         unsigned long long *blob_sym_index, int max_blobs, RxState *state, RxPublished *host_pub,
                         cudaStream_t st);
cudaError_t launch_decode_blobs(const uint8_t *blobs, int nbursts, amps_recc_words *out, cudaStream_t st);

}  // namespace amps
