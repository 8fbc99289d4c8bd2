// rx_kernels.cuh -- launch interface of the fused RECC receive kernels (implementation: rx_kernels.cu)
#pragma once
#include "spec.cuh"
#include "../../include/amps_b200.h"

namespace amps {

// Geometry of the front-end kernel.  One "pass" = kS2Tiles tiles = 2*TB blocks of 25 samples
// -> TB demodulated outputs (one per thread).
constexpr int kTB      = 192;              // threads per CTA == 25-sample blocks per tile
constexpr int kStages  = 2;                // TMA ring depth
constexpr int kTile    = kTB * kD1;        // 4800 input samples, 38400 bytes
constexpr int kPass    = 2 * kTile;        // 9600 input samples per stage-2 pass
constexpr int kHist    = kPass;            // input history carried between calls (one pass)
constexpr int kPorch   = 304;              // v history mirrored in front of the ring
constexpr int kVRing   = 4 * kTB;          // two passes of 400 kS/s samples

struct RxFrontParams {
    const float2 *chunk;     // logical samples [0, npass*kPass)
    const float2 *tail;      // logical samples [-kHist, 0)
    float        *dring;     // demod ring, indexed by (absolute demod index & dmask)
    float2       *ydump;     // optional: complex baseband of this call (npass*kTB entries) or nullptr
    uint64_t      q_base;    // absolute demod index of this call's first output
    uint32_t      dmask;
    uint32_t      npass;
    uint32_t      blk_base;  // absolute 25-sample block index (mod 2^32) of logical sample 0
    uint32_t      fcw25;     // NCO phase step per block (25 * fcw mod 2^32)
    float2        w[kD1];    // NCO phasors inside a block
    float         g[75];     // CIC^3 taps (73 + 2 zeros)
    float         h2[300];   // channel filter (299 + pad)
};

struct RxState {             // device-resident stream state
    unsigned long long lo;          // next demod position to search
    unsigned long long resume_at;   // positions below this are inside an already captured burst
    unsigned long long nrec_total;  // bursts published since stream start (monotonic)
    unsigned int       ncand;       // candidates found by the detect kernel (reset by select)
    unsigned int       cand_overflow;
};

struct RxPublished {         // mirror of the counters in mapped pinned host memory, written by the select kernel
    unsigned long long nrec_total;
    unsigned int       cand_overflow;
    unsigned int       pad;
};

struct Candidate { unsigned long long pos; float corr; unsigned int pad; };

constexpr int kMaxCand = 8192;
constexpr int kMaxAccept = 512;    // bursts one call can publish

size_t rx_front_smem_bytes();
cudaError_t rx_configure_device();
cudaError_t launch_rx_front(const RxFrontParams &p, int grid, cudaStream_t st);
cudaError_t launch_rx_detect(const float *dring, uint32_t dmask, RxState *state, Candidate *cand,
                             unsigned long long scan_lo, unsigned long long scan_hi, cudaStream_t st);
// records are assembled in `scratch` (device memory, kMaxAccept entries) and then streamed into
// host_ring[(nrec_total + a) % ring_len] (mapped pinned host memory, device-visible alias)
cudaError_t launch_rx_select(const float *dring, uint32_t dmask, RxState *state, Candidate *cand,
                             unsigned long long scan_hi, amps_burst *scratch, amps_burst *host_ring,
                             unsigned int ring_len, RxPublished *host_pub, cudaStream_t st);
cudaError_t launch_decode_blobs(const uint8_t *blobs, int nbursts, amps_recc_words *out, cudaStream_t st);

}  // namespace amps
