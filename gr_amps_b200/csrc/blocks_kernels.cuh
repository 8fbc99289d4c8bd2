// blocks_kernels.cuh -- launch interface of the byte-level block kernels (blocks_kernels.cu)
#pragma once
#include "spec.cuh"

namespace amps {

constexpr int kReccBuf = 65536;        // d_symbufsz (lib/recc_impl.cc:68)
constexpr int kReccWindow = 4096;      // d_windowsz (lib/recc_impl.cc:69)
constexpr int kFoccFrameBits = 463;    // lib/focc_impl.cc:246

struct ReccCompatState {
    uint8_t  buf[kReccBuf];
    uint32_t len;
    int32_t  pending;                  // offset of a found trigger, -1 if none
    unsigned long long appended;       // bytes appended since stream start (bookkeeping only, not in the reference)
};

cudaError_t launch_recc_compat(ReccCompatState *st, const uint8_t *in, const int *chunk_sizes, int nchunks,
                               uint8_t *blobs_out, int max_blobs, int *nblobs_out, cudaStream_t stream);
cudaError_t launch_focc_bytes(const uint8_t *slots, const int *sched, unsigned long long first, unsigned long long n,
                              unsigned int sps, int busy_idle, uint8_t *out, cudaStream_t stream);
cudaError_t launch_focc_bits(const uint8_t *slots, const int *sched, unsigned long long first_bit, unsigned long long n,
                             int busy_idle, uint8_t *out, cudaStream_t stream);
cudaError_t launch_fvc_bytes(const uint8_t *bits, unsigned long long first, unsigned long long n, unsigned int sps,
                             uint8_t *out, cudaStream_t stream);

}  // namespace amps
