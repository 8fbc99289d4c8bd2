// blocks_kernels.cu -- the byte-level gr-amps blocks on the GPU:
//   recc_compat_kernel : amps.recc (lib/recc_impl.cc:93-145) including its buffer quirks
//   focc_bytes_kernel  : amps.focc byte stream (lib/focc_impl.cc:178-218, 582-647; lib/amps_packet.h:47-76)
//   fvc_bytes_kernel   : amps.fvc byte stream (lib/fvc_impl.cc:71-88, 152-193)
// These are integer / byte kernels: parity with the reference is bit-exact by construction.
#include "blocks_kernels.cuh"
#include "recc_compat.cuh"

namespace amps {

// One CTA executes a whole schedule of work() calls, in order, on the device-resident symbol buffer.
__global__ void __launch_bounds__(256) recc_compat_kernel(ReccCompatState *st, const uint8_t *__restrict__ in,
                                                         const int *__restrict__ chunk_sizes, int nchunks, uint8_t *blobs_out,
                                                         int max_blobs, int *nblobs_out) {
    const int nblobs = recc_compat_run(st, in, nchunks, [chunk_sizes](int c) { return (unsigned int)chunk_sizes[c]; }, blobs_out,
                                       max_blobs, (unsigned long long *)nullptr);
    if (threadIdx.x == 0) *nblobs_out = nblobs;
}

cudaError_t launch_recc_compat(ReccCompatState *st, const uint8_t *in, const int *chunk_sizes, int nchunks,
                               uint8_t *blobs_out, int max_blobs, int *nblobs_out, cudaStream_t stream) {
    recc_compat_kernel<<<1, 256, 0, stream>>>(st, in, chunk_sizes, nchunks, blobs_out, max_blobs, nblobs_out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// FOCC: byte x of the requested run sits at absolute byte a = first + x of a sequence of frames;
// frame k of the sequence is slot table sched[k] (463 slots: 0, 1, or 2 = busy/idle bit).
// bit 0 -> (+1 x sps, -1 x sps), bit 1 -> (-1 x sps, +1 x sps); -1 is stored as 0xFF.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) focc_bytes_kernel(const uint8_t *__restrict__ slots, const int *__restrict__ sched,
                                                        unsigned long long first, unsigned long long n, unsigned int sps,
                                                        int busy_idle, uint8_t *__restrict__ out) {
    const unsigned long long x = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    const unsigned long long a = first + x;
    const unsigned long long frame_bytes = 926ull * sps;
    const unsigned long long fk = a / frame_bytes;
    const unsigned int within = (unsigned int)(a - fk * frame_bytes);
    const unsigned int slot = within / (2u * sps);
    const bool second_half = (within - slot * 2u * sps) >= sps;
    unsigned int bit = slots[(size_t)sched[fk] * kFoccFrameBits + slot];
    if (bit == 2u) bit = busy_idle ? 1u : 0u;
    const bool high = bit ? second_half : !second_half;
    out[x] = high ? 0x01 : 0xFF;
}

cudaError_t launch_focc_bytes(const uint8_t *slots, const int *sched, unsigned long long first, unsigned long long n,
                              unsigned int sps, int busy_idle, uint8_t *out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const unsigned int grid = (unsigned int)((n + 255) / 256);
    focc_bytes_kernel<<<grid, 256, 0, stream>>>(slots, sched, first, n, sps, busy_idle, out);
    return cudaGetLastError();
}

// FOCC as data bits (one byte per 10 kbit/s bit, busy/idle resolved): the input of the forward path's bit entry point
__global__ void __launch_bounds__(256) focc_bits_kernel(const uint8_t *__restrict__ slots, const int *__restrict__ sched,
                                                       unsigned long long first_bit, unsigned long long n, int busy_idle,
                                                       uint8_t *__restrict__ out) {
    const unsigned long long x = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    const unsigned long long a = first_bit + x;
    const unsigned long long fk = a / kFoccFrameBits;
    unsigned int bit = slots[(size_t)sched[fk] * kFoccFrameBits + (unsigned int)(a - fk * kFoccFrameBits)];
    if (bit == 2u) bit = busy_idle ? 1u : 0u;
    out[x] = (uint8_t)bit;
}

cudaError_t launch_focc_bits(const uint8_t *slots, const int *sched, unsigned long long first_bit, unsigned long long n,
                             int busy_idle, uint8_t *out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const unsigned int grid = (unsigned int)((n + 255) / 256);
    focc_bits_kernel<<<grid, 256, 0, stream>>>(slots, sched, first_bit, n, busy_idle, out);
    return cudaGetLastError();
}

// FVC: byte x of the run is byte (first + x) of the replay of `bits`.
__global__ void __launch_bounds__(256) fvc_bytes_kernel(const uint8_t *__restrict__ bits, unsigned long long first,
                                                       unsigned long long n, unsigned int sps, uint8_t *__restrict__ out) {
    const unsigned long long x = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    const unsigned long long p = first + x;
    const unsigned long long bi = p / (2ull * sps);
    const bool second_half = (p - bi * 2ull * sps) >= sps;
    const bool high = bits[bi] ? second_half : !second_half;
    out[x] = high ? 0x01 : 0xFF;
}

cudaError_t launch_fvc_bytes(const uint8_t *bits, unsigned long long first, unsigned long long n, unsigned int sps,
                             uint8_t *out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const unsigned int grid = (unsigned int)((n + 255) / 256);
    fvc_bytes_kernel<<<grid, 256, 0, stream>>>(bits, first, n, sps, out);
    return cudaGetLastError();
}

}  // namespace amps
