// blocks_kernels.cu -- the byte-level gr-amps blocks on the GPU:
//   recc_compat_kernel : amps.recc (lib/recc_impl.cc:93-145) including its buffer quirks
//   focc_bytes_kernel  : amps.focc byte stream (lib/focc_impl.cc:178-218, 582-647; lib/amps_packet.h:47-76)
//   fvc_bytes_kernel   : amps.fvc byte stream (lib/fvc_impl.cc:71-88, 152-193)
// These are integer / byte kernels: parity with the reference is bit-exact by construction.
#include "blocks_kernels.cuh"

namespace amps {

__device__ __constant__ uint8_t c_recc_trig[kTrig] = {
    0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,
    0,1,0,1,0,1,1,0,1,0,1,0,0,1,1,0,1,0,0,1,1,0};

// block-wide memmove with memmove semantics for dst < src (forward, read-all-then-write-all per batch)
__device__ void block_move_down(uint8_t *dst, const uint8_t *src, unsigned int n) {
    const unsigned int t = threadIdx.x, nt = blockDim.x;
    for (unsigned int base = 0; base < n; base += nt) {
        const unsigned int i = base + t;
        uint8_t v = 0;
        if (i < n) v = src[i];
        __syncthreads();
        if (i < n) dst[i] = v;
        __syncthreads();
    }
}

// One CTA executes a whole schedule of work() calls, in order, on the device-resident symbol buffer.
__global__ void __launch_bounds__(256) recc_compat_kernel(ReccCompatState *st, const uint8_t *__restrict__ in,
                                                         const int *__restrict__ chunk_sizes, int nchunks, uint8_t *blobs_out,
                                                         int max_blobs, int *nblobs_out) {
    __shared__ int s_first;
    __shared__ unsigned int s_len;
    __shared__ int s_pending;
    __shared__ int s_nblobs;
    const unsigned int t = threadIdx.x, nt = blockDim.x;
    if (t == 0) { s_len = st->len; s_pending = st->pending; s_nblobs = 0; }
    __syncthreads();
    size_t in_off = 0;
    for (int c = 0; c < nchunks; ++c) {
        const unsigned int n = (unsigned int)chunk_sizes[c];
        unsigned int len = s_len;
        int pending = s_pending;
        __syncthreads();
        if (n < 1u) continue;
        // wrap: copies the CAPACITY tail [61440, 65536), not the data tail, and forgets a pending trigger (:104-108)
        if (len + n > (unsigned)kReccBuf) {
            for (unsigned int i = t; i < (unsigned)kReccWindow; i += nt) st->buf[i] = st->buf[kReccBuf - kReccWindow + i];
            len = kReccWindow;
            pending = -1;
            __syncthreads();
        }
        for (unsigned int i = t; i < n; i += nt) st->buf[len + i] = in[in_off + i];      // append (:110-111)
        in_off += n;
        len += n;
        __syncthreads();
        if (len > (unsigned)kTrig) {
            if (pending < 0) {
                // memmem over the last min(len, n + 73) bytes, first match wins (:115-119)
                unsigned int searchsz = n + kTrig - 1;
                if (searchsz > len) searchsz = len;
                const unsigned int base = len - searchsz;
                if (t == 0) s_first = 0x7fffffff;
                __syncthreads();
                for (unsigned int p = t; p + kTrig <= searchsz; p += nt) {
                    bool ok = true;
                    for (int k = 0; k < kTrig; ++k)
                        if (st->buf[base + p + k] != c_recc_trig[k]) { ok = false; break; }
                    if (ok) atomicMin(&s_first, (int)(base + p));
                }
                __syncthreads();
                if (s_first != 0x7fffffff) pending = s_first;
                __syncthreads();
            }
            if (pending >= 0) {
                const long captured = (long)len - pending - kTrig;
                if (captured > kCapture) {                                               // strict (:124-125)
                    const int slot = s_nblobs;
                    if (slot < max_blobs)
                        for (unsigned int i = t; i < (unsigned)kCapture; i += nt)
                            blobs_out[(size_t)slot * kCapture + i] = st->buf[pending + kTrig + i];
                    __syncthreads();
                    // the LAST `pending` bytes move to the front; len shrinks by `pending` (:129-134)
                    const unsigned int k = (unsigned int)pending;
                    if (k > 0) block_move_down(st->buf, st->buf + (len - k), k);
                    len -= k;
                    pending = -1;
                    if (t == 0) s_nblobs = slot + 1;
                }
            }
        }
        __syncthreads();
        if (t == 0) { s_len = len; s_pending = pending; }
        __syncthreads();
    }
    if (t == 0) { st->len = s_len; st->pending = s_pending; *nblobs_out = s_nblobs; }
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
