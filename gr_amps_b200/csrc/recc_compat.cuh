// recc_compat.cuh -- the amps.recc state machine (lib/recc_impl.cc:93-145, buffer quirks included) as a block-wide
// device routine, shared by the byte-input block (blocks_kernels.cu) and the M&M timing mode of the IQ path
// (rx_kernels.cu).  Integer / byte work: bit-exact with the reference by construction.
#pragma once
#include "blocks_kernels.cuh"

namespace amps {

static __device__ __constant__ uint8_t c_recc_trig[kTrig] = {
    0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,0,1,1,0,
    0,1,0,1,0,1,1,0,1,0,1,0,0,1,1,0,1,0,0,1,1,0};

// block-wide memmove with memmove semantics for dst < src (forward, read-all-then-write-all per batch)
__device__ inline void block_move_down(uint8_t *dst, const uint8_t *src, unsigned int n) {
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

// Runs work() calls 0 .. nchunks-1 in order; call c consumes size_of(c) bytes of `in`.  Every captured blob goes to
// blobs_out[slot * kCapture] (slots beyond max_blobs are counted but not stored); blob_sym_index, if given, receives
// the index, in the stream of all bytes ever appended, of the blob's first byte (meaningful while no wrap quirk hit).
// Returns the number of blobs.  All threads of the CTA must call it.
template <typename SizeFn>
__device__ int recc_compat_run(ReccCompatState *st, const uint8_t *__restrict__ in, int nchunks, SizeFn size_of,
                               uint8_t *blobs_out, int max_blobs, unsigned long long *blob_sym_index) {
    __shared__ int s_first;
    __shared__ unsigned int s_len;
    __shared__ int s_pending;
    __shared__ int s_nblobs;
    const unsigned int t = threadIdx.x, nt = blockDim.x;
    if (t == 0) { s_len = st->len; s_pending = st->pending; s_nblobs = 0; }
    __syncthreads();
    size_t in_off = 0;
    unsigned long long appended = st->appended;
    for (int c = 0; c < nchunks; ++c) {
        const unsigned int n = size_of(c);
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
        appended += n;
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
                    if (slot < max_blobs) {
                        for (unsigned int i = t; i < (unsigned)kCapture; i += nt)
                            blobs_out[(size_t)slot * kCapture + i] = st->buf[pending + kTrig + i];
                        if (blob_sym_index && t == 0) blob_sym_index[slot] = appended - (unsigned long long)captured;
                    }
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
    if (t == 0) { st->len = s_len; st->pending = s_pending; st->appended = appended; }
    __syncthreads();
    return s_nblobs;
}

}  // namespace amps
