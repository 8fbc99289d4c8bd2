// c_api_blocks.cu -- extern "C" entry points of the byte-level blocks: amps_recc_* (compat sink),
// amps_focc_* and amps_fvc_* (Manchester half-symbol sources).  Host state machines mirror the
// reference's work()/message-handler behaviour; every byte that leaves a source and every blob
// that leaves the sink is produced by a CUDA kernel (blocks_kernels.cu).
#include "blocks_kernels.cuh"
#include "common.h"
#include "proto.h"

#include <cstring>
#include <deque>
#include <new>
#include <vector>

using namespace amps;

// =============================================================================================
// recc (lib/recc_impl.cc)
// =============================================================================================
struct amps_recc {
    int device = 0;
    cudaStream_t stream = nullptr;
    ReccCompatState *d_state = nullptr;
    uint8_t *d_in = nullptr;  size_t in_cap = 0;
    int *d_sizes = nullptr;   int sizes_cap = 0;
    uint8_t *d_blobs = nullptr; int blobs_cap = 0;
    int *d_nblobs = nullptr;
    std::vector<uint8_t> h_blobs;
};

static int recc_reserve(amps_recc *h, size_t total, int nchunks) {
    if (total > h->in_cap) {
        cudaFree(h->d_in); h->d_in = nullptr;
        h->in_cap = total * 2 + 4096;
        CK(cudaMalloc(&h->d_in, h->in_cap));
    }
    if (nchunks > h->sizes_cap) {
        cudaFree(h->d_sizes); h->d_sizes = nullptr;
        h->sizes_cap = nchunks * 2 + 16;
        CK(cudaMalloc(&h->d_sizes, sizeof(int) * h->sizes_cap));
    }
    const int need = (int)(total / (kTrig + kCapture)) + 2;      // each publish consumes > 3448 fresh symbols
    if (need > h->blobs_cap) {
        cudaFree(h->d_blobs); h->d_blobs = nullptr;
        h->blobs_cap = need * 2;
        CK(cudaMalloc(&h->d_blobs, (size_t)h->blobs_cap * kCapture));
    }
    return AMPS_OK;
}

static int recc_init(amps_recc *h) {
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaMalloc(&h->d_state, sizeof(ReccCompatState)));
    CK(cudaMemset(h->d_state, 0, sizeof(ReccCompatState)));            // zero-initialised buffer (:75)
    const int32_t none = -1;
    CK(cudaMemcpy(&h->d_state->pending, &none, sizeof none, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->d_nblobs, sizeof(int)));
    return AMPS_OK;
}

extern "C" int amps_recc_create(int device, amps_recc **out) {
    if (!out) return set_error(AMPS_E_INVAL, "null argument");
    *out = nullptr;
    int st = select_device(device);
    if (st != AMPS_OK) return st;
    amps_recc *h = new (std::nothrow) amps_recc();
    if (!h) return set_error(AMPS_E_NOMEM, "out of host memory");
    h->device = device;
    st = recc_init(h);
    if (st != AMPS_OK) { amps_recc_destroy(h); return st; }           // (the error text set by the failing call survives the clean-up)
    *out = h;
    return AMPS_OK;
}

extern "C" int amps_recc_destroy(amps_recc *h) {
    if (!h) return AMPS_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaFree(h->d_state); cudaFree(h->d_in); cudaFree(h->d_sizes); cudaFree(h->d_blobs); cudaFree(h->d_nblobs);
    delete h;
    return AMPS_OK;
}

extern "C" int amps_recc_work_chunks(amps_recc *h, const uint8_t *in, const int *chunk_sizes, int nchunks,
                                     amps_blob_cb cb, void *user) {
    if (!h || nchunks < 0 || (nchunks && (!in || !chunk_sizes))) return set_error(AMPS_E_INVAL, "bad argument");
    size_t total = 0;
    for (int i = 0; i < nchunks; ++i) {
        if (chunk_sizes[i] >= kReccBuf - kReccWindow)
            return set_error(AMPS_E_OVERFLOW, "noutput_items must be < 61440 (the reference asserts, lib/recc_impl.cc:103)");
        if (chunk_sizes[i] > 0) total += (size_t)chunk_sizes[i];
    }
    if (nchunks == 0) return AMPS_OK;
    CK(cudaSetDevice(h->device));
    int rc = recc_reserve(h, total, nchunks);
    if (rc != AMPS_OK) return rc;
    // chunks with n < 1 are no-ops in the reference (:98-101); keep their slots so indices line up
    std::vector<int> sizes(chunk_sizes, chunk_sizes + nchunks);
    std::vector<uint8_t> packed;
    const uint8_t *src = in;
    bool has_neg = false;
    for (int s : sizes) has_neg |= s < 0;
    if (has_neg) return set_error(AMPS_E_INVAL, "negative chunk size");
    if (total) CK(cudaMemcpyAsync(h->d_in, src, total, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_sizes, sizes.data(), sizeof(int) * nchunks, cudaMemcpyHostToDevice, h->stream));
    CKL(launch_recc_compat(h->d_state, h->d_in, h->d_sizes, nchunks, h->d_blobs, h->blobs_cap, h->d_nblobs, h->stream));
    int nb = 0;
    CK(cudaMemcpyAsync(&nb, h->d_nblobs, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (nb > h->blobs_cap) nb = h->blobs_cap;
    if (nb > 0) {
        h->h_blobs.resize((size_t)nb * kCapture);
        CK(cudaMemcpy(h->h_blobs.data(), h->d_blobs, (size_t)nb * kCapture, cudaMemcpyDeviceToHost));
        if (cb) for (int i = 0; i < nb; ++i) cb(h->h_blobs.data() + (size_t)i * kCapture, user);   // message_port_pub("bursts") (:126)
    }
    return AMPS_OK;
}

extern "C" int amps_recc_work(amps_recc *h, const uint8_t *in, int n, amps_blob_cb cb, void *user) {
    if (n < 1) return AMPS_OK;                                          // the reference prints and returns 0 (:98-101)
    return amps_recc_work_chunks(h, in, &n, 1, cb, user);
}

// =============================================================================================
// focc (lib/focc_impl.cc)
// =============================================================================================
static const int kEphemeralPool = 4096;

struct amps_focc {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_gen = nullptr;          // recorded behind the last generate on whatever stream it ran
    bool gen_pending = false;
    unsigned sps = 1;
    int nsuper = 0;
    std::vector<bool> filler;              // per superframe slot
    uint8_t *d_slots = nullptr;            // (nsuper + kEphemeralPool) x 463 slot tables
    int pool_next = 0;
    std::deque<int> queue;                 // ids of queued ephemeral frames (frame_queue, :565-580)
    // emission state
    int frame_idx = 0;                     // position in the superframe
    int cur_id = 0;                        // slot-table id of the frame being sent (ephemeral ids >= nsuper)
    int bit = 0, off = 0;
    bool at_end = false;
    int busy_idle = 1;                     // lib/amps_common.h:7, set to 1 by the ctor (:111)
    // staging
    int *d_sched = nullptr;  size_t sched_cap = 0;
    uint8_t *d_out = nullptr; size_t out_cap = 0;
    std::vector<int> h_sched;
};

static inline int focc_burst_last_bit(int bit) {      // last bit index of the burst that contains `bit`
    if (bit <= 22) return 22;
    return 22 + 22 * ((bit - 23) / 22 + 1);
}

// step over the END marker we are sitting on; at the end of a frame this advances the superframe and
// lets a queued frame replace a filler slot (lib/focc_impl.cc:491-507)
static void focc_step_over_end(amps_focc *h) {
    if (h->bit == kFoccFrameBits) {
        h->frame_idx = (h->frame_idx + 1) % h->nsuper;
        h->cur_id = h->frame_idx;
        if (h->filler[(size_t)h->frame_idx] && !h->queue.empty()) {
            h->cur_id = h->queue.front();
            h->queue.pop_front();
        }
        h->bit = 0;
    }
    h->at_end = false;
}

static int focc_upload_frame(amps_focc *h, int id, const uint8_t *wa, const uint8_t *wb) {
    const auto slots = focc_frame_slots(wa, wb);
    CK(cudaMemcpy(h->d_slots + (size_t)id * kFoccFrameBits, slots.data(), kFoccFrameBits, cudaMemcpyHostToDevice));
    return AMPS_OK;
}

static int focc_init(amps_focc *h, int aggressive_registration) {
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_gen, cudaEventDisableTiming));
    // superframe (lib/focc_impl.cc:383-406 / :420-466)
    std::vector<Word28> words;
    const int halves = aggressive_registration ? 2 : 1;
    for (int half = 0; half < halves; ++half) {
        words.push_back(overhead_word_1(0, 16, true, false, false, aggressive_registration ? 4 : 3)); h->filler.push_back(false);
        words.push_back(overhead_word_2(0, true, true, true, true, 0, 23, true, true, 23, false)); h->filler.push_back(false);
        words.push_back(access_type_parameters_global_action(0, false)); h->filler.push_back(false);
        if (aggressive_registration) { words.push_back(registration_increment_global_action(0, 100, false)); h->filler.push_back(false); }
        words.push_back(registration_id(0, (aggressive_registration && half == 1) ? 500 : 0, true)); h->filler.push_back(false);
        const int nfill = aggressive_registration ? 14 : 15;
        for (int i = 0; i < nfill; ++i) { words.push_back(control_filler_word()); h->filler.push_back(true); }
    }
    h->nsuper = (int)words.size();
    CK(cudaMalloc(&h->d_slots, (size_t)(h->nsuper + kEphemeralPool) * kFoccFrameBits));
    for (int i = 0; i < h->nsuper; ++i) {
        int rc = focc_upload_frame(h, i, words[(size_t)i].data(), words[(size_t)i].data());
        if (rc != AMPS_OK) return rc;
    }
    h->frame_idx = 0; h->cur_id = 0; h->bit = 0; h->off = 0; h->at_end = false;
    return AMPS_OK;
}

extern "C" int amps_focc_create(unsigned long symrate, int aggressive_registration, int device, amps_focc **out) {
    if (!out) return set_error(AMPS_E_INVAL, "null argument");
    *out = nullptr;
    if (symrate < 20000) return set_error(AMPS_E_INVAL, "symrate must be >= 20000 (samples_per_sym = symrate / 20000)");
    int st = select_device(device);
    if (st != AMPS_OK) return st;
    amps_focc *h = new (std::nothrow) amps_focc();
    if (!h) return set_error(AMPS_E_NOMEM, "out of host memory");
    h->device = device;
    h->sps = (unsigned)(symrate / 20000);
    st = focc_init(h, aggressive_registration);
    if (st != AMPS_OK) { amps_focc_destroy(h); return st; }
    *out = h;
    return AMPS_OK;
}

extern "C" int amps_focc_destroy(amps_focc *h) {
    if (!h) return AMPS_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->ev_gen) cudaEventDestroy(h->ev_gen);
    cudaFree(h->d_slots); cudaFree(h->d_sched); cudaFree(h->d_out);
    delete h;
    return AMPS_OK;
}

extern "C" int amps_focc_set_busy_idle(amps_focc *h, int idle) {
    if (!h) return set_error(AMPS_E_INVAL, "null handle");
    h->busy_idle = idle ? 1 : 0;
    return AMPS_OK;
}

extern "C" int amps_focc_push_words(amps_focc *h, long stream, const uint8_t *words28, long nwords) {
    if (!h || nwords < 0 || (nwords && !words28)) return set_error(AMPS_E_INVAL, "bad argument");
    if (stream < 1 || stream > 3) return set_error(AMPS_E_INVAL, "stream must be 1 (A), 2 (B) or 3 (BOTH)");
    CK(cudaSetDevice(h->device));
    // one slot stays reserved: the frame that was popped and is being transmitted (cur_id) is not in the queue any more
    if ((long)h->queue.size() + nwords >= kEphemeralPool) return set_error(AMPS_E_OVERFLOW, "too many queued FOCC frames");
    // a slot about to be rewritten may still be read by the last generate, which ran on the CALLER's stream
    if (h->gen_pending) { CK(cudaEventSynchronize(h->ev_gen)); h->gen_pending = false; }
    CK(cudaStreamSynchronize(h->stream));
    const Word28 fill = control_filler_word();
    for (long i = 0; i < nwords; ++i) {                                 // one ephemeral frame per word (:529-562)
        const uint8_t *w = words28 + 28 * i;
        const int id = h->nsuper + h->pool_next;
        h->pool_next = (h->pool_next + 1) % kEphemeralPool;
        int rc = focc_upload_frame(h, id, stream == 2 ? fill.data() : w, stream == 1 ? fill.data() : w);
        if (rc != AMPS_OK) return rc;
        h->queue.push_back(id);
    }
    return AMPS_OK;
}

// Plans `n` bytes of output starting at the current state: fills h->h_sched with the slot-table ids of
// the frames crossed, returns the byte offset inside the first one, and advances the state exactly as
// repeated work() calls would.
static unsigned long long focc_plan(amps_focc *h, unsigned long long n) {
    const unsigned long long two = 2ull * h->sps;
    h->h_sched.clear();
    if (h->at_end) focc_step_over_end(h);
    h->h_sched.push_back(h->cur_id);
    const unsigned long long first = (unsigned long long)h->bit * two + (unsigned long long)h->off;
    unsigned long long left = n;
    while (left > 0) {
        if (h->at_end) {
            const bool frame_end = h->bit == kFoccFrameBits;
            focc_step_over_end(h);
            if (frame_end) h->h_sched.push_back(h->cur_id);
        }
        const int last = focc_burst_last_bit(h->bit);
        const unsigned long long avail = (unsigned long long)(last - h->bit + 1) * two - (unsigned long long)h->off;
        const unsigned long long take = left < avail ? left : avail;
        const unsigned long long pos = (unsigned long long)h->off + take;
        h->bit += (int)(pos / two);
        h->off = (int)(pos % two);
        left -= take;
        if (take == avail) h->at_end = true;       // landed exactly on the END marker that closes the burst
    }
    return first;
}

static int focc_emit(amps_focc *h, unsigned long long first, unsigned long long n, uint8_t *d_out, cudaStream_t st) {
    if (h->h_sched.size() > h->sched_cap) {
        cudaFree(h->d_sched); h->d_sched = nullptr;
        h->sched_cap = h->h_sched.size() * 2 + 64;
        CK(cudaMalloc(&h->d_sched, sizeof(int) * h->sched_cap));
    }
    CK(cudaMemcpyAsync(h->d_sched, h->h_sched.data(), sizeof(int) * h->h_sched.size(), cudaMemcpyHostToDevice, st));
    CKL(launch_focc_bytes(h->d_slots, h->d_sched, first, n, h->sps, h->busy_idle, d_out, st));
    CK(cudaEventRecord(h->ev_gen, st));
    h->gen_pending = true;
    return AMPS_OK;
}

static int focc_host_out(amps_focc *h, size_t n) {
    if (n > h->out_cap) {
        cudaFree(h->d_out); h->d_out = nullptr;
        h->out_cap = n * 2 + 4096;
        CK(cudaMalloc(&h->d_out, h->out_cap));
    }
    return AMPS_OK;
}

extern "C" int amps_focc_work(amps_focc *h, uint8_t *out, int noutput_items, int *produced) {
    if (!h || !produced || (noutput_items > 0 && !out)) return set_error(AMPS_E_INVAL, "null argument");
    if (noutput_items < 1) { *produced = -1; return AMPS_OK; }           // WORK_DONE (:590-593)
    *produced = 0;
    CK(cudaSetDevice(h->device));
    if (h->at_end) { focc_step_over_end(h); return AMPS_OK; }            // returns at every FOCC_END, possibly with 0 items (:630-632)
    // one work() call emits at most the rest of the current burst
    const unsigned long long two = 2ull * h->sps;
    const int last = focc_burst_last_bit(h->bit);
    const unsigned long long avail = (unsigned long long)(last - h->bit + 1) * two - (unsigned long long)h->off;
    const unsigned long long take = (unsigned long long)noutput_items < avail ? (unsigned long long)noutput_items : avail;
    const unsigned long long first = focc_plan(h, take);
    int rc = focc_host_out(h, (size_t)take);
    if (rc != AMPS_OK) return rc;
    rc = focc_emit(h, first, take, h->d_out, h->stream);
    if (rc != AMPS_OK) return rc;
    CK(cudaMemcpyAsync(out, h->d_out, (size_t)take, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->at_end && take < (unsigned long long)noutput_items) focc_step_over_end(h);   // the loop reached FOCC_END inside this call
    *produced = (int)take;
    return AMPS_OK;
}

extern "C" int amps_focc_generate_dev(amps_focc *h, void *d_out, size_t n, void *cuda_stream) {
    if (!h || (n && !d_out)) return set_error(AMPS_E_INVAL, "null argument");
    if (n == 0) return AMPS_OK;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    CK(cudaStreamSynchronize(st));                                       // h_sched / d_sched are reused
    const unsigned long long first = focc_plan(h, n);
    return focc_emit(h, first, n, static_cast<uint8_t *>(d_out), st);
}

// The same stream as data bits (1 byte per bit): advances the state exactly as generating nbits * 2 * sps bytes would.
extern "C" int amps_focc_generate_bits_dev(amps_focc *h, void *d_out, size_t nbits, void *cuda_stream) {
    if (!h || (nbits && !d_out)) return set_error(AMPS_E_INVAL, "null argument");
    if (nbits == 0) return AMPS_OK;
    if (h->off != 0) return set_error(AMPS_E_STATE, "the stream is in the middle of a bit (a byte-level call stopped there)");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    CK(cudaStreamSynchronize(st));
    const unsigned long long two = 2ull * h->sps;
    const unsigned long long first = focc_plan(h, (unsigned long long)nbits * two);
    if (h->h_sched.size() > h->sched_cap) {
        cudaFree(h->d_sched); h->d_sched = nullptr;
        h->sched_cap = h->h_sched.size() * 2 + 64;
        CK(cudaMalloc(&h->d_sched, sizeof(int) * h->sched_cap));
    }
    CK(cudaMemcpyAsync(h->d_sched, h->h_sched.data(), sizeof(int) * h->h_sched.size(), cudaMemcpyHostToDevice, st));
    CKL(launch_focc_bits(h->d_slots, h->d_sched, first / two, nbits, h->busy_idle, static_cast<uint8_t *>(d_out), st));
    CK(cudaEventRecord(h->ev_gen, st));
    h->gen_pending = true;
    return AMPS_OK;
}

extern "C" int amps_focc_generate_bits(amps_focc *h, uint8_t *out, size_t nbits) {
    if (!h || (nbits && !out)) return set_error(AMPS_E_INVAL, "null argument");
    if (nbits == 0) return AMPS_OK;
    CK(cudaSetDevice(h->device));
    int rc = focc_host_out(h, nbits);
    if (rc != AMPS_OK) return rc;
    rc = amps_focc_generate_bits_dev(h, h->d_out, nbits, h->stream);
    if (rc != AMPS_OK) return rc;
    CK(cudaMemcpyAsync(out, h->d_out, nbits, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return AMPS_OK;
}

extern "C" int amps_focc_generate(amps_focc *h, uint8_t *out, size_t n) {
    if (!h || (n && !out)) return set_error(AMPS_E_INVAL, "null argument");
    if (n == 0) return AMPS_OK;
    CK(cudaSetDevice(h->device));
    int rc = focc_host_out(h, n);
    if (rc != AMPS_OK) return rc;
    rc = amps_focc_generate_dev(h, h->d_out, n, h->stream);
    if (rc != AMPS_OK) return rc;
    CK(cudaMemcpyAsync(out, h->d_out, n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return AMPS_OK;
}

// =============================================================================================
// fvc (lib/fvc_impl.cc)
// =============================================================================================
struct amps_fvc {
    int device = 0;
    cudaStream_t stream = nullptr;
    unsigned sps = 1;
    std::vector<uint8_t> bits;             // d_curdata as bits; only ever grows (:131-142)
    uint8_t *d_bits = nullptr; size_t bits_cap = 0; size_t bits_on_dev = 0;
    unsigned long long replay_len = 0, replay_pos = 0;     // bytes of the snapshot being replayed (d_curqueue)
    uint64_t timer = 0;                    // timerhack
    uint8_t *d_out = nullptr; size_t out_cap = 0;
};

extern "C" int amps_fvc_create(unsigned long symrate, int device, amps_fvc **out) {
    if (!out) return set_error(AMPS_E_INVAL, "null argument");
    *out = nullptr;
    if (symrate < 20000) return set_error(AMPS_E_INVAL, "symrate must be >= 20000");
    int st = select_device(device);
    if (st != AMPS_OK) return st;
    amps_fvc *h = new (std::nothrow) amps_fvc();
    if (!h) return set_error(AMPS_E_NOMEM, "out of host memory");
    h->device = device;
    h->sps = (unsigned)(symrate / 20000);
    const cudaError_t ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) { delete h; return set_cuda_error(ce, "cudaStreamCreateWithFlags"); }
    *out = h;
    return AMPS_OK;
}

extern "C" int amps_fvc_destroy(amps_fvc *h) {
    if (!h) return AMPS_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaFree(h->d_bits); cudaFree(h->d_out);
    delete h;
    return AMPS_OK;
}

extern "C" int amps_fvc_push_words(amps_fvc *h, const uint8_t *words28, long nwords, int has_timer, uint64_t timer) {
    if (!h || nwords < 0 || (nwords && !words28)) return set_error(AMPS_E_INVAL, "bad argument");
    if (has_timer) h->timer = timer;                                     // (:124-127)
    for (long i = 0; i < nwords; ++i) {
        const auto t = fvc_word_train(words28 + 28 * i);
        h->bits.insert(h->bits.end(), t.begin(), t.end());
    }
    return AMPS_OK;
}

extern "C" int amps_fvc_work(amps_fvc *h, uint8_t *out, int noutput_items, int *produced, int *fvc_off) {
    if (!h || !produced || (noutput_items > 0 && !out)) return set_error(AMPS_E_INVAL, "null argument");
    if (fvc_off) *fvc_off = 0;
    *produced = 0;
    if (noutput_items < 1) return AMPS_OK;
    CK(cudaSetDevice(h->device));
    if (h->bits.empty()) {
        // no word was ever queued: claim noutput_items and leave the buffer untouched, exactly as the reference does
        // (:159-161; the stream is muted downstream, grc/ampsbs.grc:1555-1601)
        *produced = noutput_items;
        return AMPS_OK;
    }
    if (h->replay_pos == h->replay_len) {                               // replay queue empty: new snapshot (:162-173)
        if (h->timer >= 1) {
            h->timer--;
            if (h->timer == 0 && fvc_off) *fvc_off = 1;                  // "fvc off" PDU on command_out
        }
        h->replay_len = (unsigned long long)h->bits.size() * 2ull * h->sps;
        h->replay_pos = 0;
        if (h->bits.size() > h->bits_on_dev) {                           // bring the grown train to the device
            if (h->bits.size() > h->bits_cap) {
                cudaFree(h->d_bits); h->d_bits = nullptr;
                h->bits_cap = h->bits.size() * 2 + 4096;
                CK(cudaMalloc(&h->d_bits, h->bits_cap));
            }
            CK(cudaMemcpy(h->d_bits, h->bits.data(), h->bits.size(), cudaMemcpyHostToDevice));
            h->bits_on_dev = h->bits.size();
        }
    }
    const unsigned long long left = h->replay_len - h->replay_pos;
    const unsigned long long take = (unsigned long long)noutput_items < left ? (unsigned long long)noutput_items : left;
    if (take > h->out_cap) {
        cudaFree(h->d_out); h->d_out = nullptr;
        h->out_cap = (size_t)take * 2 + 4096;
        CK(cudaMalloc(&h->d_out, h->out_cap));
    }
    CKL(launch_fvc_bytes(h->d_bits, h->replay_pos, take, h->sps, h->d_out, h->stream));
    CK(cudaMemcpyAsync(out, h->d_out, (size_t)take, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->replay_pos += take;
    *produced = (int)take;
    return AMPS_OK;
}

// fvc as DATA BITS (one byte per bit, 0xFF = muted while no word was ever pushed): same replay / timerhack semantics as
// amps_fvc_work, in units of bits; only valid on a bit boundary of the replay.  The bits are control data that already
// live on the host (the word train), so this is a copy, not a kernel.
extern "C" int amps_fvc_work_bits(amps_fvc *h, uint8_t *out_bits, int nbits, int *produced, int *fvc_off) {
    if (!h || !produced || (nbits > 0 && !out_bits)) return set_error(AMPS_E_INVAL, "null argument");
    if (fvc_off) *fvc_off = 0;
    *produced = 0;
    if (nbits < 1) return AMPS_OK;
    const unsigned long long two = 2ull * h->sps;
    if (h->bits.empty()) {
        std::memset(out_bits, 0xFF, (size_t)nbits);
        *produced = nbits;
        return AMPS_OK;
    }
    if (h->replay_pos % two) return set_error(AMPS_E_STATE, "the replay is in the middle of a bit (a byte-level call stopped there)");
    if (h->replay_pos == h->replay_len) {
        if (h->timer >= 1) {
            h->timer--;
            if (h->timer == 0 && fvc_off) *fvc_off = 1;
        }
        h->replay_len = (unsigned long long)h->bits.size() * two;
        h->replay_pos = 0;
    }
    const unsigned long long left = (h->replay_len - h->replay_pos) / two;
    const unsigned long long take = (unsigned long long)nbits < left ? (unsigned long long)nbits : left;
    std::memcpy(out_bits, h->bits.data() + h->replay_pos / two, (size_t)take);
    h->replay_pos += take * two;
    *produced = (int)take;
    return AMPS_OK;
}
