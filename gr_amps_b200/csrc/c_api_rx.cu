// c_api_rx.cu -- extern "C" entry points of the receive side: amps_recc_iq_* (fused IQ path) and
// amps_recc_decode_* (message-only burst decoder).  See include/amps_b200.h for the contract and
// the reference interfaces each entry point replaces.
#include "common.h"
#include "design.h"
#include "rx_kernels.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

using namespace amps;

struct amps_recc_iq {
    int          device = 0;
    int          sm_count = 0;
    cudaStream_t stream = nullptr;       // own stream for the host-buffer path
    uint32_t     max_samples = 0;
    uint32_t     max_records = 0;
    uint32_t     flags = 0;
    std::vector<float> lpf;
    uint32_t     fcw = 0;
    bool         native400 = false;      // samp_rate == 400e3: the reference's own rate, no CIC stage
    uint32_t     pass_in = kPass;        // input samples per pass (API granularity)
    uint32_t     hist = kHist;           // input samples of history carried between calls
    uint32_t     decim = kD1 * kD2;      // input samples per demodulated sample

    RxFrontParams fp{};                  // constant part filled at create
    bool         sc16 = false;           // AMPS_RX_INPUT_SC16: interleaved int16 I/Q instead of float
    bool         sc16_unit = false;      // sc16 scale is a power of two: folded into the NCO tables, no multiply in the kernel
    size_t       isz = sizeof(float2);   // bytes per complex input sample (8 or 4)
    uint8_t     *d_stage = nullptr;      // host path: [carry | new chunk]
    uint8_t     *d_tail[2] = {nullptr, nullptr};
    int          tail_cur = 0;
    float       *d_dring = nullptr;
    uint32_t    *d_hring = nullptr;      // hard decisions (d >= 0), 1 bit per demod sample, same ring indexing
    uint32_t     dmask = 0;
    float2      *d_ydump = nullptr;
    size_t       ydump_cap = 0;          // complex samples
    uint64_t     ydump_first = 0, ydump_count = 0;
    RxState     *d_state = nullptr;
    Candidate   *d_cand = nullptr;
    Accepted    *d_acc = nullptr;        // bursts accepted by the last select (kMaxAccept entries)
    amps_burst  *h_ring = nullptr;       // mapped pinned host ring the select kernel publishes into
    RxPublished *h_pub = nullptr;        // mapped pinned counters
    uint64_t     consumed = 0;           // bursts already handed to the caller
    uint64_t     lost = 0;               // bursts overwritten in the ring before they were collected
    cudaStream_t last_stream = nullptr;
    cudaStream_t side = nullptr;         // detect / select / capture run here, overlapped with the next front kernel
    cudaEvent_t  ev_front = nullptr;     // front kernel of the current call finished
    cudaEvent_t  ev_side[2] = {nullptr, nullptr};   // detection of call k finished (k & 1)
    uint64_t     call_no = 0;
    bool         serial = false;         // AMPS_RX_SERIAL=1: no overlap (profiling / A-B measurements)
    int          diag = 0;               // AMPS_RX_DIAG (measurement aid): 1 = no capture launch, 2 = no detect/select launch
    bool         front_only = false;     // AMPS_RX_FRONT_ONLY=1: no detection at all (pipeline measurements only: no bursts come out)
    // AMPS_RX_TIMING_MM: the reference graph's serial tail instead of the feed-forward detector
    bool         mm_mode = false;
    MmState     *d_mm = nullptr;
    float       *d_mmtab = nullptr;
    uint8_t     *d_sym = nullptr;
    uint32_t     sym_cap = 0;
    ReccCompatState *d_compat = nullptr;
    uint8_t     *d_blobs = nullptr;
    unsigned long long *d_blob_idx = nullptr;

    size_t       carry = 0;              // unprocessed samples sitting at the front of d_stage
    uint64_t     samples_in = 0;         // samples handed to the kernels
    uint64_t     total_d = 0;            // demod samples produced
    uint64_t     scan_hi = 0;            // positions below this have been searched
    uint64_t     bursts = 0, launches = 0;
    // AMPS_RX_TIME_KERNELS: ring of event pairs around the front-end kernel
    static constexpr int kEv = 256;
    cudaEvent_t  ev0[kEv] = {}, ev1[kEv] = {};
    uint64_t     ev_count = 0;
};

static int mm_reset(amps_recc_iq *h) {
    MmState m;
    std::memset(&m, 0, sizeof m);
    m.omega = 10.0f;                                                   // grc/ampsbs.grc:1807 (omega = 10, mu = 0)
    CK(cudaMemcpy(h->d_mm, &m, sizeof m, cudaMemcpyHostToDevice));
    CK(cudaMemset(h->d_compat, 0, sizeof(ReccCompatState)));
    const int32_t none = -1;
    CK(cudaMemcpy(&h->d_compat->pending, &none, sizeof none, cudaMemcpyHostToDevice));
    return AMPS_OK;
}

static int rx_alloc(amps_recc_iq *h) {
    const size_t max_d = (size_t)h->max_samples / h->decim + kPassOut;
    size_t cap = 1;
    // two calls' worth: the detection of call k overlaps the front kernel of call k+1
    while (cap < 2 * max_d + (size_t)kSpan + 4096) cap <<= 1;
    h->dmask = (uint32_t)(cap - 1);
    CK(cudaMalloc(&h->d_stage, ((size_t)h->max_samples + h->pass_in) * h->isz));
    for (int i = 0; i < 2; ++i) {
        CK(cudaMalloc(&h->d_tail[i], (size_t)h->hist * h->isz));
        CK(cudaMemset(h->d_tail[i], 0, (size_t)h->hist * h->isz));
    }
    CK(cudaMalloc(&h->d_dring, cap * sizeof(float)));
    CK(cudaMemset(h->d_dring, 0, cap * sizeof(float)));
    CK(cudaMalloc(&h->d_hring, cap / 8));
    CK(cudaMemset(h->d_hring, 0, cap / 8));
    if (h->flags & AMPS_RX_DUMP_BASEBAND) {
        h->ydump_cap = max_d;
        CK(cudaMalloc(&h->d_ydump, h->ydump_cap * sizeof(float2)));
    }
    CK(cudaMalloc(&h->d_state, sizeof(RxState)));
    CK(cudaMemset(h->d_state, 0, sizeof(RxState)));
    CK(cudaMalloc(&h->d_cand, sizeof(Candidate) * kMaxCand * 2));      // candidates + the select kernel's sorted copy
    CK(cudaMalloc(&h->d_acc, sizeof(Accepted) * kMaxAccept));
    if (h->mm_mode) {
        h->sym_cap = (uint32_t)(max_d / 8 + 64);
        const std::vector<float> tab = mmse_interp_table();
        CK(cudaMalloc(&h->d_mm, sizeof(MmState)));
        CK(cudaMalloc(&h->d_mmtab, tab.size() * sizeof(float)));
        CK(cudaMemcpy(h->d_mmtab, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&h->d_sym, h->sym_cap));
        CK(cudaMalloc(&h->d_compat, sizeof(ReccCompatState)));
        CK(cudaMalloc(&h->d_blobs, (size_t)kMaxAccept * kCapture));
        CK(cudaMalloc(&h->d_blob_idx, (size_t)kMaxAccept * sizeof(unsigned long long)));
        int rc = mm_reset(h);
        if (rc != AMPS_OK) return rc;
    }
    CK(cudaHostAlloc(&h->h_ring, sizeof(amps_burst) * h->max_records, cudaHostAllocMapped));
    CK(cudaHostAlloc(&h->h_pub, sizeof(RxPublished), cudaHostAllocMapped));
    std::memset(h->h_pub, 0, sizeof(RxPublished));
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_front, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&h->ev_side[i], cudaEventDisableTiming));
    if (h->flags & AMPS_RX_TIME_KERNELS)
        for (int i = 0; i < amps_recc_iq::kEv; ++i) { CK(cudaEventCreate(&h->ev0[i])); CK(cudaEventCreate(&h->ev1[i])); }
    return AMPS_OK;
}

extern "C" int amps_recc_iq_create(const amps_recc_iq_params *params, amps_recc_iq **out) {
    if (!params || !out) return set_error(AMPS_E_INVAL, "null argument");
    *out = nullptr;
    if (params->samp_rate != 10e6 && params->samp_rate != 400e3)
        return set_error(AMPS_E_INVAL, "samp_rate must be 10e6 (25 x the reference's rate) or 400e3 (the reference's own rate)");
    if (params->max_samples == 0) return set_error(AMPS_E_INVAL, "max_samples must be > 0");
    if (params->lpf_taps && (params->n_lpf_taps == 0 || params->n_lpf_taps > (uint32_t)kMaxLpf))
        return set_error(AMPS_E_INVAL, "n_lpf_taps must be in 1..299");
    int st = select_device(params->device);
    if (st != AMPS_OK) return st;
    amps_recc_iq *h = new (std::nothrow) amps_recc_iq();
    if (!h) return set_error(AMPS_E_NOMEM, "out of host memory");
    h->device = params->device;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, params->device);
    h->sm_count = prop.multiProcessorCount;
    // round the per-call capacity up to whole passes
    h->native400 = params->samp_rate == 400e3;
    if (h->native400) { h->pass_in = kPass400; h->hist = kPass400; h->decim = kD2; }
    h->max_samples = (uint32_t)(((uint64_t)params->max_samples + h->pass_in - 1) / h->pass_in * h->pass_in);
    h->max_records = params->max_bursts ? params->max_bursts : 256;
    h->flags = params->flags;
    h->mm_mode = (params->flags & AMPS_RX_TIMING_MM) != 0;
    h->sc16 = (params->flags & AMPS_RX_INPUT_SC16) != 0;
    h->isz = h->sc16 ? sizeof(short2) : sizeof(float2);
    { const char *e = std::getenv("AMPS_RX_SERIAL"); h->serial = e && e[0] == '1'; }
    { const char *e = std::getenv("AMPS_RX_FRONT_ONLY"); h->front_only = e && e[0] == '1'; }
    { const char *e = std::getenv("AMPS_RX_DIAG"); h->diag = e ? std::atoi(e) : 0; }
    if (params->lpf_taps) h->lpf.assign(params->lpf_taps, params->lpf_taps + params->n_lpf_taps);
    else h->lpf = firdes_low_pass(3.0, 400e3, 10e3, 4500.0, WIN_BLACKMAN);     // grc/ampsbs.grc:138-184
    h->fcw = nco_fcw(params->center_freq, params->samp_rate);

    std::memset(&h->fp, 0, sizeof h->fp);
    h->fp.fcw25 = (uint32_t)(25u * h->fcw);
    h->fp.in_scale = params->sc16_scale != 0.0f ? params->sc16_scale : 1.0f / 32768.0f;
    nco_block_table(h->fcw, kD1, reinterpret_cast<float *>(h->fp.w));
    if (h->sc16) {
        int e = 0;
        const float m = std::frexp(h->fp.in_scale, &e);
        if (m == 0.5f && e > -40 && e < 40) {                       // power of two: (I s) w == I (s w) exactly
            h->sc16_unit = true;
            for (int k = 0; k < kD1; ++k) { h->fp.w[k].x *= h->fp.in_scale; h->fp.w[k].y *= h->fp.in_scale; }
        }
    }
    for (int k = 0; k < kD1; ++k) h->fp.wj[k] = make_float2(-h->fp.w[k].y, h->fp.w[k].x);
    std::vector<float> cic;
    cic3_taps(kD1, cic);
    for (size_t i = 0; i < cic.size(); ++i) h->fp.g[i] = cic[i];
    for (size_t i = 0; i < h->lpf.size(); ++i) h->fp.h2[i] = h->lpf[i];

    cudaError_t ce = rx_configure_device();
    if (ce != cudaSuccess) { delete h; return set_cuda_error(ce, "rx_configure_device"); }
    st = rx_alloc(h);
    if (st != AMPS_OK) { amps_recc_iq_destroy(h); return st; }
    *out = h;
    return AMPS_OK;
}

extern "C" int amps_recc_iq_destroy(amps_recc_iq *h) {
    if (!h) return AMPS_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->ev_front) cudaEventDestroy(h->ev_front);
    for (int i = 0; i < 2; ++i) if (h->ev_side[i]) cudaEventDestroy(h->ev_side[i]);
    for (int i = 0; i < amps_recc_iq::kEv; ++i) { if (h->ev0[i]) cudaEventDestroy(h->ev0[i]); if (h->ev1[i]) cudaEventDestroy(h->ev1[i]); }
    cudaFree(h->d_stage); cudaFree(h->d_tail[0]); cudaFree(h->d_tail[1]); cudaFree(h->d_dring); cudaFree(h->d_hring);
    cudaFree(h->d_ydump); cudaFree(h->d_state); cudaFree(h->d_cand); cudaFree(h->d_acc);
    cudaFree(h->d_mm); cudaFree(h->d_mmtab); cudaFree(h->d_sym); cudaFree(h->d_compat); cudaFree(h->d_blobs); cudaFree(h->d_blob_idx);
    if (h->h_ring) cudaFreeHost(h->h_ring);
    if (h->h_pub) cudaFreeHost(h->h_pub);
    delete h;
    return AMPS_OK;
}

extern "C" int amps_recc_iq_reset(amps_recc_iq *h) {
    if (!h) return set_error(AMPS_E_INVAL, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    for (int i = 0; i < 2; ++i) CK(cudaMemset(h->d_tail[i], 0, (size_t)h->hist * h->isz));
    CK(cudaMemset(h->d_state, 0, sizeof(RxState)));
    CK(cudaMemset(h->d_dring, 0, ((size_t)h->dmask + 1) * sizeof(float)));
    if (h->mm_mode) { int rc = mm_reset(h); if (rc != AMPS_OK) return rc; }
    std::memset(h->h_pub, 0, sizeof(RxPublished));
    h->consumed = 0; h->call_no = 0;
    h->tail_cur = 0; h->carry = 0; h->samples_in = 0; h->total_d = 0; h->scan_hi = 0;
    h->ydump_first = 0; h->ydump_count = 0;
    return AMPS_OK;
}

extern "C" int amps_recc_iq_granularity(const amps_recc_iq *h) { return h ? (int)h->pass_in : kPass; }

// Enqueue everything for `npass` passes whose samples start at d_chunk.
static int rx_enqueue(amps_recc_iq *h, const uint8_t *d_chunk, uint32_t npass, cudaStream_t st) {
    RxFrontParams p = h->fp;
    p.chunk = d_chunk;
    p.tail = h->d_tail[h->tail_cur];
    p.dring = h->d_dring;
    p.hring = h->d_hring;
    p.dmask = h->dmask;
    p.q_base = h->total_d;
    p.npass = npass;
    p.blk_base = (uint32_t)(h->samples_in / kD1);
    p.n_base = h->samples_in;
    p.ydump = h->d_ydump;
    // whole passes per CTA, grid sized so that (nearly) every CTA gets the same count within one wave
    const uint32_t resident = (uint32_t)rx_front_ctas_per_sm(h->sc16) * (uint32_t)h->sm_count;
    p.pass_per_cta = (npass + resident - 1u) / resident;
    const int grid = (int)((npass + p.pass_per_cta - 1u) / p.pass_per_cta);
    // the demod ring holds two calls: do not overwrite what the detection of call k-2 may still read
    const int par = (int)(h->call_no & 1u);
    if (h->call_no >= 2) CK(cudaStreamWaitEvent(st, h->ev_side[par], 0));
    const bool timed = (h->flags & AMPS_RX_TIME_KERNELS) != 0;
    const int  evi = (int)(h->ev_count % amps_recc_iq::kEv);
    if (timed) CK(cudaEventRecord(h->ev0[evi], st));
    p.tail_out = h->native400 ? nullptr : h->d_tail[h->tail_cur ^ 1];
    if (h->native400) CKL(launch_rx_front400(p, grid, st, h->sc16, h->sc16_unit));
    else CKL(launch_rx_front(p, grid, st, h->sc16, h->sc16_unit));
    if (timed) { CK(cudaEventRecord(h->ev1[evi], st)); h->ev_count++; }
    h->launches++;
    // history for the next call = the tail of this one (the 10 MS/s front kernel copies it itself)
    if (h->native400)
        CK(cudaMemcpyAsync(h->d_tail[h->tail_cur ^ 1], d_chunk + ((size_t)npass * h->pass_in - h->hist) * h->isz, (size_t)h->hist * h->isz,
                           cudaMemcpyDeviceToDevice, st));
    h->tail_cur ^= 1;
    h->ydump_first = h->total_d;
    h->ydump_count = (uint64_t)npass * kPassOut;
    h->samples_in += (uint64_t)npass * h->pass_in;
    h->total_d += (uint64_t)npass * kPassOut;
    // search every position whose capture is complete -- on the side stream, so that the next call's
    // front kernel (HBM-bound, 2 CTAs/SM) overlaps these small latency-bound kernels
    CK(cudaEventRecord(h->ev_front, st));
    cudaStream_t sd = h->serial ? st : h->side;
    if (!h->serial) CK(cudaStreamWaitEvent(sd, h->ev_front, 0));
    if (h->front_only) {
        // measurement aid: nothing after the front kernel
    } else if (h->mm_mode) {
        // serial tail: M&M + slicer over everything demodulated so far, amps.recc on the new half-symbols, then
        // one CTA per blob decodes and publishes it
        CKL(launch_rx_mm(h->d_dring, h->dmask, h->total_d, h->d_mm, h->d_mmtab, h->d_sym, h->sym_cap, h->d_compat, h->d_blobs,
                         h->d_blob_idx, kMaxAccept, h->d_state, h->h_pub, sd));
        int max_new = (int)((uint64_t)npass * kPassOut / (8u * (unsigned)kMmQuantum)) + 2;   // <= one blob per work() quantum
        if (max_new > kMaxAccept) max_new = kMaxAccept;
        CKL(launch_rx_capture(h->d_dring, h->dmask, h->d_state, h->d_acc, max_new, h->h_ring, h->max_records, h->h_pub, h->decim, sd,
                              h->d_blobs, h->d_blob_idx));
        h->launches += 3;
    } else if (h->total_d > (uint64_t)kSpan) {
        const uint64_t hi = h->total_d - (uint64_t)kSpan;
        const uint64_t lo = h->scan_hi > 64 ? h->scan_hi - 64 : 0;
        if (hi > h->scan_hi) {
            if (!(h->diag & 2)) {
            CKL(launch_rx_detect(h->d_dring, h->d_hring, h->dmask, h->d_state, h->d_cand, lo, hi, 0, sd));
            CKL(launch_rx_select(h->d_state, h->d_cand, h->d_acc, hi, h->h_pub, sd));
            }
            // at most one burst per kBurstLen searched positions (+1 for a run deferred from the last call)
            const int max_new = (int)((hi - lo) / (uint64_t)kBurstLen) + 2;
            if (!(h->diag & 1))
            CKL(launch_rx_capture(h->d_dring, h->dmask, h->d_state, h->d_acc, max_new, h->h_ring, h->max_records, h->h_pub, h->decim, sd));
            h->launches += 3;
            h->scan_hi = hi;
        }
    }
    CK(cudaEventRecord(h->ev_side[par], sd));
    h->call_no++;
    h->last_stream = st;
    return AMPS_OK;
}

static int rx_submit_dev(amps_recc_iq *h, const void *d_iq, size_t nsamples, void *cuda_stream, bool sc16) {
    if (!h || (!d_iq && nsamples)) return set_error(AMPS_E_INVAL, "null argument");
    if (h->sc16 != sc16) return set_error(AMPS_E_STATE, sc16 ? "handle was not created with AMPS_RX_INPUT_SC16" : "handle was created with AMPS_RX_INPUT_SC16: use the _sc16 entry points");
    if (nsamples == 0) return AMPS_OK;
    if (nsamples % h->pass_in) return set_error(AMPS_E_ALIGN, "nsamples must be a multiple of amps_recc_iq_granularity()");
    if (reinterpret_cast<uintptr_t>(d_iq) & 15u) return set_error(AMPS_E_ALIGN, "d_iq must be 16-byte aligned");
    if (nsamples > h->max_samples) return set_error(AMPS_E_OVERFLOW, "nsamples exceeds max_samples");
    if (h->carry) return set_error(AMPS_E_STATE, "host-path samples are pending; reset() or keep using work()");
    CK(cudaSetDevice(h->device));
    return rx_enqueue(h, static_cast<const uint8_t *>(d_iq), (uint32_t)(nsamples / h->pass_in), static_cast<cudaStream_t>(cuda_stream));
}
extern "C" int amps_recc_iq_submit_dev(amps_recc_iq *h, const void *d_iq, size_t nsamples, void *cuda_stream) {
    return rx_submit_dev(h, d_iq, nsamples, cuda_stream, false);
}
extern "C" int amps_recc_iq_submit_sc16_dev(amps_recc_iq *h, const void *d_iq, size_t nsamples, void *cuda_stream) {
    return rx_submit_dev(h, d_iq, nsamples, cuda_stream, true);
}

// Wait for the stream; afterwards records [h->consumed, h->consumed + *n_out) sit in the host ring.
static int rx_fetch(amps_recc_iq *h, uint64_t *n_out) {
    *n_out = 0;
    CK(cudaStreamSynchronize(h->side));
    CK(cudaStreamSynchronize(h->last_stream));
    if (h->h_pub->cand_overflow)
        return set_error(AMPS_E_OVERFLOW, h->mm_mode ? "more than 512 bursts captured in one call"
                                                     : "trigger candidate list overflowed (more than 8192 matches in one call)");
    const uint64_t total = h->h_pub->nrec_total;
    if (total - h->consumed > h->max_records) {           // the ring wrapped over uncollected records
        h->lost += total - h->consumed - h->max_records;
        h->consumed = total - h->max_records;
    }
    *n_out = total - h->consumed;
    return AMPS_OK;
}

extern "C" int amps_recc_iq_collect(amps_recc_iq *h, amps_burst *out, int max, int *n_out) {
    if (!h || !n_out || (max > 0 && !out)) return set_error(AMPS_E_INVAL, "null argument");
    *n_out = 0;
    CK(cudaSetDevice(h->device));
    uint64_t n = 0;
    int rc = rx_fetch(h, &n);
    if (rc != AMPS_OK) return rc;
    const uint64_t give = n < (uint64_t)max ? n : (uint64_t)max;     // the rest stays for the next collect
    for (uint64_t i = 0; i < give; ++i) out[i] = h->h_ring[(h->consumed + i) % h->max_records];
    h->consumed += give;
    h->bursts += give;
    *n_out = (int)give;
    return AMPS_OK;
}

extern "C" int amps_recc_iq_peek(amps_recc_iq *h, const amps_burst **ring, uint32_t *ring_len, uint64_t *first, uint64_t *count) {
    if (!h || !ring || !ring_len || !first || !count) return set_error(AMPS_E_INVAL, "null argument");
    CK(cudaSetDevice(h->device));
    uint64_t n = 0;
    int rc = rx_fetch(h, &n);
    if (rc != AMPS_OK) return rc;
    *ring = h->h_ring; *ring_len = h->max_records; *first = h->consumed; *count = n;
    return AMPS_OK;
}

extern "C" int amps_recc_iq_poll(amps_recc_iq *h, const amps_burst **ring, uint32_t *ring_len, uint64_t *first, uint64_t *count) {
    if (!h || !ring || !ring_len || !first || !count) return set_error(AMPS_E_INVAL, "null argument");
    const uint64_t total = *reinterpret_cast<volatile unsigned long long *>(&h->h_pub->nrec_total);
    if (total - h->consumed > h->max_records) {           // the ring wrapped over uncollected records
        h->lost += total - h->consumed - h->max_records;
        h->consumed = total - h->max_records;
    }
    *ring = h->h_ring; *ring_len = h->max_records; *first = h->consumed; *count = total - h->consumed;
    return AMPS_OK;
}

extern "C" int amps_recc_iq_consume(amps_recc_iq *h, uint64_t count) {
    if (!h) return set_error(AMPS_E_INVAL, "null handle");
    if (h->consumed + count > h->h_pub->nrec_total) return set_error(AMPS_E_INVAL, "consuming more bursts than were published");
    h->consumed += count;
    h->bursts += count;
    return AMPS_OK;
}

static int rx_work(amps_recc_iq *h, const void *iq_host, size_t nsamples, amps_burst_cb cb, void *user, bool sc16) {
    if (!h || (!iq_host && nsamples)) return set_error(AMPS_E_INVAL, "null argument");
    if (h->sc16 != sc16) return set_error(AMPS_E_STATE, sc16 ? "handle was not created with AMPS_RX_INPUT_SC16" : "handle was created with AMPS_RX_INPUT_SC16: use the _sc16 entry points");
    if (nsamples > h->max_samples) return set_error(AMPS_E_OVERFLOW, "nsamples exceeds max_samples");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    if (nsamples)
        CK(cudaMemcpyAsync(h->d_stage + h->carry * h->isz, iq_host, nsamples * h->isz, cudaMemcpyHostToDevice, st));
    const size_t avail = h->carry + nsamples;
    const uint32_t npass = (uint32_t)(avail / h->pass_in);
    h->last_stream = st;
    if (npass) {
        int rc = rx_enqueue(h, h->d_stage, npass, st);
        if (rc != AMPS_OK) return rc;
        const size_t left = avail - (size_t)npass * h->pass_in;
        if (left)
            CK(cudaMemcpyAsync(h->d_stage, h->d_stage + (size_t)npass * h->pass_in * h->isz, left * h->isz, cudaMemcpyDeviceToDevice, st));
        h->carry = left;
    } else {
        h->carry = avail;
    }
    // deliver bursts in stream order, like message_port_pub("bursts", ...) from work() (lib/recc_impl.cc:126);
    // the callback sees the record in place in the pinned ring
    uint64_t n = 0;
    int rc = rx_fetch(h, &n);
    if (rc != AMPS_OK) return rc;
    if (cb) for (uint64_t i = 0; i < n; ++i) cb(&h->h_ring[(h->consumed + i) % h->max_records], user);
    h->consumed += n;
    h->bursts += n;
    return AMPS_OK;
}
extern "C" int amps_recc_iq_work(amps_recc_iq *h, const float *iq_host, size_t nsamples, amps_burst_cb cb, void *user) {
    return rx_work(h, iq_host, nsamples, cb, user, false);
}
extern "C" int amps_recc_iq_work_sc16(amps_recc_iq *h, const int16_t *iq_host, size_t nsamples, amps_burst_cb cb, void *user) {
    return rx_work(h, iq_host, nsamples, cb, user, true);
}

extern "C" int amps_recc_iq_read_demod(amps_recc_iq *h, uint64_t first, float *out, size_t n) {
    if (!h || !out) return set_error(AMPS_E_INVAL, "null argument");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    if (first + n > h->total_d) return set_error(AMPS_E_INVAL, "range beyond the demodulated stream");
    if (h->total_d - first > (uint64_t)h->dmask + 1) return set_error(AMPS_E_INVAL, "range no longer in the demod ring");
    size_t done = 0;
    while (done < n) {
        const size_t idx = (size_t)((first + done) & h->dmask);
        size_t run = (size_t)h->dmask + 1 - idx;
        if (run > n - done) run = n - done;
        CK(cudaMemcpy(out + done, h->d_dring + idx, run * sizeof(float), cudaMemcpyDeviceToHost));
        done += run;
    }
    return AMPS_OK;
}

extern "C" int amps_recc_iq_read_baseband(amps_recc_iq *h, uint64_t first, float *out_iq, size_t n) {
    if (!h || !out_iq) return set_error(AMPS_E_INVAL, "null argument");
    if (!h->d_ydump) return set_error(AMPS_E_STATE, "handle was not created with AMPS_RX_DUMP_BASEBAND");
    if (first < h->ydump_first || first + n > h->ydump_first + h->ydump_count)
        return set_error(AMPS_E_INVAL, "only the last call's baseband is kept");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out_iq, h->d_ydump + (first - h->ydump_first), n * sizeof(float2), cudaMemcpyDeviceToHost));
    return AMPS_OK;
}

extern "C" int amps_recc_iq_stats(const amps_recc_iq *h, uint64_t *samples_in, uint64_t *demod_out, uint64_t *bursts,
                                  uint64_t *kernel_launches) {
    if (!h) return set_error(AMPS_E_INVAL, "null handle");
    if (samples_in) *samples_in = h->samples_in;
    if (demod_out) *demod_out = h->total_d;
    if (bursts) *bursts = h->bursts;
    if (kernel_launches) *kernel_launches = h->launches;
    return AMPS_OK;
}

extern "C" int amps_recc_iq_front_times(amps_recc_iq *h, float *ms_out, int cap, int *n_out) {
    if (!h || !n_out || (cap > 0 && !ms_out)) return set_error(AMPS_E_INVAL, "null argument");
    *n_out = 0;
    if (!(h->flags & AMPS_RX_TIME_KERNELS)) return set_error(AMPS_E_STATE, "handle was not created with AMPS_RX_TIME_KERNELS");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->last_stream));
    uint64_t have = h->ev_count < (uint64_t)amps_recc_iq::kEv ? h->ev_count : (uint64_t)amps_recc_iq::kEv;
    if (have > (uint64_t)cap) have = (uint64_t)cap;
    for (uint64_t k = 0; k < have; ++k) {
        const int i = (int)((h->ev_count - have + k) % amps_recc_iq::kEv);
        CK(cudaEventElapsedTime(&ms_out[k], h->ev0[i], h->ev1[i]));
    }
    *n_out = (int)have;
    return AMPS_OK;
}

extern "C" int amps_recc_iq_get_taps(const amps_recc_iq *h, float *lpf_out, int cap) {
    if (!h) return set_error(AMPS_E_INVAL, "null handle");
    const int n = (int)h->lpf.size();
    if (lpf_out) for (int i = 0; i < n && i < cap; ++i) lpf_out[i] = h->lpf[i];
    return n;
}

// --------------------------------------------------------------------------------------------
// recc_decode (message-only block)
// --------------------------------------------------------------------------------------------
struct amps_recc_decode {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint8_t *d_blobs = nullptr;
    amps_recc_words *d_out = nullptr;
    int cap = 0;
};

static int decode_reserve(amps_recc_decode *h, int n) {
    if (n <= h->cap) return AMPS_OK;
    cudaFree(h->d_blobs); cudaFree(h->d_out);
    h->d_blobs = nullptr; h->d_out = nullptr; h->cap = 0;
    CK(cudaMalloc(&h->d_blobs, (size_t)n * kCapture));
    CK(cudaMalloc(&h->d_out, (size_t)n * sizeof(amps_recc_words)));
    h->cap = n;
    return AMPS_OK;
}

extern "C" int amps_recc_decode_create(int device, amps_recc_decode **out) {
    if (!out) return set_error(AMPS_E_INVAL, "null argument");
    *out = nullptr;
    int st = select_device(device);
    if (st != AMPS_OK) return st;
    amps_recc_decode *h = new (std::nothrow) amps_recc_decode();
    if (!h) return set_error(AMPS_E_NOMEM, "out of host memory");
    h->device = device;
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    st = decode_reserve(h, 16);
    if (st != AMPS_OK) { amps_recc_decode_destroy(h); return st; }
    *out = h;
    return AMPS_OK;
}

extern "C" int amps_recc_decode_destroy(amps_recc_decode *h) {
    if (!h) return AMPS_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaFree(h->d_blobs); cudaFree(h->d_out);
    delete h;
    return AMPS_OK;
}

extern "C" int amps_recc_decode_bursts(amps_recc_decode *h, const uint8_t *blobs, int nbursts, amps_recc_words *out) {
    if (!h || !blobs || !out || nbursts < 0) return set_error(AMPS_E_INVAL, "bad argument");
    if (nbursts == 0) return AMPS_OK;
    CK(cudaSetDevice(h->device));
    int st = decode_reserve(h, nbursts);
    if (st != AMPS_OK) return st;
    CK(cudaMemcpyAsync(h->d_blobs, blobs, (size_t)nbursts * kCapture, cudaMemcpyHostToDevice, h->stream));
    CKL(launch_decode_blobs(h->d_blobs, nbursts, h->d_out, h->stream));
    CK(cudaMemcpyAsync(out, h->d_out, (size_t)nbursts * sizeof(amps_recc_words), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return AMPS_OK;
}

extern "C" int amps_recc_decode_burst(amps_recc_decode *h, const uint8_t *blob3374, amps_recc_words *out) {
    return amps_recc_decode_bursts(h, blob3374, 1, out);
}
