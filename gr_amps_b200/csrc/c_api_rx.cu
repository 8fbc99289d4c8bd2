// c_api_rx.cu -- extern "C" entry points of the receive side: amps_recc_iq_* (fused IQ path, single channel and batched)
// and amps_recc_decode_* (message-only burst decoder).  See include/amps_b200.h for the contract and
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

struct amps_recc_iq_batch;

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
    uint32_t     gran = kUnit;           // input samples per processing quantum (API granularity)
    uint32_t     hist = kHist;           // input samples of history carried between calls
    uint32_t     decim = kD1 * kD2;      // input samples per demodulated sample
    uint32_t     gran_out = kUnitOut;    // demodulated samples per quantum

    RxFrontParams1   fp{};               // 10 MS/s: constant part filled at create
    RxFront400Params fp400{};            // 400 kS/s: constant part filled at create
    bool         sc16 = false;           // AMPS_RX_INPUT_SC16: interleaved int16 I/Q instead of float
    bool         sc16_unit = false;      // sc16 scale is a power of two: folded into the NCO tables, no multiply in the kernel
    size_t       isz = sizeof(float2);   // bytes per complex input sample (8 or 4)
    uint8_t     *d_stage = nullptr;      // host path: [carry | new chunk]
    uint8_t     *d_tail[2] = {nullptr, nullptr};   // history (+ device-path carry) of the next call, double-buffered
    int          tail_cur = 0;
    float       *d_dring = nullptr;
    uint32_t    *d_hring = nullptr;      // hard decisions (d >= 0), 1 bit per demod sample, same ring indexing
    uint32_t     dmask = 0;
    float2      *d_ydump = nullptr;
    size_t       ydump_cap = 0;          // complex samples
    uint64_t     ydump_first = 0, ydump_count = 0;
    RxState     *d_state = nullptr;
    Candidate   *d_cand = nullptr;
    Accepted    *d_acc = nullptr;        // bursts accepted by the select of call slot 0 .. kRxDepth-1 (kRxDepth x kMaxAccept entries)
    uint32_t    *d_flags = nullptr;      // boundary counters of the front kernel (kMaxGrid words, zero between launches)
    amps_burst  *h_ring = nullptr;       // mapped pinned host ring the capture kernel publishes into
    RxPublished *h_pub = nullptr;        // mapped pinned counters
    uint64_t     consumed = 0;           // bursts already handed to the caller
    uint64_t     lost = 0;               // bursts overwritten in the ring before they were collected
    uint32_t     overflow_seen = 0;      // candidate-list overflows already reported to the caller
    cudaStream_t last_stream = nullptr;
    cudaStream_t side = nullptr;         // search + selection (and the M&M / 400 kS/s tails) run here, overlapped with the next front kernel
    cudaStream_t side2 = nullptr;        // ... and the capture here: the search of call k+1 does not wait for the capture of call k
    cudaEvent_t  ev_sel = nullptr;       // selection of the current call finished (side -> side2)
    int          cap_par = -1;           // call slot of the last capture launched on side2 (-1: none since the last join)
    cudaEvent_t  ev_front = nullptr;     // front kernel of the current call finished
    cudaEvent_t  ev_side[kRxDepth] = {};            // side-stream work of call k finished (k mod kRxDepth)
    bool         ev_side_valid[kRxDepth] = {};      // ... and whether call k used the side stream at all
    uint64_t     call_no = 0;
    bool         serial = false;         // AMPS_RX_SERIAL=1: no overlap (profiling / A-B measurements)
    cudaStream_t copy = nullptr;         // host path, big calls: uploads run here, a piece ahead of the kernels on `stream`
    cudaEvent_t  ev_copy = nullptr;
    uint32_t     piece = 0;              // samples per uploaded piece (2^24, rounded to the granularity; AMPS_RX_PIECE overrides, 0 = off)
    bool         one_side = false;       // AMPS_RX_ONE_SIDE=1: search and capture share one side stream (A/B measurements)
    bool         front_only = false;     // AMPS_RX_FRONT_ONLY=1: no capture (pipeline measurements only: no bursts come out)
    int          grid_cap = 0;           // AMPS_RX_GRID (test hook): cap on the front kernel's grid
    bool         want_prof = false;      // AMPS_RX_PROF=1 (measurement aid): per-CTA time stamps of the last front launch
    unsigned long long *d_prof = nullptr;
    bool         nosearch = false;       // AMPS_RX_NOSEARCH=1 (measurement aid): no trigger search / selection in the front kernel
    amps_recc_iq_batch *batch = nullptr; // the handle is driven through a batch
    // AMPS_RX_TIMING_MM: the reference graph's serial tail instead of the feed-forward detector
    bool         mm_mode = false;
    bool         fused = false;          // AMPS_RX_FUSED_SEARCH: trigger search + selection inside the front kernel (2 launches per call)
    MmState     *d_mm = nullptr;
    float       *d_mmtab = nullptr;
    uint8_t     *d_sym = nullptr;
    uint32_t     sym_cap = 0;
    ReccCompatState *d_compat = nullptr;
    uint8_t     *d_blobs = nullptr;
    unsigned long long *d_blob_idx = nullptr;

    size_t       carry = 0;              // host path: unprocessed samples sitting at the front of d_stage
    uint32_t     dev_carry = 0;          // device path: unprocessed samples sitting behind the history in d_tail[tail_cur]
    uint64_t     samples_in = 0;         // samples handed to the kernels
    uint64_t     total_d = 0;            // demod samples produced
    uint64_t     groups_done = 0;        // 400 kS/s: trigger-search groups already searched
    uint64_t     bursts = 0, launches = 0;
    // AMPS_RX_TIME_KERNELS: ring of event pairs around the front-end kernel
    static constexpr int kEv = 256;
    cudaEvent_t  ev0[kEv] = {}, ev1[kEv] = {};
    uint64_t     ev_count = 0;
};

// A set of channels (handles) served by ONE front launch + ONE capture launch per call: K carriers of one GPU, or K
// independent buffers.  The handles keep their own rings and records; the batch owns the streams and the call counter.
struct amps_recc_iq_batch {
    int          device = 0;
    int          sm_count = 0;
    std::vector<amps_recc_iq *> ch;
    bool         sc16 = false, sc16_unit = false;
    size_t       isz = sizeof(float2);
    cudaStream_t stream = nullptr, side = nullptr, side2 = nullptr, last_stream = nullptr;
    cudaEvent_t  ev_front = nullptr, ev_sel = nullptr, ev_side[kRxDepth] = {};
    uint64_t     call_no = 0;
    uint64_t     launches = 0;
    uint8_t     *d_stage = nullptr;      // shared-buffer host path: [carry | new chunk], every channel reads it
    uint32_t     stage_cap = 0;          // samples
    size_t       carry = 0;
    bool         timed = false;
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

static size_t tail_samples(const amps_recc_iq *h) { return (size_t)h->hist + h->gran; }

static int rx_alloc(amps_recc_iq *h) {
    const size_t max_d = ((size_t)h->max_samples + h->gran) / h->decim + kPassOut;
    size_t cap = 1;
    // kRxDepth + 1 calls' worth: the search / capture of call k may still read its part while the front kernels of calls
    // k+1 .. k+kRxDepth-1 write theirs (small pipelined calls are then bound by the front kernel, not by the side stream)
    while (cap < (size_t)(kRxDepth + 1) * max_d + (size_t)kSpan + 4096) cap <<= 1;
    h->dmask = (uint32_t)(cap - 1);
    for (int i = 0; i < 2; ++i) {
        CK(cudaMalloc(&h->d_tail[i], tail_samples(h) * h->isz));
        CK(cudaMemset(h->d_tail[i], 0, tail_samples(h) * h->isz));
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
    CK(cudaMalloc(&h->d_cand, sizeof(Candidate) * kMaxCand * 2));      // candidates + the select's sorted copy
    CK(cudaMalloc(&h->d_acc, sizeof(Accepted) * kMaxAccept * kRxDepth));
    if (h->want_prof) { CK(cudaMalloc(&h->d_prof, sizeof(unsigned long long) * 16 * kMaxGrid)); CK(cudaMemset(h->d_prof, 0, sizeof(unsigned long long) * 16 * kMaxGrid)); }
    CK(cudaMalloc(&h->d_flags, sizeof(uint32_t) * kMaxGrid));
    CK(cudaMemset(h->d_flags, 0, sizeof(uint32_t) * kMaxGrid));
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
    CK(cudaStreamCreateWithFlags(&h->side2, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_front, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_sel, cudaEventDisableTiming));
    for (int i = 0; i < kRxDepth; ++i) CK(cudaEventCreateWithFlags(&h->ev_side[i], cudaEventDisableTiming));
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
    h->native400 = params->samp_rate == 400e3;
    if (h->native400) { h->gran = kPass400; h->hist = kPass400; h->decim = kD2; h->gran_out = kPassOut; }
    // round the per-call capacity up to whole processing quanta
    h->max_samples = (uint32_t)(((uint64_t)params->max_samples + h->gran - 1) / h->gran * h->gran);
    h->max_records = params->max_bursts ? params->max_bursts : 256;
    h->flags = params->flags;
    h->mm_mode = (params->flags & AMPS_RX_TIMING_MM) != 0;
    h->fused = (params->flags & AMPS_RX_FUSED_SEARCH) != 0;
    { const char *e = std::getenv("AMPS_RX_FUSED"); if (e && e[0] == '1') h->fused = true; }
    h->sc16 = (params->flags & AMPS_RX_INPUT_SC16) != 0;
    h->isz = h->sc16 ? sizeof(short2) : sizeof(float2);
    { const char *e = std::getenv("AMPS_RX_SERIAL"); h->serial = e && e[0] == '1'; }
    { const char *e = std::getenv("AMPS_RX_FRONT_ONLY"); h->front_only = e && e[0] == '1'; }
    { const char *e = std::getenv("AMPS_RX_ONE_SIDE"); h->one_side = e && e[0] == '1'; }
    { const char *e = std::getenv("AMPS_RX_PIECE"); const long v = e ? std::atol(e) : (1l << 24); h->piece = (uint32_t)(v > 0 ? v : 0); }
    { const char *e = std::getenv("AMPS_RX_GRID"); h->grid_cap = e ? std::atoi(e) : 0; }
    { const char *e = std::getenv("AMPS_RX_NOSEARCH"); h->nosearch = e && e[0] == '1'; }
    { const char *e = std::getenv("AMPS_RX_PROF"); h->want_prof = e && e[0] == '1'; }
    if (params->lpf_taps) h->lpf.assign(params->lpf_taps, params->lpf_taps + params->n_lpf_taps);
    else h->lpf = firdes_low_pass(3.0, 400e3, 10e3, 4500.0, WIN_BLACKMAN);     // grc/ampsbs.grc:138-184
    h->fcw = nco_fcw(params->center_freq, params->samp_rate);

    std::memset(&h->fp, 0, sizeof h->fp);
    std::memset(&h->fp400, 0, sizeof h->fp400);
    RxChan &c = h->fp.ch[0];
    c.fcw25 = (uint32_t)(25u * h->fcw);
    c.in_scale = params->sc16_scale != 0.0f ? params->sc16_scale : 1.0f / 32768.0f;
    nco_block_table(h->fcw, kD1, reinterpret_cast<float *>(c.w));
    if (h->sc16) {
        int e = 0;
        const float m = std::frexp(c.in_scale, &e);
        if (m == 0.5f && e > -40 && e < 40) {                       // power of two: (I s) w == I (s w) exactly
            h->sc16_unit = true;
            for (int k = 0; k < kD1; ++k) { c.w[k].x *= c.in_scale; c.w[k].y *= c.in_scale; }
        }
    }
    for (int k = 0; k < kD1; ++k) h->fp.wj0[k] = make_float2(-c.w[k].y, c.w[k].x);
    std::vector<float> cic;
    cic3_taps(kD1, cic);
    for (size_t i = 0; i < cic.size(); ++i) h->fp.g[i] = cic[i];
    for (size_t i = 0; i < h->lpf.size(); ++i) h->fp.h2[i] = h->lpf[i];
    h->fp.nchan = 1;
    h->fp400.fcw25 = c.fcw25;
    h->fp400.in_scale = c.in_scale;
    for (int k = 0; k < kD1; ++k) { h->fp400.w[k] = c.w[k]; h->fp400.wj[k] = h->fp.wj0[k]; }
    for (size_t i = 0; i < h->lpf.size(); ++i) h->fp400.h2[i] = h->lpf[i];

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
    if (h->side2) cudaStreamDestroy(h->side2);
    if (h->copy) cudaStreamDestroy(h->copy);
    if (h->ev_copy) cudaEventDestroy(h->ev_copy);
    if (h->ev_front) cudaEventDestroy(h->ev_front);
    if (h->ev_sel) cudaEventDestroy(h->ev_sel);
    for (int i = 0; i < kRxDepth; ++i) if (h->ev_side[i]) cudaEventDestroy(h->ev_side[i]);
    for (int i = 0; i < amps_recc_iq::kEv; ++i) { if (h->ev0[i]) cudaEventDestroy(h->ev0[i]); if (h->ev1[i]) cudaEventDestroy(h->ev1[i]); }
    cudaFree(h->d_stage); cudaFree(h->d_tail[0]); cudaFree(h->d_tail[1]); cudaFree(h->d_dring); cudaFree(h->d_hring);
    cudaFree(h->d_ydump); cudaFree(h->d_state); cudaFree(h->d_cand); cudaFree(h->d_acc); cudaFree(h->d_flags); cudaFree(h->d_prof);
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
    for (int i = 0; i < 2; ++i) CK(cudaMemset(h->d_tail[i], 0, tail_samples(h) * h->isz));
    CK(cudaMemset(h->d_state, 0, sizeof(RxState)));
    CK(cudaMemset(h->d_flags, 0, sizeof(uint32_t) * kMaxGrid));
    CK(cudaMemset(h->d_dring, 0, ((size_t)h->dmask + 1) * sizeof(float)));
    if (h->mm_mode) { int rc = mm_reset(h); if (rc != AMPS_OK) return rc; }
    std::memset(h->h_pub, 0, sizeof(RxPublished));
    h->consumed = 0; h->call_no = 0; h->overflow_seen = 0;
    for (int i = 0; i < kRxDepth; ++i) h->ev_side_valid[i] = false;
    h->cap_par = -1;
    h->tail_cur = 0; h->carry = 0; h->dev_carry = 0; h->samples_in = 0; h->total_d = 0; h->groups_done = 0;
    h->ydump_first = 0; h->ydump_count = 0;
    return AMPS_OK;
}

extern "C" int amps_recc_iq_granularity(const amps_recc_iq *h) { return h ? (int)h->gran : kUnit; }

// ---- per-channel pieces of a 10 MS/s call -----------------------------------------------------
// Fills the launch arguments of one channel for `units` whole units out of [tail carry | chunk] and advances the handle's
// host-side stream position.  par = call parity of whoever owns the call counter (the handle or its batch).
static void chan_begin(amps_recc_iq *h, RxChan &c, const uint8_t *d_chunk, uint32_t nchunk, uint32_t units, uint32_t par) {
    c = h->fp.ch[0];                                   // constant part (NCO tables, scale)
    c.chunk = d_chunk;
    c.tail = h->d_tail[h->tail_cur];
    c.tail_out = h->d_tail[h->tail_cur ^ 1];
    c.dring = h->d_dring;
    c.hring = h->d_hring;
    c.ydump = h->d_ydump;
    c.state = h->d_state;
    c.cand = h->d_cand;
    c.acc = h->d_acc;
    c.host_pub = h->h_pub;
    c.flags = h->d_flags;
    c.q_base = h->total_d;
    c.dmask = h->dmask;
    c.units = units;
    c.carry = h->dev_carry;
    c.nchunk = nchunk;
    c.blk_base = (uint32_t)(h->samples_in / kD1);
    c.par = par;
    c.search = (h->fused && !h->mm_mode && !h->nosearch) ? 1u : 0u;
    h->tail_cur ^= 1;
    h->dev_carry = h->dev_carry + nchunk - units * (uint32_t)kUnit;
    h->ydump_first = h->total_d;
    h->ydump_count = (uint64_t)units * kUnitOut;
    h->samples_in += (uint64_t)units * kUnit;
    h->total_d += (uint64_t)units * kUnitOut;
}

static void chan_capture(const amps_recc_iq *h, RxCaptureChan &cc, uint32_t par, uint32_t cta_first, uint32_t cta_count) {
    cc.dring = h->d_dring;
    cc.state = h->d_state;
    cc.acc = h->d_acc + (size_t)par * kMaxAccept;
    cc.host_ring = h->h_ring;
    cc.host_pub = h->h_pub;
    cc.blobs = h->mm_mode ? h->d_blobs : nullptr;
    cc.blob_sym_index = h->mm_mode ? h->d_blob_idx : nullptr;
    cc.dmask = h->dmask;
    cc.ring_len = h->max_records;
    cc.decim = h->decim;
    cc.par = par;
    cc.cta_first = cta_first;
    cc.cta_count = cta_count;
}

// the stand-alone search of a channel: every group whose lookahead is complete and that was not searched yet
static uint32_t chan_search(amps_recc_iq *h, RxSearchChan &sc, uint32_t par, uint32_t cta_first) {
    const uint64_t g_hi = h->total_d / kUnitOut >= (uint64_t)kGroupLag ? h->total_d / kUnitOut - kGroupLag : 0;
    sc.dring = h->d_dring; sc.hring = h->d_hring; sc.state = h->d_state; sc.cand = h->d_cand; sc.acc = h->d_acc; sc.host_pub = h->h_pub;
    sc.g_lo = h->groups_done;
    sc.g_hi = g_hi > h->groups_done ? g_hi : h->groups_done;
    sc.total_d = h->total_d;
    sc.dmask = h->dmask; sc.par = par; sc.cta_first = cta_first;
    sc.cta_count = rx_search_ctas(sc.g_hi - sc.g_lo);
    sc.host_ring = h->h_ring; sc.ring_len = h->max_records; sc.decim = h->decim;
    h->groups_done = sc.g_hi;
    return sc.cta_count;
}

// the M&M tail of a channel
static void chan_mm(const amps_recc_iq *h, RxMmChan &m, uint32_t par) {
    m.dring = h->d_dring; m.mm = h->d_mm; m.sym = h->d_sym; m.cs = h->d_compat; m.blobs = h->d_blobs; m.blob_sym_index = h->d_blob_idx;
    m.state = h->d_state; m.host_pub = h->h_pub; m.total_d = h->total_d; m.dmask = h->dmask; m.sym_cap = h->sym_cap; m.par = par; m.pad = 0;
}
static uint32_t mm_capture_ctas(uint64_t outputs) {                      // <= one blob per emulated work() quantum
    const uint64_t mx = outputs / (8u * (unsigned)kMmQuantum) + 2;
    return (uint32_t)(mx > (uint64_t)kMaxAccept ? (uint64_t)kMaxAccept : mx);
}

// bursts one call of `outputs` demodulated samples can make capturable: they are at least kBurstLen apart, +1 for a run
// that became decidable at the edge, +1 for leftovers of a call that had more than the list holds
static uint32_t capture_ctas(uint64_t outputs) {
    uint64_t n = outputs / (uint64_t)kBurstLen + 2;
    return (uint32_t)(n > (uint64_t)kMaxAccept ? (uint64_t)kMaxAccept : n);
}

// fewer than one unit in [carry | chunk]: nothing to launch, the samples join the carry behind the history
static int chan_append_carry(amps_recc_iq *h, const uint8_t *d_chunk, uint32_t nchunk, cudaStream_t st) {
    if (nchunk)
        CK(cudaMemcpyAsync(h->d_tail[h->tail_cur] + ((size_t)h->hist + h->dev_carry) * h->isz, d_chunk, (size_t)nchunk * h->isz,
                           cudaMemcpyDeviceToDevice, st));
    h->dev_carry += nchunk;
    return AMPS_OK;
}

// Enqueue everything for one call of a single 10 MS/s channel: [tail carry | nchunk samples at d_chunk].
// in_order: the caller synchronises right after this call (the host-buffer path), so nothing is gained by moving the search /
// capture to the side stream -- they follow the front kernel on `st` and four driver calls (two event records, two waits) go.
static int rx_enqueue10(amps_recc_iq *h, const uint8_t *d_chunk, uint32_t nchunk, cudaStream_t st, bool in_order = false) {
    AMPS_NVTX("amps_recc_iq: enqueue (front + search + capture)");
    const uint32_t units = (h->dev_carry + nchunk) / (uint32_t)kUnit;
    h->last_stream = st;
    if (units == 0) return chan_append_carry(h, d_chunk, nchunk, st);
    const uint32_t par = (uint32_t)(h->call_no % kRxDepth);
    // the demod ring holds kRxDepth + 1 calls, the accepted-burst lists kRxDepth: do not overwrite what the side-stream work of
    // call k - kRxDepth may still read
    if (h->ev_side_valid[par]) CK(cudaStreamWaitEvent(st, h->ev_side[par], 0));
    RxFrontParams1 p = h->fp;
    p.prof = h->d_prof;
    chan_begin(h, p.ch[0], d_chunk, nchunk, units, par);
    const uint32_t tiles = rx_tiles_of(units);
    p.tile_cum[0] = 0; p.tile_cum[1] = tiles;
    uint32_t resident = (uint32_t)rx_front_ctas_per_sm(h->sc16) * (uint32_t)h->sm_count;
    // between one and five tiles per SM a CTA that has its SM to itself gets through its two warm-up tiles and its 2 .. 5 own
    // ones sooner than two CTAs that share the SM get through two warm-up tiles and 1 .. 2 own ones each (measured,
    // tools/grid_probe.py: 1.5 M .. 3.5 M samples 14.5 .. 16.5 us against 15.2 .. 18.0 us; equal at 4.2 M, worse beyond)
    if (!h->sc16 && h->grid_cap == 0 && tiles > resident && tiles <= 5u * (uint32_t)h->sm_count) resident = (uint32_t)h->sm_count;
    if (h->grid_cap > 0 && resident > (uint32_t)h->grid_cap) resident = (uint32_t)h->grid_cap;
    const uint32_t grid = rx_make_deal(p.deal, tiles, resident);
    const bool timed = (h->flags & AMPS_RX_TIME_KERNELS) != 0;
    const int  evi = (int)(h->ev_count % amps_recc_iq::kEv);
    if (timed) CK(cudaEventRecord(h->ev0[evi], st));
    CKL(launch_rx_front(p, (int)grid, st, h->sc16, h->sc16_unit, h->fused && !h->mm_mode && !h->nosearch));
    if (timed) { CK(cudaEventRecord(h->ev1[evi], st)); h->ev_count++; }
    h->launches++;
    // capture (and the M&M tail) on the side stream, so that the next call's front kernel (HBM-bound, 2 CTAs/SM) overlaps
    // these small latency-bound kernels
    const bool serial = h->serial || in_order;
    cudaStream_t sd = serial ? st : h->side;
    if (!serial) {
        CK(cudaEventRecord(h->ev_front, st));
        CK(cudaStreamWaitEvent(sd, h->ev_front, 0));
    }
    if (!h->front_only) {
        RxCaptureParams cp;
        cp.nchan = 1;
        uint32_t nc = capture_ctas((uint64_t)units * kUnitOut);
        if (h->mm_mode) {
            // serial tail: M&M + slicer over everything demodulated so far, amps.recc on the new half-symbols, then
            // one CTA per blob decodes and publishes it
            static thread_local RxMmParams mp;
            mp.nchan = 1; mp.max_blobs = kMaxAccept; mp.table = h->d_mmtab;
            chan_mm(h, mp.ch[0], par);
            CKL(launch_rx_mm(mp, sd));
            nc = mm_capture_ctas((uint64_t)units * kUnitOut);
            h->launches += 2;
        }
        // a call that can make at most two bursts capturable: the search kernel's last CTA captures them itself (2 launches)
        const bool small = !h->mm_mode && !h->fused && !h->nosearch && nc <= 2u;
        if (!h->mm_mode && !h->fused && !h->nosearch) {
            // trigger search + selection as their own launch, overlapped (like the capture) with the next call's front kernel
            RxSearchParams sp;
            sp.nchan = 1;
            sp.prof = h->d_prof;
            const uint32_t ns = chan_search(h, sp.ch[0], par, 0);
            // (a search that captures by itself publishes the record count: the captures still running on side2 go first)
            if (small && !serial && h->cap_par >= 0) { CK(cudaStreamWaitEvent(sd, h->ev_side[h->cap_par], 0)); h->cap_par = -1; }
            CKL(launch_rx_search(sp, (int)ns, small, sd));
            h->launches++;
            if (!small && !serial && !h->one_side) {
                // the capture gets a stream of its own: search + selection of consecutive calls are one chain (the candidate
                // list), the captures another -- a pipelined call costs the longer of the two, not their sum
                CK(cudaEventRecord(h->ev_sel, sd));
                sd = h->side2;
                CK(cudaStreamWaitEvent(sd, h->ev_sel, 0));
                h->cap_par = (int)par;
            }
        }
        if (!small) {
            chan_capture(h, cp.ch[0], par, 0, nc);
            CKL(launch_rx_capture(cp, (int)nc, sd));
            h->launches++;
        }
    }
    h->ev_side_valid[par] = !serial;
    if (!serial) CK(cudaEventRecord(h->ev_side[par], sd));
    h->call_no++;
    return AMPS_OK;
}

// Enqueue everything for `npass` whole passes of the 400 kS/s front end whose samples start at d_chunk.
static int rx_enqueue400(amps_recc_iq *h, const uint8_t *d_chunk, uint32_t npass, cudaStream_t st) {
    AMPS_NVTX("amps_recc_iq: enqueue 400 kS/s");
    RxFront400Params p = h->fp400;
    p.chunk = d_chunk;
    p.tail = h->d_tail[h->tail_cur];
    p.dring = h->d_dring;
    p.hring = h->d_hring;
    p.dmask = h->dmask;
    p.q_base = h->total_d;
    p.npass = npass;
    p.n_base = h->samples_in;
    p.ydump = h->d_ydump;
    const uint32_t resident = 4u * (uint32_t)h->sm_count;
    p.pass_per_cta = (npass + resident - 1u) / resident;
    const int grid = (int)((npass + p.pass_per_cta - 1u) / p.pass_per_cta);
    const uint32_t par = (uint32_t)(h->call_no % kRxDepth);
    if (h->ev_side_valid[par]) CK(cudaStreamWaitEvent(st, h->ev_side[par], 0));
    const bool timed = (h->flags & AMPS_RX_TIME_KERNELS) != 0;
    const int  evi = (int)(h->ev_count % amps_recc_iq::kEv);
    if (timed) CK(cudaEventRecord(h->ev0[evi], st));
    CKL(launch_rx_front400(p, grid, st, h->sc16, h->sc16_unit));
    if (timed) { CK(cudaEventRecord(h->ev1[evi], st)); h->ev_count++; }
    h->launches++;
    // history for the next call = the tail of this one
    CK(cudaMemcpyAsync(h->d_tail[h->tail_cur ^ 1], d_chunk + ((size_t)npass * h->gran - h->hist) * h->isz, (size_t)h->hist * h->isz,
                       cudaMemcpyDeviceToDevice, st));
    h->tail_cur ^= 1;
    h->ydump_first = h->total_d;
    h->ydump_count = (uint64_t)npass * kPassOut;
    h->samples_in += (uint64_t)npass * h->gran;
    h->total_d += (uint64_t)npass * kPassOut;
    CK(cudaEventRecord(h->ev_front, st));
    cudaStream_t sd = h->serial ? st : h->side;
    if (!h->serial) CK(cudaStreamWaitEvent(sd, h->ev_front, 0));
    if (!h->front_only) {
        RxCaptureParams cp;
        cp.nchan = 1;
        uint32_t nc = capture_ctas((uint64_t)npass * kPassOut);
        if (h->mm_mode) {
            static thread_local RxMmParams mp;
            mp.nchan = 1; mp.max_blobs = kMaxAccept; mp.table = h->d_mmtab;
            chan_mm(h, mp.ch[0], par);
            CKL(launch_rx_mm(mp, sd));
            nc = mm_capture_ctas((uint64_t)npass * kPassOut);
            h->launches += 2;
        } else {
            // search every group whose lookahead is complete, then select (one launch)
            RxSearchParams sp;
            sp.nchan = 1;
            const uint32_t ns = chan_search(h, sp.ch[0], par, 0);
            CKL(launch_rx_search(sp, (int)ns, false, sd));
            h->launches++;
        }
        chan_capture(h, cp.ch[0], par, 0, nc);
        CKL(launch_rx_capture(cp, (int)nc, sd));
        h->launches++;
    }
    CK(cudaEventRecord(h->ev_side[par], sd));
    h->ev_side_valid[par] = true;
    h->call_no++;
    h->last_stream = st;
    return AMPS_OK;
}

static int rx_submit_dev(amps_recc_iq *h, const void *d_iq, size_t nsamples, void *cuda_stream, bool sc16) {
    if (!h || (!d_iq && nsamples)) return set_error(AMPS_E_INVAL, "null argument");
    if (h->batch) return set_error(AMPS_E_STATE, "the handle belongs to a batch: use amps_recc_iq_batch_*");
    if (h->sc16 != sc16) return set_error(AMPS_E_STATE, sc16 ? "handle was not created with AMPS_RX_INPUT_SC16" : "handle was created with AMPS_RX_INPUT_SC16: use the _sc16 entry points");
    if (nsamples == 0) return AMPS_OK;
    if (reinterpret_cast<uintptr_t>(d_iq) & 15u) return set_error(AMPS_E_ALIGN, "d_iq must be 16-byte aligned");
    if (nsamples > h->max_samples) return set_error(AMPS_E_OVERFLOW, "nsamples exceeds max_samples");
    if (h->carry) return set_error(AMPS_E_STATE, "host-path samples are pending; reset() or keep using work()");
    CK(cudaSetDevice(h->device));
    if (h->native400) {
        if (nsamples % h->gran) return set_error(AMPS_E_ALIGN, "nsamples must be a multiple of amps_recc_iq_granularity() at 400 kS/s");
        return rx_enqueue400(h, static_cast<const uint8_t *>(d_iq), (uint32_t)(nsamples / h->gran), static_cast<cudaStream_t>(cuda_stream));
    }
    // any length whose byte count is a multiple of 16 (the bulk-copy engine's granule): what does not fill a 1600-sample unit
    // is carried on the device to the next call
    if ((nsamples * h->isz) & 15u) return set_error(AMPS_E_ALIGN, sc16 ? "nsamples must be a multiple of 4" : "nsamples must be a multiple of 2");
    return rx_enqueue10(h, static_cast<const uint8_t *>(d_iq), (uint32_t)nsamples, static_cast<cudaStream_t>(cuda_stream));
}
extern "C" int amps_recc_iq_submit_dev(amps_recc_iq *h, const void *d_iq, size_t nsamples, void *cuda_stream) {
    return rx_submit_dev(h, d_iq, nsamples, cuda_stream, false);
}
extern "C" int amps_recc_iq_submit_sc16_dev(amps_recc_iq *h, const void *d_iq, size_t nsamples, void *cuda_stream) {
    return rx_submit_dev(h, d_iq, nsamples, cuda_stream, true);
}

// Wait for the stream; afterwards records [h->consumed, h->consumed + *n_out) sit in the host ring.
// *overflowed is set ONCE per overflow of the candidate list (the bursts that were captured are still delivered).
static int rx_fetch(amps_recc_iq *h, uint64_t *n_out, bool *overflowed) {
    AMPS_NVTX("amps_recc_iq: fetch bursts");
    *n_out = 0;
    *overflowed = false;
    if (h->batch) {
        CK(cudaStreamSynchronize(h->batch->side));
        CK(cudaStreamSynchronize(h->batch->side2));
        CK(cudaStreamSynchronize(h->batch->last_stream));       // (a null handle is the default stream)
    } else {
        CK(cudaStreamSynchronize(h->side));
        CK(cudaStreamSynchronize(h->side2));
        CK(cudaStreamSynchronize(h->last_stream));
    }
    const uint32_t ov = h->h_pub->cand_overflow;
    if (ov != h->overflow_seen) { h->overflow_seen = ov; *overflowed = true; }
    const uint64_t total = h->h_pub->nrec_total;
    if (total - h->consumed > h->max_records) {           // the ring wrapped over uncollected records
        h->lost += total - h->consumed - h->max_records;
        h->consumed = total - h->max_records;
    }
    *n_out = total - h->consumed;
    return AMPS_OK;
}
static int overflow_status(const amps_recc_iq *h) {
    return set_error(AMPS_E_OVERFLOW, h->mm_mode ? "more than 512 bursts captured in one call: the surplus was dropped (reported once; the stream goes on)"
                                                 : "trigger candidate list overflowed (more than 8192 undecided matches): candidates were dropped (reported once; the stream goes on)");
}

extern "C" int amps_recc_iq_collect(amps_recc_iq *h, amps_burst *out, int max, int *n_out) {
    if (!h || !n_out || (max > 0 && !out)) return set_error(AMPS_E_INVAL, "null argument");
    *n_out = 0;
    CK(cudaSetDevice(h->device));
    uint64_t n = 0;
    bool ov = false;
    int rc = rx_fetch(h, &n, &ov);
    if (rc != AMPS_OK) return rc;
    const uint64_t give = n < (uint64_t)max ? n : (uint64_t)max;     // the rest stays for the next collect
    for (uint64_t i = 0; i < give; ++i) out[i] = h->h_ring[(h->consumed + i) % h->max_records];
    h->consumed += give;
    h->bursts += give;
    *n_out = (int)give;
    return ov ? overflow_status(h) : AMPS_OK;
}

extern "C" int amps_recc_iq_peek(amps_recc_iq *h, const amps_burst **ring, uint32_t *ring_len, uint64_t *first, uint64_t *count) {
    if (!h || !ring || !ring_len || !first || !count) return set_error(AMPS_E_INVAL, "null argument");
    CK(cudaSetDevice(h->device));
    uint64_t n = 0;
    bool ov = false;
    int rc = rx_fetch(h, &n, &ov);
    if (rc != AMPS_OK) return rc;
    *ring = h->h_ring; *ring_len = h->max_records; *first = h->consumed; *count = n;
    return ov ? overflow_status(h) : AMPS_OK;
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
    AMPS_NVTX("amps_recc_iq_work");
    if (!h || (!iq_host && nsamples)) return set_error(AMPS_E_INVAL, "null argument");
    if (h->batch) return set_error(AMPS_E_STATE, "the handle belongs to a batch: use amps_recc_iq_batch_*");
    if (h->sc16 != sc16) return set_error(AMPS_E_STATE, sc16 ? "handle was not created with AMPS_RX_INPUT_SC16" : "handle was created with AMPS_RX_INPUT_SC16: use the _sc16 entry points");
    if (nsamples > h->max_samples) return set_error(AMPS_E_OVERFLOW, "nsamples exceeds max_samples");
    if (h->dev_carry) return set_error(AMPS_E_STATE, "device-path samples are pending; reset() or keep using submit_dev()");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    // the staging buffer of the host path exists from the first host call on (device-resident and batched use never needs it)
    if (!h->d_stage) CK(cudaMalloc(&h->d_stage, ((size_t)h->max_samples + h->gran) * h->isz));
    // a caller that used submit_dev() before and did not collect: this call's kernels run in order on `st`, behind whatever
    // search / capture work of those calls is still on the side streams
    for (int i = 0; i < kRxDepth; ++i)
        if (h->ev_side_valid[i]) { CK(cudaStreamWaitEvent(st, h->ev_side[i], 0)); h->ev_side_valid[i] = false; }
    h->cap_par = -1;
    h->last_stream = st;
    const size_t piece = (size_t)(h->piece / h->gran) * h->gran;
    if (!h->native400 && !h->d_ydump && piece && nsamples >= 2 * piece) {
        // a big call goes up in pieces on a stream of its own and the kernels of piece i run while piece i+1 is on the link:
        // what the caller waits for after the last byte has arrived is one piece's kernels, not the whole call's
        const uint8_t *src = static_cast<const uint8_t *>(iq_host);
        size_t up = 0, enq = 0;                          // samples uploaded / samples handed to the kernels (staging coordinates)
        while (up < nsamples) {
            size_t n = nsamples - up;
            if (n >= 2 * piece) n = piece;               // (the last piece is between one and two pieces long)
            CK(cudaMemcpyAsync(h->d_stage + (h->carry + up) * h->isz, src + up * h->isz, n * h->isz, cudaMemcpyHostToDevice, h->copy));
            CK(cudaEventRecord(h->ev_copy, h->copy));
            CK(cudaStreamWaitEvent(st, h->ev_copy, 0));
            up += n;
            const size_t nq = (h->carry + up - enq) / h->gran;
            if (nq) {
                int rc = rx_enqueue10(h, h->d_stage + enq * h->isz, (uint32_t)(nq * h->gran), st, /*in_order=*/true);
                if (rc != AMPS_OK) return rc;
                enq += nq * h->gran;
            }
        }
        const size_t left = h->carry + nsamples - enq;
        if (left) CK(cudaMemcpyAsync(h->d_stage, h->d_stage + enq * h->isz, left * h->isz, cudaMemcpyDeviceToDevice, st));
        h->carry = left;
        uint64_t nb = 0;
        bool ovf = false;
        int rc = rx_fetch(h, &nb, &ovf);
        if (rc != AMPS_OK) return rc;
        if (cb) for (uint64_t i = 0; i < nb; ++i) cb(&h->h_ring[(h->consumed + i) % h->max_records], user);
        h->consumed += nb;
        h->bursts += nb;
        return ovf ? overflow_status(h) : AMPS_OK;
    }
    if (nsamples)
        CK(cudaMemcpyAsync(h->d_stage + h->carry * h->isz, iq_host, nsamples * h->isz, cudaMemcpyHostToDevice, st));
    const size_t avail = h->carry + nsamples;
    const uint32_t nq = (uint32_t)(avail / h->gran);       // whole processing quanta (units at 10 MS/s, passes at 400 kS/s)
    if (nq) {
        int rc = h->native400 ? rx_enqueue400(h, h->d_stage, nq, st) : rx_enqueue10(h, h->d_stage, nq * h->gran, st, /*in_order=*/true);
        if (rc != AMPS_OK) return rc;
        const size_t left = avail - (size_t)nq * h->gran;
        if (left)
            CK(cudaMemcpyAsync(h->d_stage, h->d_stage + (size_t)nq * h->gran * h->isz, left * h->isz, cudaMemcpyDeviceToDevice, st));
        h->carry = left;
    } else {
        h->carry = avail;
    }
    // deliver bursts in stream order, like message_port_pub("bursts", ...) from work() (lib/recc_impl.cc:126);
    // the callback sees the record in place in the pinned ring
    uint64_t n = 0;
    bool ov = false;
    int rc = rx_fetch(h, &n, &ov);
    if (rc != AMPS_OK) return rc;
    if (cb) for (uint64_t i = 0; i < n; ++i) cb(&h->h_ring[(h->consumed + i) % h->max_records], user);
    h->consumed += n;
    h->bursts += n;
    return ov ? overflow_status(h) : AMPS_OK;
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

static int event_times(cudaEvent_t *ev0, cudaEvent_t *ev1, uint64_t ev_count, int ring, float *ms_out, int cap, int *n_out) {
    uint64_t have = ev_count < (uint64_t)ring ? ev_count : (uint64_t)ring;
    if (have > (uint64_t)cap) have = (uint64_t)cap;
    for (uint64_t k = 0; k < have; ++k) {
        const int i = (int)((ev_count - have + k) % ring);
        CK(cudaEventElapsedTime(&ms_out[k], ev0[i], ev1[i]));
    }
    *n_out = (int)have;
    return AMPS_OK;
}

extern "C" int amps_recc_iq_front_times(amps_recc_iq *h, float *ms_out, int cap, int *n_out) {
    if (!h || !n_out || (cap > 0 && !ms_out)) return set_error(AMPS_E_INVAL, "null argument");
    *n_out = 0;
    if (!(h->flags & AMPS_RX_TIME_KERNELS)) return set_error(AMPS_E_STATE, "handle was not created with AMPS_RX_TIME_KERNELS");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->last_stream));
    return event_times(h->ev0, h->ev1, h->ev_count, amps_recc_iq::kEv, ms_out, cap, n_out);
}

// test aid (pure host arithmetic, no device needed): how `tiles` tiles of `nchan` channels would be dealt to CTAs
extern "C" int amps_b200_debug_deal(uint32_t tiles, uint32_t resident, uint32_t nchan, uint32_t equal_tiles, uint32_t *grid_out,
                                    uint32_t *lo_out, uint32_t *owner_out) {
    if (!grid_out || !lo_out || !owner_out || tiles == 0) return set_error(AMPS_E_INVAL, "null argument or no tiles");
    RxDeal d;
    const uint32_t grid = rx_make_deal(d, tiles, resident, nchan, equal_tiles);
    *grid_out = grid;
    for (uint32_t s = 0; s <= grid; ++s) lo_out[s] = deal_lo(d, s);
    for (uint32_t t = 0; t < tiles; ++t) owner_out[t] = deal_owner(d, t);
    return AMPS_OK;
}

extern "C" int amps_recc_iq_debug_prof(amps_recc_iq *h, unsigned long long *out, int ctas) {
    if (!h || !out || ctas < 0 || ctas > kMaxGrid) return set_error(AMPS_E_INVAL, "bad argument");
    if (!h->d_prof) return set_error(AMPS_E_STATE, "create the handle with AMPS_RX_PROF=1 in the environment");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, h->d_prof, sizeof(unsigned long long) * 16 * (size_t)ctas, cudaMemcpyDeviceToHost));
    return AMPS_OK;
}

extern "C" int amps_recc_iq_get_taps(const amps_recc_iq *h, float *lpf_out, int cap) {
    if (!h) return set_error(AMPS_E_INVAL, "null handle");
    const int n = (int)h->lpf.size();
    if (lpf_out) for (int i = 0; i < n && i < cap; ++i) lpf_out[i] = h->lpf[i];
    return n;
}

// --------------------------------------------------------------------------------------------
// batched calls: K channels, one front launch + one capture launch (per kMaxBatch channels)
// --------------------------------------------------------------------------------------------
extern "C" int amps_recc_iq_batch_create(amps_recc_iq *const *handles, int count, uint32_t flags, amps_recc_iq_batch **out) {
    if (!handles || !out || count < 1) return set_error(AMPS_E_INVAL, "bad argument");
    *out = nullptr;
    const amps_recc_iq *h0 = handles[0];
    for (int i = 0; i < count; ++i) {
        const amps_recc_iq *h = handles[i];
        if (!h) return set_error(AMPS_E_INVAL, "null handle in the batch");
        if (h->batch) return set_error(AMPS_E_STATE, "a handle already belongs to a batch");
        if (h->native400) return set_error(AMPS_E_INVAL, "batches take 10 MS/s handles only");
        if (h->device != h0->device || h->sc16 != h0->sc16 || h->sc16_unit != h0->sc16_unit || h->fused != h0->fused || h->mm_mode != h0->mm_mode)
            return set_error(AMPS_E_INVAL, "all handles of a batch must share the device, the input format, the timing mode and AMPS_RX_FUSED_SEARCH");
        if (h->call_no || h->carry || h->dev_carry) return set_error(AMPS_E_STATE, "handles must be fresh (or reset) when they join a batch");
        for (int j = 0; j < i; ++j) if (handles[j] == h) return set_error(AMPS_E_INVAL, "the same handle twice in a batch");
    }
    CK(cudaSetDevice(h0->device));
    amps_recc_iq_batch *b = new (std::nothrow) amps_recc_iq_batch();
    if (!b) return set_error(AMPS_E_NOMEM, "out of host memory");
    b->device = h0->device; b->sm_count = h0->sm_count; b->sc16 = h0->sc16; b->sc16_unit = h0->sc16_unit; b->isz = h0->isz;
    b->ch.assign(handles, handles + count);
    b->timed = (flags & AMPS_RX_TIME_KERNELS) != 0;
    cudaError_t e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->side2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_front, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_sel, cudaEventDisableTiming);
    for (int i = 0; i < kRxDepth && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&b->ev_side[i], cudaEventDisableTiming);
    if (b->timed)
        for (int i = 0; i < amps_recc_iq_batch::kEv && e == cudaSuccess; ++i) { e = cudaEventCreate(&b->ev0[i]); if (e == cudaSuccess) e = cudaEventCreate(&b->ev1[i]); }
    if (e != cudaSuccess) { amps_recc_iq_batch_destroy(b); return set_cuda_error(e, "batch streams / events"); }
    for (amps_recc_iq *h : b->ch) h->batch = b;
    *out = b;
    return AMPS_OK;
}

extern "C" int amps_recc_iq_batch_destroy(amps_recc_iq_batch *b) {
    if (!b) return AMPS_OK;
    cudaSetDevice(b->device);
    cudaDeviceSynchronize();
    for (amps_recc_iq *h : b->ch) if (h->batch == b) h->batch = nullptr;
    if (b->stream) cudaStreamDestroy(b->stream);
    if (b->side) cudaStreamDestroy(b->side);
    if (b->side2) cudaStreamDestroy(b->side2);
    if (b->ev_front) cudaEventDestroy(b->ev_front);
    if (b->ev_sel) cudaEventDestroy(b->ev_sel);
    for (int i = 0; i < kRxDepth; ++i) if (b->ev_side[i]) cudaEventDestroy(b->ev_side[i]);
    for (int i = 0; i < amps_recc_iq_batch::kEv; ++i) { if (b->ev0[i]) cudaEventDestroy(b->ev0[i]); if (b->ev1[i]) cudaEventDestroy(b->ev1[i]); }
    cudaFree(b->d_stage);
    delete b;
    return AMPS_OK;
}

extern "C" int amps_recc_iq_batch_size(const amps_recc_iq_batch *b) { return b ? (int)b->ch.size() : 0; }

// one call: channel i gets nsamples[i] new samples at d_iq[i] (device pointers; several channels may share one buffer)
static int batch_enqueue(amps_recc_iq_batch *b, const void *const *d_iq, const size_t *nsamples, cudaStream_t st) {
    AMPS_NVTX("amps_recc_iq_batch: enqueue");
    const uint32_t par = (uint32_t)(b->call_no % kRxDepth);
    b->last_stream = st;
    if (b->call_no >= (uint64_t)kRxDepth) CK(cudaStreamWaitEvent(st, b->ev_side[par], 0));
    const uint32_t resident = (uint32_t)rx_front_ctas_per_sm(b->sc16) * (uint32_t)b->sm_count;
    const int evi = (int)(b->ev_count % amps_recc_iq_batch::kEv);
    bool ev0_done = !b->timed;
    cudaStream_t sd = b->side;
    static thread_local RxFrontParamsB p;                  // 24 KB: not on the stack
    size_t i = 0;
    const size_t K = b->ch.size();
    std::vector<RxCaptureParams> caps;
    std::vector<uint32_t> cap_grid;
    std::vector<RxSearchParams> srch;
    std::vector<uint32_t> srch_grid;
    const bool mm = b->ch[0]->mm_mode;
    const bool split = !mm && !b->ch[0]->fused && !b->ch[0]->nosearch;
    std::vector<RxMmParams> mms;
    while (i < K) {
        // next group of up to kMaxBatch channels that have at least one whole unit
        const amps_recc_iq *h0 = b->ch[0];
        std::memcpy(p.g, h0->fp.g, sizeof p.g);
        std::memcpy(p.h2, h0->fp.h2, sizeof p.h2);
        p.prof = nullptr;
        uint32_t n = 0, tiles = 0, cap_ctas = 0, srch_ctas = 0, eq_tiles = 0;
        bool all_equal = true;
        RxCaptureParams cp;
        RxSearchParams sp;
        RxMmParams mp;
        mp.max_blobs = kMaxAccept; mp.table = b->ch[0]->d_mmtab;
        p.tile_cum[0] = 0;
        for (; i < K && n < (uint32_t)kMaxBatch; ++i) {
            amps_recc_iq *h = b->ch[i];
            const uint32_t nchunk = (uint32_t)nsamples[i];
            const uint32_t units = (h->dev_carry + nchunk) / (uint32_t)kUnit;
            if (units == 0) { int rc = chan_append_carry(h, static_cast<const uint8_t *>(d_iq[i]), nchunk, st); if (rc != AMPS_OK) return rc; continue; }
            chan_begin(h, p.ch[n], static_cast<const uint8_t *>(d_iq[i]), nchunk, units, par);
            const uint32_t ct = rx_tiles_of(units);
            if (n == 0) eq_tiles = ct; else if (ct != eq_tiles) all_equal = false;
            tiles += ct;
            p.tile_cum[n + 1] = tiles;
            const uint32_t nc = mm ? mm_capture_ctas((uint64_t)units * kUnitOut) : capture_ctas((uint64_t)units * kUnitOut);
            if (mm) chan_mm(h, mp.ch[n], par);
            chan_capture(h, cp.ch[n], par, cap_ctas, nc);
            cap_ctas += nc;
            if (split) srch_ctas += chan_search(h, sp.ch[n], par, srch_ctas);
            ++n;
        }
        if (n == 0) continue;
        p.nchan = n;
        cp.nchan = n;
        const uint32_t grid = rx_make_deal(p.deal, tiles, resident, n, all_equal ? eq_tiles : 0u);
        if (!ev0_done) { CK(cudaEventRecord(b->ev0[evi], st)); ev0_done = true; }      // (after the host-side set-up)
        CKL(launch_rx_front_batch(p, (int)grid, st, b->sc16, b->sc16_unit, !split && !mm));
        b->launches++;
        caps.push_back(cp);
        cap_grid.push_back(cap_ctas);
        if (split) { sp.nchan = n; srch.push_back(sp); srch_grid.push_back(srch_ctas); }
        if (mm) { mp.nchan = n; mms.push_back(mp); }
    }
    if (b->timed && !ev0_done) CK(cudaEventRecord(b->ev0[evi], st));                 // (no channel had a whole unit)
    if (b->timed) { CK(cudaEventRecord(b->ev1[evi], st)); b->ev_count++; }
    CK(cudaEventRecord(b->ev_front, st));
    CK(cudaStreamWaitEvent(sd, b->ev_front, 0));
    if (b->ch[0]->front_only) caps.clear();               // (AMPS_RX_FRONT_ONLY: a measurement switch, see tools/roofline_sweep.py)
    for (size_t k = 0; k < caps.size(); ++k) {
        if (split) { CKL(launch_rx_search(srch[k], (int)srch_grid[k], false, sd)); b->launches++; }
        if (mm) { CKL(launch_rx_mm(mms[k], sd)); b->launches += 2; }
        if (!split) { CKL(launch_rx_capture(caps[k], (int)cap_grid[k], sd)); b->launches++; }
    }
    if (split && !caps.empty() && b->ch[0]->one_side) {
        for (size_t k = 0; k < caps.size(); ++k) { CKL(launch_rx_capture(caps[k], (int)cap_grid[k], sd)); b->launches++; }
    } else if (split && !caps.empty()) {
        // the captures on a stream of their own, behind every selection of this call (see rx_enqueue10)
        CK(cudaEventRecord(b->ev_sel, sd));
        sd = b->side2;
        CK(cudaStreamWaitEvent(sd, b->ev_sel, 0));
        for (size_t k = 0; k < caps.size(); ++k) { CKL(launch_rx_capture(caps[k], (int)cap_grid[k], sd)); b->launches++; }
    }
    CK(cudaEventRecord(b->ev_side[par], sd));
    b->call_no++;
    return AMPS_OK;
}

extern "C" int amps_recc_iq_batch_submit_dev(amps_recc_iq_batch *b, const void *const *d_iq, const size_t *nsamples, void *cuda_stream) {
    if (!b || !d_iq || !nsamples) return set_error(AMPS_E_INVAL, "null argument");
    if (b->carry) return set_error(AMPS_E_STATE, "host-path samples are pending; keep using amps_recc_iq_batch_work_shared()");
    for (size_t i = 0; i < b->ch.size(); ++i) {
        const amps_recc_iq *h = b->ch[i];
        if (!d_iq[i] && nsamples[i]) return set_error(AMPS_E_INVAL, "null device pointer");
        if (reinterpret_cast<uintptr_t>(d_iq[i]) & 15u) return set_error(AMPS_E_ALIGN, "d_iq[i] must be 16-byte aligned");
        if ((nsamples[i] * h->isz) & 15u) return set_error(AMPS_E_ALIGN, b->sc16 ? "nsamples[i] must be a multiple of 4" : "nsamples[i] must be a multiple of 2");
        if (nsamples[i] > h->max_samples) return set_error(AMPS_E_OVERFLOW, "nsamples[i] exceeds the channel's max_samples");
    }
    CK(cudaSetDevice(b->device));
    return batch_enqueue(b, d_iq, nsamples, static_cast<cudaStream_t>(cuda_stream));
}

// ONE host buffer -> uploaded once -> every channel of the batch (its own center_freq) demodulates it: the carriers of one
// wideband capture share the PCIe transfer.
extern "C" int amps_recc_iq_batch_work_shared(amps_recc_iq_batch *b, const void *iq_host, size_t nsamples, amps_batch_burst_cb cb, void *user) {
    AMPS_NVTX("amps_recc_iq_batch_work_shared");
    if (!b || (!iq_host && nsamples)) return set_error(AMPS_E_INVAL, "null argument");
    uint32_t cap = b->ch[0]->max_samples;
    for (const amps_recc_iq *h : b->ch) { if (h->max_samples < cap) cap = h->max_samples; if (h->dev_carry) return set_error(AMPS_E_STATE, "device-path samples are pending"); }
    if (nsamples > cap) return set_error(AMPS_E_OVERFLOW, "nsamples exceeds max_samples");
    CK(cudaSetDevice(b->device));
    if (!b->d_stage) { b->stage_cap = cap + (uint32_t)kUnit; CK(cudaMalloc(&b->d_stage, (size_t)b->stage_cap * b->isz)); }
    cudaStream_t st = b->stream;
    if (nsamples) CK(cudaMemcpyAsync(b->d_stage + b->carry * b->isz, iq_host, nsamples * b->isz, cudaMemcpyHostToDevice, st));
    const size_t avail = b->carry + nsamples;
    const size_t nq = avail / kUnit;
    b->last_stream = st;
    if (nq) {
        std::vector<const void *> ptrs(b->ch.size(), b->d_stage);
        std::vector<size_t> ns(b->ch.size(), nq * kUnit);
        int rc = batch_enqueue(b, ptrs.data(), ns.data(), st);
        if (rc != AMPS_OK) return rc;
        const size_t left = avail - nq * kUnit;
        if (left) CK(cudaMemcpyAsync(b->d_stage, b->d_stage + nq * kUnit * b->isz, left * b->isz, cudaMemcpyDeviceToDevice, st));
        b->carry = left;
    } else {
        b->carry = avail;
    }
    bool any_ov = false;
    for (size_t i = 0; i < b->ch.size(); ++i) {
        amps_recc_iq *h = b->ch[i];
        uint64_t n = 0;
        bool ov = false;
        int rc = rx_fetch(h, &n, &ov);
        if (rc != AMPS_OK) return rc;
        if (cb) for (uint64_t k = 0; k < n; ++k) cb((int)i, &h->h_ring[(h->consumed + k) % h->max_records], user);
        h->consumed += n;
        h->bursts += n;
        any_ov |= ov;
    }
    return any_ov ? overflow_status(b->ch[0]) : AMPS_OK;
}

extern "C" int amps_recc_iq_batch_front_times(amps_recc_iq_batch *b, float *ms_out, int cap, int *n_out) {
    if (!b || !n_out || (cap > 0 && !ms_out)) return set_error(AMPS_E_INVAL, "null argument");
    *n_out = 0;
    if (!b->timed) return set_error(AMPS_E_STATE, "batch was not created with AMPS_RX_TIME_KERNELS");
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->last_stream));
    return event_times(b->ev0, b->ev1, b->ev_count, amps_recc_iq_batch::kEv, ms_out, cap, n_out);
}

extern "C" int amps_recc_iq_batch_stats(const amps_recc_iq_batch *b, uint64_t *calls, uint64_t *kernel_launches) {
    if (!b) return set_error(AMPS_E_INVAL, "null handle");
    if (calls) *calls = b->call_no;
    if (kernel_launches) *kernel_launches = b->launches;
    return AMPS_OK;
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
