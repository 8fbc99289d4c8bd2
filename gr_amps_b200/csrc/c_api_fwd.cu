// c_api_fwd.cu -- extern "C" entry points of the fused forward path (amps_fwd_*), see include/amps_b200.h.
#include "common.h"
#include "design.h"
#include "fwd_kernels.cuh"

#include <cmath>
#include <cstring>
#include <new>
#include <vector>

using namespace amps;

struct amps_fwd {
    int device = 0, sm_count = 0, ncar = 0;
    cudaStream_t stream = nullptr;
    uint32_t max_sym = 0;
    FwdParams fp{};
    std::vector<float> taps[kFwdMaxCar];
    uint8_t *d_sym[kFwdMaxCar] = {};          // host-path staging
    int32_t *d_sloc[kFwdMaxCar] = {}, *d_btot[kFwdMaxCar] = {}, *d_boff[kFwdMaxCar] = {};
    uint8_t *d_hsym[2][kFwdMaxCar] = {};
    int32_t *d_hS[2][kFwdMaxCar] = {};
    int32_t *d_carry = nullptr;
    int hist_cur = 0;
    float2 *d_out = nullptr;                  // host-path staging
    uint64_t sym_total = 0;
    // Manchester-bit fast path
    FwdBitsParams bp{};
    float2 *d_resp = nullptr;                 // [ncar][2][360] per-bit responses of the x4 interpolator
    float *d_fast = nullptr;                  // [ncar][kFbFastLen] the same, grouped three bits per lookup
    uint8_t *d_bits[kFwdMaxCar] = {};         // host-path staging
    uint8_t *d_hbits[2][kFwdMaxCar] = {};
    int hbits_cur = 0;
    uint64_t bit_total = 0;
    int mode = 0;                             // 0 unset, 1 symbol stream, 2 bit stream (no mixing without reset)
    // voice legs
    bool voice = false;
    VoicePrepParams vp{};
    float *d_audio = nullptr;                 // host-path staging
    float *d_E = nullptr;                     // x25 resampler taps
    double *d_hx[2] = {};
    unsigned long long *d_delta = nullptr, *d_phase = nullptr, *d_vbtot = nullptr, *d_vboff = nullptr;
    float2 *d_vph[kFwdVoiceLegs] = {};
    float2 *d_vhist[2][kFwdVoiceLegs] = {};
    int vhist_cur = 0;
    uint64_t audio_total = 0;
    uint32_t max_audio = 0;
};

static int voice_reset_state(amps_fwd *h) {
    if (!h->voice) return AMPS_OK;
    for (int b = 0; b < 2; ++b) {
        CK(cudaMemset(h->d_hx[b], 0, sizeof(double) * kVoiceImp));
        for (int l = 0; l < kFwdVoiceLegs; ++l) CK(cudaMemset(h->d_vhist[b][l], 0, sizeof(float2) * kFwdVoiceHist));
    }
    CK(cudaMemset(h->d_phase, 0, sizeof(unsigned long long)));
    h->vhist_cur = 0;
    h->audio_total = 0;
    return AMPS_OK;
}

static int fwd_reset_state(amps_fwd *h) {
    for (int b = 0; b < 2; ++b)
        for (int c = 0; c < kFwdMaxCar; ++c) {
            CK(cudaMemset(h->d_hsym[b][c], 0, kFwdHistLen));
            CK(cudaMemset(h->d_hS[b][c], 0, sizeof(int32_t) * kFwdHistLen));
        }
    CK(cudaMemset(h->d_carry, 0, sizeof(int32_t) * kFwdMaxCar));
    for (int b = 0; b < 2; ++b)
        for (int c = 0; c < kFwdMaxCar; ++c) CK(cudaMemset(h->d_hbits[b][c], 0xFF, 16));    // "muted" before the stream starts
    h->hist_cur = 0; h->hbits_cur = 0;
    h->sym_total = 0; h->bit_total = 0; h->mode = 0;
    return voice_reset_state(h);
}

extern "C" int amps_fwd_create(const amps_fwd_params *p, amps_fwd **out) {
    if (!p || !out) return set_error(AMPS_E_INVAL, "null argument");
    *out = nullptr;
    if (p->samp_rate != 10e6 || p->symrate != 100e3)
        return set_error(AMPS_E_INVAL, "samp_rate must be 10e6 and symrate 100e3 (x4 reference interpolator, x25 CIC)");
    if (p->ncarriers < 1 || p->ncarriers > kFwdMaxCar) return set_error(AMPS_E_INVAL, "ncarriers must be 1..3");
    if (p->max_samples == 0) return set_error(AMPS_E_INVAL, "max_samples must be > 0");
    int st = select_device(p->device);
    if (st != AMPS_OK) return st;
    amps_fwd *h = new (std::nothrow) amps_fwd();
    if (!h) return set_error(AMPS_E_NOMEM, "out of host memory");
    h->device = p->device;
    h->ncar = p->ncarriers;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, p->device);
    h->sm_count = prop.multiProcessorCount;
    h->max_sym = (p->max_samples + kFwdInterp - 1) / kFwdInterp;
    std::memset(&h->fp, 0, sizeof h->fp);
    h->fp.ncar = h->ncar;
    h->fp.scale = p->out_scale;
    // frequency_modulator_fc sensitivity 2 pi max_deviation / symrate (grc/ampsbs.grc:614) as a 32-bit phase step
    h->fp.fcw_fm = (uint32_t)(uint64_t)std::llround(p->max_deviation / p->symrate * 4294967296.0);
    std::vector<float> cic5;                                       // boxcar5^3 / 125, 13 taps
    cic3_taps(5, cic5);
    for (int u = 0; u < 15; ++u) h->fp.G2[u] = u < (int)cic5.size() ? p->out_scale * 5.0f * cic5[(size_t)u] : 0.0f;
    for (int c = 0; c < h->ncar; ++c) {
        // pfb interpolator taps at the reference's 400 kS/s: firdes.low_pass(1, 400e3, 10e3, tw) (Hamming), :2172,:2227
        h->taps[c] = firdes_low_pass(1.0, 400e3, 10e3, p->lpf_transition[c], WIN_HAMMING);
        const int n = (int)h->taps[c].size();
        if (n > 4 * kFwdMaxTap4) { delete h; return set_error(AMPS_E_INVAL, "interpolator has more than 324 taps (transition too narrow)"); }
        h->fp.ntap4[c] = (n + 3) / 4;
        for (int i = 0; i < n; ++i) h->fp.taps[c][i] = h->taps[c][(size_t)i];
        const uint32_t fcw = nco_fcw(-p->carrier_freq[c], p->samp_rate);      // shift UP by carrier_freq
        h->fp.fcw_mix25[c] = (uint32_t)(25u * fcw);
        std::vector<float> ph(2 * 75);
        nco_block_table(fcw, 75, ph.data());                       // e^{j phi_c(u)}, u < 75
        for (int u = 0; u < 15; ++u) {                              // tap u of the 2 MS/s stage sits 5 u output samples later
            const float g = u < (int)cic5.size() ? 5.0f * cic5[(size_t)u] : 0.0f;
            h->fp.C1[c][u] = make_float2(g * ph[2 * (size_t)(5 * u)], g * ph[2 * (size_t)(5 * u) + 1]);
            h->fp.C1j[c][u] = make_float2(-h->fp.C1[c][u].y, h->fp.C1[c][u].x);
        }
        h->fp.w25[c] = make_float2(ph[50], ph[51]);
    }
    // ---- polyphase work split over the eight warps of fwd_fused_kernel<false> (FwdParams::seg): at most three warps per
    //      carrier (one owner + two helpers), the next warp always going to the carrier with the most tap groups per warp
    {
        int nslot[kFwdMaxCar] = {0, 0, 0};
        for (int c = 0; c < h->ncar; ++c) nslot[c] = 1;
        for (int used = h->ncar; used < kFwdThreads / 32; ++used) {
            int best = -1;
            for (int c = 0; c < h->ncar; ++c)
                if (nslot[c] < 3 && (best < 0 || h->fp.ntap4[c] * nslot[best] > h->fp.ntap4[best] * nslot[c])) best = c;
            if (best < 0) break;
            nslot[best]++;
        }
        int w = 0, helper = 0;
        for (int c = 0; c < h->ncar; ++c) {
            const int first_helper = helper;
            for (int sl = 0; sl < nslot[c]; ++sl, ++w) {
                FwdParams::Seg &sg = h->fp.seg[w];
                sg.c = (int8_t)c; sg.slot = (int8_t)sl; sg.nslot = (int8_t)nslot[c];
                sg.hidx = (int8_t)(sl == 0 ? first_helper : helper++);
                sg.k0 = (int16_t)((long)h->fp.ntap4[c] * sl / nslot[c]);
                sg.k1 = (int16_t)((long)h->fp.ntap4[c] * (sl + 1) / nslot[c]);
            }
        }
        for (; w < kFwdThreads / 32; ++w) { h->fp.seg[w].c = -1; h->fp.seg[w].slot = 0; h->fp.seg[w].nslot = 0; h->fp.seg[w].hidx = 0; h->fp.seg[w].k0 = h->fp.seg[w].k1 = 0; }
    }
    // ---- Manchester-bit fast path tables: response of the x4 interpolator to one bit's 10 FM samples
    {
        std::memset(&h->bp, 0, sizeof h->bp);
        h->bp.ncar = h->ncar;
        std::vector<float> resp((size_t)h->ncar * 2 * kFbRespLen * 2, 0.0f);
        std::vector<float> fast((size_t)h->ncar * kFbFastLen, 0.0f);
        std::vector<double> r0((size_t)2 * kFbRespLen);             // R0' = (R0 + conj R1) / 2 of the current carrier, in double
        const double kTwoPi = 6.283185307179586476925286766559;
        for (int c = 0; c < h->ncar; ++c) {
            const int nt = (int)h->taps[c].size();
            for (int b = 0; b < 2; ++b) {
                for (int u = 0; u < kFbRespLen; ++u) {
                    double re = 0, im = 0;
                    int S = 0;
                    for (int i = 0; i < kFbSymPerBit; ++i) {
                        const int half_sym = i < 5 ? (b ? -1 : +1) : (b ? +1 : -1);     // bit 1 -> (low, high), bit 0 -> (high, low)
                        S += half_sym;
                        const int k = u - 4 * i;
                        if (k < 0 || k >= nt) continue;
                        const uint32_t psi = (uint32_t)S * h->fp.fcw_fm;
                        const double ang = kTwoPi * ((double)psi / 4294967296.0);
                        re += (double)h->taps[c][(size_t)k] * std::cos(ang);
                        im += (double)h->taps[c][(size_t)k] * std::sin(ang);
                    }
                    const size_t o = (((size_t)c * 2 + (size_t)b) * kFbRespLen + (size_t)u) * 2;
                    resp[o] = (float)re; resp[o + 1] = (float)im;
                    // a Manchester 1 mirrors a 0, so R1 = conj(R0) up to the rounding of the two cos/sin evaluations
                    if (b == 0) { r0[2 * (size_t)u] = 0.5 * re; r0[2 * (size_t)u + 1] = 0.5 * im; }
                    else { r0[2 * (size_t)u] += 0.5 * re; r0[2 * (size_t)u + 1] -= 0.5 * im; }
                }
            }
            const uint32_t fcw = nco_fcw(-p->carrier_freq[c], p->samp_rate);
            h->bp.fcw_mix1000[c] = (uint32_t)(1000u * fcw);
            std::vector<float> ph(2 * kFbMPerBit);
            nco_block_table((uint32_t)(25u * fcw), kFbMPerBit, ph.data());       // e^{j phi_c(25 u)}, u < 40
            for (int u = 0; u < kFbMPerBit; ++u) h->bp.w40[c][u] = make_float2(ph[2 * (size_t)u], ph[2 * (size_t)u + 1]);
            std::memcpy(h->bp.C1[c], h->fp.C1[c], sizeof h->bp.C1[c]);
            // grouped tables (fwd_kernels.cuh: kFbFastLen)
            float *F = fast.data() + (size_t)c * kFbFastLen;
            for (int u = 0; u < kFbMPerBit; ++u) {
                double resum = 0;
                for (int d = 0; d < kFbRespBits; ++d) resum += r0[2 * (size_t)(u + kFbMPerBit * d)];
                const double wr = (double)h->bp.w40[c][u].x, wi = (double)h->bp.w40[c][u].y;
                F[2 * u] = (float)(resum * wr); F[2 * u + 1] = (float)(resum * wi);                       // RW
                F[2 * kFbMPerBit + 2 * u] = (float)-wi; F[2 * kFbMPerBit + 2 * u + 1] = (float)wr;        // JW = j w40
                for (int G = 0; G < 3; ++G)
                    for (int pat = 0; pat < 8; ++pat) {
                        double acc = 0;
                        for (int k = 0; k < 3; ++k) {
                            const double im = r0[2 * (size_t)(u + kFbMPerBit * (8 - 3 * G - k)) + 1];
                            acc += ((pat >> k) & 1) ? -im : im;
                        }
                        F[4 * kFbMPerBit + (G * 8 + pat) * kFbMPerBit + u] = (float)acc;
                    }
            }
        }
        std::memcpy(h->bp.G2, h->fp.G2, sizeof h->bp.G2);
        CK(cudaMalloc(&h->d_fast, fast.size() * sizeof(float)));
        CK(cudaMemcpy(h->d_fast, fast.data(), fast.size() * sizeof(float), cudaMemcpyHostToDevice));
        h->bp.fast = h->d_fast;
        CK(cudaMalloc(&h->d_resp, resp.size() * sizeof(float)));
        CK(cudaMemcpy(h->d_resp, resp.data(), resp.size() * sizeof(float), cudaMemcpyHostToDevice));
        h->bp.resp = h->d_resp;
    }
    cudaError_t ce = fwd_configure_device();
    if (ce != cudaSuccess) { delete h; return set_cuda_error(ce, "fwd_configure_device"); }
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    const size_t nblk = ((size_t)h->max_sym + kFwdScanBlock - 1) / kFwdScanBlock;
    for (int c = 0; c < kFwdMaxCar; ++c) {
        CK(cudaMalloc(&h->d_sym[c], h->max_sym));
        CK(cudaMalloc(&h->d_sloc[c], sizeof(int32_t) * h->max_sym));
        CK(cudaMalloc(&h->d_btot[c], sizeof(int32_t) * nblk));
        CK(cudaMalloc(&h->d_boff[c], sizeof(int32_t) * nblk));
        for (int b = 0; b < 2; ++b) {
            CK(cudaMalloc(&h->d_hsym[b][c], kFwdHistLen));
            CK(cudaMalloc(&h->d_hS[b][c], sizeof(int32_t) * kFwdHistLen));
        }
    }
    for (int c = 0; c < kFwdMaxCar; ++c) {
        CK(cudaMalloc(&h->d_bits[c], h->max_sym / kFbSymPerBit + 16));
        for (int b = 0; b < 2; ++b) CK(cudaMalloc(&h->d_hbits[b][c], 16));
    }
    CK(cudaMalloc(&h->d_carry, sizeof(int32_t) * kFwdMaxCar));
    CK(cudaMalloc(&h->d_out, sizeof(float2) * (size_t)h->max_sym * kFwdInterp));
    st = fwd_reset_state(h);
    if (st != AMPS_OK) { amps_fwd_destroy(h); return st; }
    *out = h;
    return AMPS_OK;
}

extern "C" int amps_fwd_destroy(amps_fwd *h) {
    if (!h) return AMPS_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->stream) cudaStreamDestroy(h->stream);
    for (int c = 0; c < kFwdMaxCar; ++c) {
        cudaFree(h->d_sym[c]); cudaFree(h->d_sloc[c]); cudaFree(h->d_btot[c]); cudaFree(h->d_boff[c]);
        for (int b = 0; b < 2; ++b) { cudaFree(h->d_hsym[b][c]); cudaFree(h->d_hS[b][c]); }
    }
    for (int c = 0; c < kFwdMaxCar; ++c) { cudaFree(h->d_bits[c]); cudaFree(h->d_hbits[0][c]); cudaFree(h->d_hbits[1][c]); }
    cudaFree(h->d_carry); cudaFree(h->d_out); cudaFree(h->d_resp); cudaFree(h->d_fast);
    cudaFree(h->d_audio); cudaFree(h->d_E); cudaFree(h->d_delta); cudaFree(h->d_phase); cudaFree(h->d_vbtot); cudaFree(h->d_vboff);
    for (int b = 0; b < 2; ++b) { cudaFree(h->d_hx[b]); for (int l = 0; l < kFwdVoiceLegs; ++l) cudaFree(h->d_vhist[b][l]); }
    for (int l = 0; l < kFwdVoiceLegs; ++l) cudaFree(h->d_vph[l]);
    delete h;
    return AMPS_OK;
}

extern "C" int amps_fwd_reset(amps_fwd *h) {
    if (!h) return set_error(AMPS_E_INVAL, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    return fwd_reset_state(h);
}

extern "C" int amps_fwd_interp(const amps_fwd *h) { (void)h; return kFwdInterp; }

extern "C" int amps_fwd_get_taps(const amps_fwd *h, int carrier, float *out, int cap) {
    if (!h || carrier < 0 || carrier >= h->ncar) return set_error(AMPS_E_INVAL, "bad argument");
    const int n = (int)h->taps[carrier].size();
    if (out) for (int i = 0; i < n && i < cap; ++i) out[i] = h->taps[carrier][(size_t)i];
    return n;
}

static int fwd_submit_symbols(amps_fwd *h, const void *const *d_sym, size_t nsym, void *d_out_iq, cudaStream_t st, bool with_voice);

extern "C" int amps_fwd_submit_dev(amps_fwd *h, const void *const *d_sym, size_t nsym, void *d_out_iq, void *cuda_stream) {
    if (!h || !d_sym || (nsym && !d_out_iq)) return set_error(AMPS_E_INVAL, "null argument");
    if (h->voice) return set_error(AMPS_E_STATE, "voice legs are enabled: use amps_fwd_submit_voice_dev / amps_fwd_work_voice");
    return fwd_submit_symbols(h, d_sym, nsym, d_out_iq, static_cast<cudaStream_t>(cuda_stream), false);
}

static int fwd_submit_symbols(amps_fwd *h, const void *const *d_sym, size_t nsym, void *d_out_iq, cudaStream_t st, bool with_voice) {
    AMPS_NVTX("amps_fwd: submit half-symbols");
    if (nsym == 0) return AMPS_OK;
    if (nsym > h->max_sym) return set_error(AMPS_E_OVERFLOW, "nsym exceeds max_samples / 100");
    if (reinterpret_cast<uintptr_t>(d_out_iq) & 15u) return set_error(AMPS_E_ALIGN, "d_out_iq must be 16-byte aligned");
    for (int c = 0; c < h->ncar; ++c) if (!d_sym[c]) return set_error(AMPS_E_INVAL, "null symbol stream");
    if (h->mode == 2) return set_error(AMPS_E_STATE, "handle is streaming data bits; reset() before switching to half-symbol input");
    h->mode = 1;
    CK(cudaSetDevice(h->device));
    const int cur = h->hist_cur, nxt = cur ^ 1;
    FwdScanParams sp{};
    FwdParams p = h->fp;
    for (int c = 0; c < kFwdMaxCar; ++c) {
        const int cc = c < h->ncar ? c : 0;
        sp.sym[c] = static_cast<const uint8_t *>(d_sym[cc]);
        sp.sloc[c] = h->d_sloc[c]; sp.btot[c] = h->d_btot[c]; sp.boff[c] = h->d_boff[c];
        sp.hsym_old[c] = h->d_hsym[cur][c]; sp.hS_old[c] = h->d_hS[cur][c];
        sp.hsym_new[c] = h->d_hsym[nxt][c]; sp.hS_new[c] = h->d_hS[nxt][c];
        p.sym[c] = sp.sym[c]; p.sloc[c] = h->d_sloc[c]; p.boff[c] = h->d_boff[c];
        p.hsym[c] = h->d_hsym[cur][c]; p.hS[c] = h->d_hS[cur][c];
    }
    sp.carry = h->d_carry;
    sp.nsym = (uint32_t)nsym;
    CKL(launch_fwd_scan(sp, h->ncar, st));
    p.out = static_cast<float2 *>(d_out_iq);
    p.nsym = (uint32_t)nsym;
    p.m_base = (uint32_t)(h->sym_total * 4u);
    const uint32_t ntiles = ((uint32_t)nsym + kFwdTileSym - 1) / kFwdTileSym;
    uint32_t grid = 3u * (uint32_t)h->sm_count;                  // 74 KB smem, 72 registers: 3 CTAs per SM
    if (grid > ntiles) grid = ntiles;
    if (with_voice) CKL(launch_fwd_fused_voice(p, (int)grid, st));
    else CKL(launch_fwd_fused(p, (int)grid, st));
    h->hist_cur = nxt;
    h->sym_total += nsym;
    return AMPS_OK;
}

extern "C" int amps_fwd_work(amps_fwd *h, const uint8_t *const *sym, size_t nsym, float *out_iq_host) {
    AMPS_NVTX("amps_fwd_work");
    if (!h || !sym || (nsym && !out_iq_host)) return set_error(AMPS_E_INVAL, "null argument");
    if (nsym == 0) return AMPS_OK;
    if (nsym > h->max_sym) return set_error(AMPS_E_OVERFLOW, "nsym exceeds max_samples / 100");
    CK(cudaSetDevice(h->device));
    const void *dptr[kFwdMaxCar] = {};
    for (int c = 0; c < h->ncar; ++c) {
        if (!sym[c]) return set_error(AMPS_E_INVAL, "null symbol stream");
        CK(cudaMemcpyAsync(h->d_sym[c], sym[c], nsym, cudaMemcpyHostToDevice, h->stream));
        dptr[c] = h->d_sym[c];
    }
    int rc = amps_fwd_submit_dev(h, dptr, nsym, h->d_out, h->stream);
    if (rc != AMPS_OK) return rc;
    CK(cudaMemcpyAsync(out_iq_host, h->d_out, sizeof(float2) * nsym * kFwdInterp, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return AMPS_OK;
}

// ---------------------------------------------------------------------------------------------
// voice legs: audio @16 kS/s -> nbfm_tx -> x25 arb resampler, added in front of a carrier's mixer
// ---------------------------------------------------------------------------------------------
extern "C" int amps_fwd_enable_voice(amps_fwd *h, const amps_fwd_voice_params *vp) {
    if (!h || !vp) return set_error(AMPS_E_INVAL, "null argument");
    if (h->voice) return set_error(AMPS_E_STATE, "voice legs are already enabled");
    if (h->mode != 0) return set_error(AMPS_E_STATE, "enable voice before streaming (or after reset())");
    if (vp->audio_rate != 16000.0) return set_error(AMPS_E_INVAL, "audio_rate must be 16000 (x25 to 400 kS/s)");
    if (vp->carrier_gated >= h->ncar || vp->carrier_open >= h->ncar || (vp->carrier_gated < 0 && vp->carrier_open < 0))
        return set_error(AMPS_E_INVAL, "voice legs need a carrier index below ncarriers");
    if (!(vp->max_dev > 0) || !(vp->tau > 0)) return set_error(AMPS_E_INVAL, "max_dev and tau must be positive");
    // the SAT is generated from an 8-entry table: 6000 / 16000 = 3 / 8 turn per audio sample
    if (vp->sat_amp != 0.0 && vp->sat_freq != 6000.0) return set_error(AMPS_E_INVAL, "sat_freq must be 6000");
    CK(cudaSetDevice(h->device));
    const double kTwoPi = 6.283185307179586476925286766559;
    std::memset(&h->vp, 0, sizeof h->vp);
    const std::vector<double> g = fm_preemph_impulse(vp->audio_rate, vp->tau, -1.0, kVoiceImp);    // nbfm_tx fh = -1 (grc :742)
    for (int k = 0; k < kVoiceImp; ++k) h->vp.g[k] = g[(size_t)k];
    for (int k = 0; k < 8; ++k) h->vp.sat[k] = vp->sat_amp * std::cos(kTwoPi * (double)((3 * k) % 8) / 8.0);
    h->vp.cycles_per_unit = vp->max_dev / vp->audio_rate;
    // voice_lpf_taps (grc :138-184 sibling block): firdes.low_pass(3, 400e3, 15e3, 6e3, BLACKMAN), 225 taps
    int per = 0;
    const std::vector<float> E = arb25_taps(firdes_low_pass(3.0, 400e3, 15e3, 6e3, WIN_BLACKMAN), per);
    if (per > kFwdVoicePer) return set_error(AMPS_E_INVAL, "voice resampler has too many taps per arm");
    h->fp.vper = per;
    std::vector<float> Epad((size_t)25 * kFwdVoicePer, 0.0f);
    for (int r = 0; r < 25; ++r)
        for (int k = 0; k < per; ++k) Epad[(size_t)r * kFwdVoicePer + k] = E[(size_t)r * per + k];
    CK(cudaMalloc(&h->d_E, Epad.size() * sizeof(float)));
    CK(cudaMemcpy(h->d_E, Epad.data(), Epad.size() * sizeof(float), cudaMemcpyHostToDevice));
    h->fp.Eg = h->d_E;
    if (vp->carrier_gated == vp->carrier_open) return set_error(AMPS_E_INVAL, "the two voice legs must feed different carriers");
    h->fp.vcar[0] = vp->carrier_gated;
    h->fp.vcar[1] = vp->carrier_open;
    h->max_audio = (uint32_t)(((uint64_t)h->max_sym * 4u + 24u) / 25u);
    CK(cudaMalloc(&h->d_audio, sizeof(float) * h->max_audio));
    CK(cudaMalloc(&h->d_delta, sizeof(unsigned long long) * h->max_audio));
    CK(cudaMalloc(&h->d_phase, sizeof(unsigned long long)));
    CK(cudaMalloc(&h->d_vbtot, sizeof(unsigned long long) * (h->max_audio / 256 + 1)));
    CK(cudaMalloc(&h->d_vboff, sizeof(unsigned long long) * (h->max_audio / 256 + 1)));
    for (int b = 0; b < 2; ++b) {
        CK(cudaMalloc(&h->d_hx[b], sizeof(double) * kVoiceImp));
        for (int l = 0; l < kFwdVoiceLegs; ++l) CK(cudaMalloc(&h->d_vhist[b][l], sizeof(float2) * kFwdVoiceHist));
    }
    for (int l = 0; l < kFwdVoiceLegs; ++l) CK(cudaMalloc(&h->d_vph[l], sizeof(float2) * h->max_audio));
    h->voice = true;
    return voice_reset_state(h);
}

extern "C" int amps_fwd_submit_voice_dev(amps_fwd *h, const void *const *d_sym, const void *d_audio, size_t nsym, int audio_mute,
                                         void *d_out_iq, void *cuda_stream) {
    if (!h || !d_sym || (nsym && (!d_out_iq || !d_audio))) return set_error(AMPS_E_INVAL, "null argument");
    if (!h->voice) return set_error(AMPS_E_STATE, "amps_fwd_enable_voice() was not called");
    if (nsym == 0) return AMPS_OK;
    if (nsym % 25) return set_error(AMPS_E_ALIGN, "nsym must be a multiple of 25 (whole 16 kS/s audio samples)");
    if (nsym > h->max_sym) return set_error(AMPS_E_OVERFLOW, "nsym exceeds max_samples / 100");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const uint32_t n_audio = (uint32_t)(nsym * 4 / 25);
    const int cur = h->vhist_cur, nxt = cur ^ 1;
    VoicePrepParams vp = h->vp;
    vp.audio = static_cast<const float *>(d_audio);
    vp.hx_old = h->d_hx[cur]; vp.hx_new = h->d_hx[nxt];
    vp.delta = h->d_delta; vp.phase = h->d_phase; vp.btot = h->d_vbtot; vp.boff = h->d_vboff;
    for (int l = 0; l < kFwdVoiceLegs; ++l) {
        vp.vph[l] = h->fp.vcar[l] >= 0 ? h->d_vph[l] : nullptr;
        vp.vhist_old[l] = h->d_vhist[cur][l]; vp.vhist_new[l] = h->d_vhist[nxt][l];
    }
    vp.leg_muted[0] = audio_mute ? 1 : 0;                       // mute_xx(audio_mute) sits in front of the gated leg only
    vp.leg_muted[1] = 0;
    vp.a_base = h->audio_total;
    vp.n_audio = n_audio;
    CKL(launch_voice_prep(vp, st));
    for (int l = 0; l < kFwdVoiceLegs; ++l) { h->fp.vph[l] = h->d_vph[l]; h->fp.vhist[l] = h->d_vhist[cur][l]; }
    h->fp.n_audio = n_audio;
    int rc = fwd_submit_symbols(h, d_sym, nsym, d_out_iq, st, true);
    if (rc != AMPS_OK) return rc;
    h->vhist_cur = nxt;
    h->audio_total += n_audio;
    return AMPS_OK;
}

extern "C" int amps_fwd_work_voice(amps_fwd *h, const uint8_t *const *sym, const float *audio, size_t nsym, int audio_mute,
                                   float *out_iq_host) {
    if (!h || !sym || (nsym && (!out_iq_host || !audio))) return set_error(AMPS_E_INVAL, "null argument");
    if (!h->voice) return set_error(AMPS_E_STATE, "amps_fwd_enable_voice() was not called");
    if (nsym == 0) return AMPS_OK;
    if (nsym % 25) return set_error(AMPS_E_ALIGN, "nsym must be a multiple of 25 (whole 16 kS/s audio samples)");
    if (nsym > h->max_sym) return set_error(AMPS_E_OVERFLOW, "nsym exceeds max_samples / 100");
    CK(cudaSetDevice(h->device));
    const void *dptr[kFwdMaxCar] = {};
    for (int c = 0; c < h->ncar; ++c) {
        if (!sym[c]) return set_error(AMPS_E_INVAL, "null symbol stream");
        CK(cudaMemcpyAsync(h->d_sym[c], sym[c], nsym, cudaMemcpyHostToDevice, h->stream));
        dptr[c] = h->d_sym[c];
    }
    CK(cudaMemcpyAsync(h->d_audio, audio, sizeof(float) * (nsym * 4 / 25), cudaMemcpyHostToDevice, h->stream));
    int rc = amps_fwd_submit_voice_dev(h, dptr, h->d_audio, nsym, audio_mute, h->d_out, h->stream);
    if (rc != AMPS_OK) return rc;
    CK(cudaMemcpyAsync(out_iq_host, h->d_out, sizeof(float2) * nsym * kFwdInterp, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return AMPS_OK;
}

// ---------------------------------------------------------------------------------------------
// Manchester-bit fast path: one byte per 10 kbit/s data bit (0, 1, 0xFF = muted), 1000 output samples per bit
// ---------------------------------------------------------------------------------------------
extern "C" int amps_fwd_submit_bits_dev(amps_fwd *h, const void *const *d_bits, size_t nbits, void *d_out_iq, void *cuda_stream) {
    AMPS_NVTX("amps_fwd_submit_bits_dev");
    if (!h || !d_bits || (nbits && !d_out_iq)) return set_error(AMPS_E_INVAL, "null argument");
    if (nbits == 0) return AMPS_OK;
    if (nbits * kFbSymPerBit > h->max_sym) return set_error(AMPS_E_OVERFLOW, "nbits exceeds max_samples / 1000");
    if (reinterpret_cast<uintptr_t>(d_out_iq) & 15u) return set_error(AMPS_E_ALIGN, "d_out_iq must be 16-byte aligned");
    for (int c = 0; c < h->ncar; ++c) if (!d_bits[c]) return set_error(AMPS_E_INVAL, "null bit stream");
    if (h->mode == 1) return set_error(AMPS_E_STATE, "handle is streaming half-symbols; reset() before switching to data-bit input");
    if (h->voice) return set_error(AMPS_E_STATE, "voice legs need the half-symbol input (amps_fwd_work_voice)");
    h->mode = 2;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int cur = h->hbits_cur, nxt = cur ^ 1;
    FwdBitsParams p = h->bp;
    for (int c = 0; c < kFwdMaxCar; ++c) {
        const int cc = c < h->ncar ? c : 0;
        p.bits[c] = static_cast<const uint8_t *>(d_bits[cc]);
        p.hbits[c] = h->d_hbits[cur][c];
    }
    p.out = static_cast<float2 *>(d_out_iq);
    p.nbits = (uint32_t)nbits;
    p.bit_base = h->bit_total;
    const uint32_t ntiles = ((uint32_t)nbits + kFbTileBits - 1) / kFbTileBits;
    uint32_t grid = 3u * (uint32_t)h->sm_count;
    if (grid > ntiles) grid = ntiles;
    CKL(launch_fwd_bits(p, (int)grid, st));
    // history for the next call: the last kFbHistBits of (old history ++ these bits)
    for (int c = 0; c < h->ncar; ++c) {
        if (nbits >= (size_t)kFbHistBits) {
            CK(cudaMemcpyAsync(h->d_hbits[nxt][c], p.bits[c] + (nbits - kFbHistBits), kFbHistBits, cudaMemcpyDeviceToDevice, st));
        } else {
            CK(cudaMemcpyAsync(h->d_hbits[nxt][c], h->d_hbits[cur][c] + nbits, kFbHistBits - nbits, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(h->d_hbits[nxt][c] + (kFbHistBits - nbits), p.bits[c], nbits, cudaMemcpyDeviceToDevice, st));
        }
    }
    h->hbits_cur = nxt;
    h->bit_total += nbits;
    return AMPS_OK;
}

extern "C" int amps_fwd_work_bits(amps_fwd *h, const uint8_t *const *bits, size_t nbits, float *out_iq_host) {
    AMPS_NVTX("amps_fwd_work_bits");
    if (!h || !bits || (nbits && !out_iq_host)) return set_error(AMPS_E_INVAL, "null argument");
    if (nbits == 0) return AMPS_OK;
    if (nbits * kFbSymPerBit > h->max_sym) return set_error(AMPS_E_OVERFLOW, "nbits exceeds max_samples / 1000");
    CK(cudaSetDevice(h->device));
    const void *dptr[kFwdMaxCar] = {};
    for (int c = 0; c < h->ncar; ++c) {
        if (!bits[c]) return set_error(AMPS_E_INVAL, "null bit stream");
        CK(cudaMemcpyAsync(h->d_bits[c], bits[c], nbits, cudaMemcpyHostToDevice, h->stream));
        dptr[c] = h->d_bits[c];
    }
    int rc = amps_fwd_submit_bits_dev(h, dptr, nbits, h->d_out, h->stream);
    if (rc != AMPS_OK) return rc;
    CK(cudaMemcpyAsync(out_iq_host, h->d_out, sizeof(float2) * nbits * kFbOutPerBit, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return AMPS_OK;
}
