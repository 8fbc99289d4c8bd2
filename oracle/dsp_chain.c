/*
 * dsp_chain.c -- ORACLE (test infrastructure).  CPU restatement of the per-sample DSP chain.
 *
 * What the reference does (grc/ampsbs.grc, stock GNU Radio 3.7 blocks whose source is NOT in
 * /root/reference; equations per SURVEY.md App. B):
 *   RX @400 kS/s: freq_xlating_fir_filter_ccc(decim 2, lpf_taps[299], -160 kHz) (:1814-1872,
 *   :138-184) -> quadrature_demod_cf(1) (:774-816) -> clock_recovery_mm_ff + binary_slicer_fb
 *   (:1751-1813, :1712-1750) -> amps.recc (lib/recc_impl.cc:93-145).
 *
 * What is restated here is the 10 MS/s extrapolation BASELINE.json asks for ("kernel-spec",
 * DESIGN.md section 3).  x[n] e^{-j theta n} followed by real-tap low-pass filtering is
 * algebraically the frequency-translating FIR; the first-stage CIC^3 /25 brings 10 MS/s down to
 * the reference's 400 kS/s where the reference's own 299-tap firdes filter /2 and the quadrature
 * demod run unchanged; symbol timing is feed-forward (exact 74/74 trigger match as in
 * recc_impl.cc:118, sampling phase = soft-correlation peak inside the run of matching phases).
 *
 *   f64 flavour: direct double arithmetic with libm -- the "ideal" chain (tolerance tests).
 *   f32 flavour: the exact fp32 operation order of the CUDA kernel -- bit-exact hard decisions.
 *
 * Build with -ffp-contract=off: every fused multiply-add below is an explicit fmaf().
 */
#include "amps_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ firdes.low_pass */
int orc_firdes_low_pass(double gain, double fs, double fc, double tw, int window, float *taps, int cap) {
    static const double max_atten[3] = {53.0, 44.0, 74.0};   /* hamming, hann, blackman */
    int ntaps = (int)(max_atten[window] * fs / (22.0 * tw));
    if ((ntaps & 1) == 0) ntaps++;
    if (!taps || cap < ntaps) return ntaps;
    int M = (ntaps - 1) / 2;
    double fwT0 = 2.0 * M_PI * fc / fs;
    double *h = (double *)malloc(sizeof(double) * (size_t)ntaps);
    /* firdes computes the window in float and the products in float; mirror that */
    for (int n = -M; n <= M; n++) {
        int k = n + M;
        double wv;
        if (window == 0) wv = 0.54 - 0.46 * cos(2.0 * M_PI * k / (ntaps - 1));
        else if (window == 1) wv = 0.5 - 0.5 * cos(2.0 * M_PI * k / (ntaps - 1));
        else wv = 0.42 - 0.5 * cos(2.0 * M_PI * k / (ntaps - 1)) + 0.08 * cos(4.0 * M_PI * k / (ntaps - 1));
        float wf = (float)wv;
        if (n == 0) h[k] = (float)(fwT0 / M_PI * wf);
        else h[k] = (float)(sin(n * fwT0) / (n * M_PI) * wf);
    }
    double fmax = h[M];
    for (int n = 1; n <= M; n++) fmax += 2.0 * h[n + M];
    double g = gain / fmax;
    for (int i = 0; i < ntaps; i++) taps[i] = (float)(h[i] * g);
    free(h);
    return ntaps;
}

/* ------------------------------------------------------------------ NCO */
uint32_t orc_nco_fcw(double center_freq, double samp_rate) {
    /* rotate by -center_freq: phase increment per sample, in 2^-32 turns */
    double turns = -center_freq / samp_rate;
    turns -= floor(turns);
    return (uint32_t)(uint64_t)llround(turns * 4294967296.0);
}

#define D1 25
#define NCIC 73
static void cic_coeffs(double c[NCIC]) {
    double a[49];
    memset(a, 0, sizeof a);
    for (int i = 0; i < 25; i++) for (int j = 0; j < 25; j++) a[i + j] += 1.0;
    memset(c, 0, sizeof(double) * NCIC);
    for (int i = 0; i < 49; i++) for (int j = 0; j < 25; j++) c[i + j] += a[i];
    for (int i = 0; i < NCIC; i++) c[i] /= 15625.0;
}

/* ------------------------------------------------------------------ f64 "ideal" chain */
void orc_rx_chain_f64(const float *iq, size_t n, uint32_t fcw, const float *h2, int nh2,
                      double *y_out, double *d_out) {
    size_t nv = n / D1, nq = nv / 2;
    double c[NCIC];
    cic_coeffs(c);
    double *ur = (double *)malloc(sizeof(double) * n), *ui = (double *)malloc(sizeof(double) * n);
    for (size_t i = 0; i < n; i++) {
        uint32_t psi = (uint32_t)((uint64_t)i * fcw);
        double ang = 2.0 * M_PI * ((double)psi / 4294967296.0);
        double cr = cos(ang), ci = sin(ang);
        double xr = iq[2 * i], xi = iq[2 * i + 1];
        ur[i] = xr * cr - xi * ci;
        ui[i] = xr * ci + xi * cr;
    }
    double *vr = (double *)calloc(nv, sizeof(double)), *vi = (double *)calloc(nv, sizeof(double));
    for (size_t m = 0; m < nv; m++) {
        double ar = 0, ai = 0;
        for (int t = 0; t < NCIC; t++) {
            long idx = (long)(D1 * (m + 1)) - 1 - t;
            if (idx < 0) break;
            ar += c[t] * ur[idx];
            ai += c[t] * ui[idx];
        }
        vr[m] = ar; vi[m] = ai;
    }
    double pr = 0, pi_ = 0;
    for (size_t q = 0; q < nq; q++) {
        double ar = 0, ai = 0;
        for (int k = 0; k < nh2; k++) {
            long idx = (long)(2 * q) - k;
            if (idx < 0) break;
            ar += (double)h2[k] * vr[idx];
            ai += (double)h2[k] * vi[idx];
        }
        if (y_out) { y_out[2 * q] = ar; y_out[2 * q + 1] = ai; }
        double zr = ar * pr + ai * pi_, zi = ai * pr - ar * pi_;
        if (d_out) d_out[q] = (zr == 0.0 && zi == 0.0) ? 0.0 : atan2(zi, zr);
        pr = ar; pi_ = ai;
    }
    free(ur); free(ui); free(vr); free(vi);
}

/* ------------------------------------------------------------------ f32 kernel-spec pieces */
/* sin/cos of a 32-bit phase (2^32 = one turn): nearest-quadrant reduction, Cephes-style minimax
 * polynomials on [-pi/4, pi/4], evaluated with explicit fmaf in this exact order. */
static void spec_sincos(uint32_t psi, float *c, float *s) {
    uint32_t quad = (psi + 0x20000000u) >> 30;
    int32_t frac = (int32_t)(psi - (quad << 30));
    float a = (float)frac * 1.46291807926715968e-9f;           /* pi / 2^31 */
    float z = a * a;
    float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f), z * a, a);
    float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f), z * z,
                    fmaf(-0.5f, z, 1.0f));
    switch (quad & 3u) {
        case 0: *c = cp;  *s = sp;  break;
        case 1: *c = -sp; *s = cp;  break;
        case 2: *c = -cp; *s = -sp; break;
        default: *c = sp; *s = -cp; break;
    }
}

static float spec_atan2(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
    if (mx == 0.0f) return 0.0f;
    float r = mn / mx;
    int big = r > 0.4142135679721832f;
    float t = big ? (r - 1.0f) / (r + 1.0f) : r;
    float z = t * t;
    float p = fmaf(fmaf(fmaf(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z, -3.33329491539e-1f);
    float a = fmaf(p * z, t, t);
    if (big) a = a + 0.785398163397448309f;
    if (ay > ax) a = 1.57079632679489662f - a;
    if (x < 0.0f) a = 3.14159265358979324f - a;
    if (y < 0.0f) a = -a;
    return a;
}

/* arg(y conj(p)) in the kernel's operation order */
static float spec_qdemod(float yr, float yi, float pr, float pi_) {
    float zr = fmaf(yi, pi_, yr * pr);
    float zi = fmaf(yi, pr, -(yr * pi_));
    return spec_atan2(zi, zr);
}

/* stage 2 + demod of the fp32 kernel-spec chain: y[q] = E + O (FMA chains over even / odd taps, k ascending),
 * d[q] = atan2_spec(Im, Re) of y[q] conj(y[q-1]); shared by the 10 MS/s and the native 400 kS/s front ends */
static void spec_stage2_demod(const float *vr, const float *vi, size_t nq, const float *h2, int nh2, float *y_out, float *d_out) {
    float pr = 0, pi_ = 0;
    for (size_t q = 0; q < nq; q++) {
        float er = 0, ei = 0, orr = 0, oi = 0;
        for (int k = 0; k < nh2; k += 2) {
            long idx = (long)(2 * q) - k;
            float a = idx >= 0 ? vr[idx] : 0.0f, b = idx >= 0 ? vi[idx] : 0.0f;
            er = fmaf(h2[k], a, er); ei = fmaf(h2[k], b, ei);
        }
        for (int k = 1; k < nh2; k += 2) {
            long idx = (long)(2 * q) - k;
            float a = idx >= 0 ? vr[idx] : 0.0f, b = idx >= 0 ? vi[idx] : 0.0f;
            orr = fmaf(h2[k], a, orr); oi = fmaf(h2[k], b, oi);
        }
        float yr = er + orr, yi = ei + oi;
        if (y_out) { y_out[2 * q] = yr; y_out[2 * q + 1] = yi; }
        if (d_out) d_out[q] = spec_qdemod(yr, yi, pr, pi_);
        pr = yr; pi_ = yi;
    }
}

/* quadrature_demod_cf on its own (gain applied in the caller's precision), zero history: the f32 flavour is the kernel's
 * operation order (spec_qdemod), the f64 one is atan2 -- what GNU Radio's qa_quadrature_demod.py vector is held against */
void orc_quad_demod(const float *iq, size_t n, float *d32, double *d64) {
    float pr = 0, pi_ = 0;
    double qr = 0, qi = 0;
    for (size_t i = 0; i < n; i++) {
        float yr = iq[2 * i], yi = iq[2 * i + 1];
        if (d32) d32[i] = spec_qdemod(yr, yi, pr, pi_);
        double zr = (double)yr * qr + (double)yi * qi, zi = (double)yi * qr - (double)yr * qi;
        if (d64) d64[i] = (zr == 0.0 && zi == 0.0) ? 0.0 : atan2(zi, zr);
        pr = yr; pi_ = yi; qr = yr; qi = yi;
    }
}

/* blk0: absolute index (mod 2^32) of the stream's 25-sample block that iq[0] starts -- the NCO phase is a function of the
 * absolute sample index, so a buffer cut out of a longer stream is demodulated exactly as it was inside the stream */
void orc_rx_chain_f32_at(const float *iq, size_t n, uint32_t fcw, const float *h2, int nh2, uint32_t blk0,
                         float *y_out, float *d_out);
void orc_rx_chain_f32(const float *iq, size_t n, uint32_t fcw, const float *h2, int nh2,
                      float *y_out, float *d_out) {
    orc_rx_chain_f32_at(iq, n, fcw, h2, nh2, 0u, y_out, d_out);
}
void orc_rx_chain_f32_at(const float *iq, size_t n, uint32_t fcw, const float *h2, int nh2, uint32_t blk0,
                         float *y_out, float *d_out) {
    size_t nv = n / D1, nq = nv / 2;
    double cd[NCIC];
    float g[75];
    cic_coeffs(cd);
    for (int t = 0; t < 75; t++) g[t] = t < NCIC ? (float)cd[t] : 0.0f;
    float wr[D1], wi[D1];
    for (int k = 0; k < D1; k++) {
        uint32_t psi = (uint32_t)((uint32_t)k * fcw);
        double ang = 2.0 * M_PI * ((double)psi / 4294967296.0);
        wr[k] = (float)cos(ang); wi[k] = (float)sin(ang);
    }
    uint32_t fcw25 = (uint32_t)(25u * fcw);
    /* rotated partial sums P0', P1', P2' per 25-sample block */
    float *P = (float *)malloc(sizeof(float) * 6 * (nv ? nv : 1));
    for (size_t b = 0; b < nv; b++) {
        float pr[3] = {0, 0, 0}, pi_[3] = {0, 0, 0};
        for (int k = 0; k < D1; k++) {
            float xr = iq[2 * (D1 * b + k)], xi = iq[2 * (D1 * b + k) + 1];
            float t1 = xr * wr[k], t2 = xr * wi[k];
            float ur = fmaf(-xi, wi[k], t1);
            float ui = fmaf(xi, wr[k], t2);
            for (int j = 0; j < 3; j++) {
                int t = 25 * j + 24 - k;
                if (t >= NCIC) continue;
                pr[j] = fmaf(g[t], ur, pr[j]);
                pi_[j] = fmaf(g[t], ui, pi_[j]);
            }
        }
        float Wc, Ws;
        spec_sincos((uint32_t)(((uint32_t)b + blk0) * fcw25), &Wc, &Ws);
        for (int j = 0; j < 3; j++) {
            float a1 = pr[j] * Wc, a2 = pr[j] * Ws;
            P[6 * b + 2 * j] = fmaf(-pi_[j], Ws, a1);
            P[6 * b + 2 * j + 1] = fmaf(pi_[j], Wc, a2);
        }
    }
    float *vr = (float *)calloc(nv ? nv : 1, sizeof(float)), *vi = (float *)calloc(nv ? nv : 1, sizeof(float));
    for (size_t m = 0; m < nv; m++) {
        float r = P[6 * m], i = P[6 * m + 1];
        float r1 = m >= 1 ? P[6 * (m - 1) + 2] : 0.0f, i1 = m >= 1 ? P[6 * (m - 1) + 3] : 0.0f;
        float r2 = m >= 2 ? P[6 * (m - 2) + 4] : 0.0f, i2 = m >= 2 ? P[6 * (m - 2) + 5] : 0.0f;
        vr[m] = (r + r1) + r2;
        vi[m] = (i + i1) + i2;
    }
    spec_stage2_demod(vr, vi, nq, h2, nh2, y_out, d_out);
    free(P); free(vr); free(vi);
}

/* ------------------------------------------------------------------ native 400 kS/s front end */
/* The reference's own operating point (grc/ampsbs.grc:263): x[m] @400 kS/s -> NCO -> lpf_taps /2 -> quadrature
 * demod.  This IS freq_xlating_fir_filter_ccc + quadrature_demod_cf; no extrapolation stage. */
void orc_rx_chain400_f64(const float *iq, size_t n, uint32_t fcw, const float *h2, int nh2, double *y_out, double *d_out) {
    size_t nq = n / 2;
    double *vr = (double *)malloc(sizeof(double) * (n ? n : 1)), *vi = (double *)malloc(sizeof(double) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) {
        uint32_t psi = (uint32_t)((uint64_t)i * fcw);
        double ang = 2.0 * M_PI * ((double)psi / 4294967296.0);
        double cr = cos(ang), ci = sin(ang);
        double xr = iq[2 * i], xi = iq[2 * i + 1];
        vr[i] = xr * cr - xi * ci;
        vi[i] = xr * ci + xi * cr;
    }
    double pr = 0, pi_ = 0;
    for (size_t q = 0; q < nq; q++) {
        double ar = 0, ai = 0;
        for (int k = 0; k < nh2; k++) {
            long idx = (long)(2 * q) - k;
            if (idx < 0) break;
            ar += (double)h2[k] * vr[idx];
            ai += (double)h2[k] * vi[idx];
        }
        if (y_out) { y_out[2 * q] = ar; y_out[2 * q + 1] = ai; }
        double zr = ar * pr + ai * pi_, zi = ai * pr - ar * pi_;
        if (d_out) d_out[q] = (zr == 0.0 && zi == 0.0) ? 0.0 : atan2(zi, zr);
        pr = ar; pi_ = ai;
    }
    free(vr); free(vi);
}

void orc_rx_chain400_f32(const float *iq, size_t n, uint32_t fcw, const float *h2, int nh2, float *y_out, float *d_out) {
    size_t nq = n / 2;
    float wr[D1], wi[D1];
    for (int k = 0; k < D1; k++) {
        uint32_t psi = (uint32_t)((uint32_t)k * fcw);
        double ang = 2.0 * M_PI * ((double)psi / 4294967296.0);
        wr[k] = (float)cos(ang); wi[k] = (float)sin(ang);
    }
    uint32_t fcw25 = (uint32_t)(25u * fcw);
    float *vr = (float *)malloc(sizeof(float) * (n ? n : 1)), *vi = (float *)malloc(sizeof(float) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) {
        /* v = (x (*) w[i mod 25]) (*) W(i div 25), each product as t = (x.re*w.re, x.re*w.im); fma(x.im, (-w.im, w.re), t) */
        int k = (int)(i % D1);
        float xr = iq[2 * i], xi = iq[2 * i + 1];
        float ur = fmaf(-xi, wi[k], xr * wr[k]);
        float ui = fmaf(xi, wr[k], xr * wi[k]);
        float Wc, Ws;
        spec_sincos((uint32_t)((uint32_t)(i / D1) * fcw25), &Wc, &Ws);
        vr[i] = fmaf(-ui, Ws, ur * Wc);
        vi[i] = fmaf(ui, Wc, ur * Ws);
    }
    spec_stage2_demod(vr, vi, nq, h2, nh2, y_out, d_out);
    free(vr); free(vi);
}

/* ------------------------------------------------------------------ detection on d */
#define OS 10   /* demod samples per half-symbol */
static int trig_match(const float *d, size_t i, const uint8_t *trig) {
    for (int k = 0; k < ORC_RECC_TRIGGER_LEN; k++) {
        int hard = d[i + (size_t)OS * k] >= 0.0f;
        if (hard != trig[k]) return 0;
    }
    return 1;
}
/* A "run" is a maximal set of adjacent sampling positions that all match the trigger 74/74 (exact
 * hard match, as the memmem of recc_impl.cc:118).  One burst per run whose FIRST position is not
 * inside an already captured burst; the sampling position is the soft-correlation peak of the run
 * (first maximum); the search resumes (74+3374)*10 positions after it. */
int orc_rx_detect(const float *d, size_t nd, orc_burst *out, int max) {
    uint8_t trig[ORC_RECC_TRIGGER_LEN];
    orc_recc_trigger(trig);
    const size_t span = (size_t)OS * (ORC_RECC_TRIGGER_LEN + ORC_RECC_CAPTURE_LEN - 1);  /* last needed offset */
    if (nd <= span) return 0;
    size_t limit = nd - span;          /* candidate positions i in [0, limit) have a complete capture */
    size_t resume = 0;
    int count = 0;
    size_t i = 0;
    while (i < limit && count < max) {
        if (!trig_match(d, i, trig) || (i > 0 && trig_match(d, i - 1, trig))) { i++; continue; }
        /* i starts a run */
        size_t best = i, j = i;
        float bestc = 0;
        int first = 1, open = 0;
        for (;; j++) {
            if (j >= limit) { open = 1; break; }          /* run reaches the end of the searchable range: not decided */
            if (!trig_match(d, j, trig)) break;
            float c = 0.0f;
            for (int k = 0; k < ORC_RECC_TRIGGER_LEN; k++) {
                float v = d[j + (size_t)OS * k];
                c = c + (trig[k] ? v : -v);
            }
            if (first || c > bestc) { best = j; bestc = c; first = 0; }
        }
        if (open) break;
        if (i >= resume) {
            orc_burst *b = &out[count++];
            b->d_index = best;
            b->corr = bestc;
            for (int s = 0; s < ORC_RECC_CAPTURE_LEN; s++)
                b->symbols[s] = d[best + (size_t)OS * (ORC_RECC_TRIGGER_LEN + s)] >= 0.0f ? 1 : 0;
            resume = best + (size_t)OS * (ORC_RECC_TRIGGER_LEN + ORC_RECC_CAPTURE_LEN);
        }
        i = j;
    }
    return count;
}

/* ------------------------------------------------------------------ forward (TX) chain, f64 */
/* What the reference does @400 kS/s (grc/ampsbs.grc): char_to_float (:1159-1252) ->
 * frequency_modulator_fc(2 pi 8000 / symrate) (:574-659) -> pfb.interpolator_ccf(4, firdes.low_pass(1,
 * 400e3, 10e3, tw)) (:2120-2229) -> [mute] -> multiply by e^{j 2 pi f n / fs} (:817-942) -> add (:1006-1056)
 * -> x 0.5 (:1355-1405).  The 10 MS/s extrapolation keeps all of that at the reference's 400 kS/s and
 * appends a CIC^3 x25 interpolator per carrier before the mixers:
 *
 *   s[i] in {+1,-1,0}  ->  S[i] = sum s  ->  fm[i] = (s[i] != 0) * exp(j 2 pi frac(S[i] * fcw_fm / 2^32))
 *   a[4i+j] = sum_k T[j+4k] fm[i-k]                         (pfb interpolator, zero history)        @400 kS/s
 *   b[5m+r]  = sum_{t=0..2} G5[r+5t] a[m-t],  G5 = 5 * boxcar5^3 / 125   (x5 CIC^3)                  @2 MS/s
 *   B[q]     = sum_c b_c[q] * exp(j 2 pi frac(5 q * fcw_c / 2^32))       (mixers, carriers summed)   @2 MS/s
 *   out[5q+r] = scale * sum_{t=0..2} G5[r+5t] B[q-t]                       (x5 CIC^3, shared)          @10 MS/s
 *
 * A symbol byte of 0 mutes that symbol (the reference's mute_xx, :1508-1601, gates the interpolator output;
 * gating its input differs only in the 0.8 ms filter transient).
 */
static void fwd_chain_core(const int8_t *const *sym, int ncarriers, size_t nsym, uint32_t fcw_fm,
                           const float *const *taps, const int *ntaps, const uint32_t *fcw_mix, double scale,
                           const double *const *extra400, double *out /* nsym*100 complex */) {
    const size_t nm = nsym * 4, nq = nsym * 20, nout = nsym * 100;
    double g5[15];
    {   /* 5 * (boxcar5 * boxcar5 * boxcar5) / 125, 13 taps */
        double a[9] = {0}, c[13] = {0};
        for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) a[i + j] += 1.0;
        for (int i = 0; i < 9; i++) for (int j = 0; j < 5; j++) c[i + j] += a[i];
        for (int i = 0; i < 15; i++) g5[i] = i < 13 ? 5.0 * c[i] / 125.0 : 0.0;
    }
    double *fr = (double *)malloc(sizeof(double) * nsym), *fi = (double *)malloc(sizeof(double) * nsym);
    double *ar = (double *)malloc(sizeof(double) * nm), *ai = (double *)malloc(sizeof(double) * nm);
    double *Br = (double *)calloc(nq, sizeof(double)), *Bi = (double *)calloc(nq, sizeof(double));
    for (int c = 0; c < ncarriers; c++) {
        int32_t S = 0;
        for (size_t i = 0; i < nsym; i++) {
            S += sym[c][i];
            uint32_t psi = (uint32_t)S * fcw_fm;
            double ang = 2.0 * M_PI * ((double)psi / 4294967296.0);
            double g = sym[c][i] != 0 ? 1.0 : 0.0;
            fr[i] = g * cos(ang); fi[i] = g * sin(ang);
        }
        for (size_t i = 0; i < nsym; i++)
            for (int j = 0; j < 4; j++) {
                double sr = 0, si = 0;
                for (int k = 0; j + 4 * k < ntaps[c]; k++) {
                    if (i < (size_t)k) break;
                    double t = taps[c][j + 4 * k];
                    sr += t * fr[i - k]; si += t * fi[i - k];
                }
                ar[4 * i + j] = sr; ai[4 * i + j] = si;
            }
        if (extra400 && extra400[c])                       /* voice leg sharing this carrier's mixer (add_xx before multiply_xx) */
            for (size_t m = 0; m < nm; m++) { ar[m] += extra400[c][2 * m]; ai[m] += extra400[c][2 * m + 1]; }
        for (size_t m = 0; m < nm; m++)
            for (int r = 0; r < 5; r++) {
                double br = 0, bi = 0;
                for (int t = 0; t < 3; t++) {
                    if (m < (size_t)t) continue;
                    br += g5[r + 5 * t] * ar[m - t]; bi += g5[r + 5 * t] * ai[m - t];
                }
                size_t q = 5 * m + (size_t)r;
                uint32_t psi = (uint32_t)((uint64_t)(5 * q) * fcw_mix[c]);
                double ang = 2.0 * M_PI * ((double)psi / 4294967296.0);
                double cr = cos(ang), ci = sin(ang);
                Br[q] += br * cr - bi * ci;
                Bi[q] += br * ci + bi * cr;
            }
    }
    for (size_t q = 0; q < nq; q++)
        for (int r = 0; r < 5; r++) {
            double orr = 0, oi = 0;
            for (int t = 0; t < 3; t++) {
                if (q < (size_t)t) continue;
                orr += g5[r + 5 * t] * Br[q - t]; oi += g5[r + 5 * t] * Bi[q - t];
            }
            out[2 * (5 * q + r)] = scale * orr;
            out[2 * (5 * q + r) + 1] = scale * oi;
        }
    (void)nout;
    free(fr); free(fi); free(ar); free(ai); free(Br); free(Bi);
}

void orc_fwd_chain_f64(const int8_t *const *sym, int ncarriers, size_t nsym, uint32_t fcw_fm,
                       const float *const *taps, const int *ntaps, const uint32_t *fcw_mix, double scale,
                       double *out /* nsym*100 complex */) {
    fwd_chain_core(sym, ncarriers, nsym, fcw_fm, taps, ntaps, fcw_mix, scale, NULL, out);
}

/* the same with a complex 400 kS/s baseband (4 nsym samples, e.g. orc_voice_tx_f64's output) added to carrier c's
 * interpolator output before its mixer; extra400[c] may be NULL */
void orc_fwd_chain_voice_f64(const int8_t *const *sym, int ncarriers, size_t nsym, uint32_t fcw_fm,
                             const float *const *taps, const int *ntaps, const uint32_t *fcw_mix, double scale,
                             const double *const *extra400, double *out) {
    fwd_chain_core(sym, ncarriers, nsym, fcw_fm, taps, ntaps, fcw_mix, scale, extra400, out);
}
