/*
 * voice_tx.c -- ORACLE (test infrastructure).  The voice leg of the reference's forward graph:
 *   wavfile_source(16 kS/s) + sig_source_f(16000, COS, 6000, 0.05)  [SAT]     grc/ampsbs.grc:943-1005, 1057-1107
 *   -> analog.nbfm_tx(audio_rate=16000, quad_rate=16000, tau=75e-6, max_dev=8e3, fh=-1)          :715-773
 *   -> [mute_xx (audio_mute)] -> pfb.arb_resampler_ccf(25, voice_lpf_taps, nfilts=8)             :1994-2119
 *   -> added to the FVC leg before the +60 kHz mixer / alone into the +90 kHz mixer               :4494-4500, 4632-4638
 * All of these are GNU Radio 3.7 blocks whose source is not under /root/reference: PARITY UNPINNED.  Restated from
 * their documented behaviour, evaluated in float64:
 *   nbfm_tx   = interp_fir (factor 1: identity) -> fm_preemph(fs, tau, fh) -> frequency_modulator_fc(2 pi max_dev / fs)
 *   fm_preemph (3.7.10+, the version with the `fh` parameter the flowgraph file carries): one-pole/one-zero IIR by
 *               bilinear transform with pre-warping, fh <= 0 -> 0.925 fs/2:
 *                 w_cla = 2 fs tan(1/(2 tau fs)), w_cha = 2 fs tan(pi fh / fs), kl = -w_cla/(2 fs), kh = -w_cha/(2 fs)
 *                 z1 = (1+kl)/(1-kl), p1 = (1+kh)/(1-kh), b0 = (1-kl)/(1-kh), g = |1-p1| / (b0 |1-z1|)  (0 dB at DC);
 *                 y[n] = g b0 x[n] - g b0 z1 x[n-1] + p1 y[n-1]
 *   frequency_modulator_fc: phi += k x, out = exp(j phi); here phi is kept in cycles modulo 1 (no precision loss)
 *   pfb_arb_resampler_ccf(rate 25, taps, 8 arms): polyphase interpolation by 8 with linear interpolation between
 *               adjacent arms (taps and first-difference taps); with rate 25 the arm position of output 25 i + r is
 *               8 r / 25 exactly, so  out[25 i + r] = sum_k E_r[k] in[i - k],
 *               E_r[k] = h[j + 8k] + a (h[j + 1 + 8k] - h[j + 8k]),  j = floor(8 r / 25),  a = frac(8 r / 25),
 *               h = taps zero-padded; zero history.  (GNU Radio advances the fractional position in float; the
 *               closed form is the drift-free statement of the same thing.)
 * A muted audio sample is a zero at the resampler input (mute_xx sits between nbfm_tx and the resampler).
 */
#include "amps_oracle.h"
#include <math.h>
#include <stdlib.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

void orc_fm_preemph_taps(double fs, double tau, double fh, double b[2], double a[2]) {
    if (fh <= 0.0 || fh >= fs / 2.0) fh = 0.925 * fs / 2.0;
    const double w_cl = 1.0 / tau, w_ch = 2.0 * M_PI * fh;
    const double w_cla = 2.0 * fs * tan(w_cl / (2.0 * fs)), w_cha = 2.0 * fs * tan(w_ch / (2.0 * fs));
    const double kl = -w_cla / (2.0 * fs), kh = -w_cha / (2.0 * fs);
    const double z1 = (1.0 + kl) / (1.0 - kl), p1 = (1.0 + kh) / (1.0 - kh), b0 = (1.0 - kl) / (1.0 - kh);
    /* H(z = -1) = 1 as designed; GNU Radio rescales so that the gain at DC is 0 dB instead */
    const double g = fabs(1.0 - p1) / (b0 * fabs(1.0 - z1));
    b[0] = g * b0; b[1] = -g * b0 * z1;
    a[0] = 1.0; a[1] = -p1;
}

/* effective x25 polyphase taps of the arb resampler: E[r * per + k], per = ceil(ntaps / 8); returns per */
int orc_arb25_taps(const float *taps, int ntaps, double *E /* 25 * ceil(ntaps/8) */) {
    const int per = (ntaps + 7) / 8;
    for (int r = 0; r < 25; r++) {
        const int j = (8 * r) / 25;
        const double a = (double)((8 * r) % 25) / 25.0;
        for (int k = 0; k < per; k++) {
            const int n0 = j + 8 * k, n1 = n0 + 1;
            const double h0 = n0 < ntaps ? (double)taps[n0] : 0.0, h1 = n1 < ntaps ? (double)taps[n1] : 0.0;
            E[r * per + k] = h0 + a * (h1 - h0);
        }
    }
    return per;
}

/* audio[n_a] (16 kS/s, float) -> complex 400 kS/s baseband out[2 * 25 * n_a].  mute: per audio sample (NULL = never).
 * sat_amp: amplitude of the 6 kHz supervisory tone added to the audio (0.05 in the flowgraph). */
void orc_voice_tx_f64(const float *audio, size_t n_a, double sat_amp, const uint8_t *mute, const float *taps, int ntaps,
                      double *out) {
    const double fs = 16000.0;
    double b[2], a[2];
    orc_fm_preemph_taps(fs, 75e-6, -1.0, b, a);
    const double k_cycles = 8000.0 / fs;                       /* 2 pi max_dev / fs radians = max_dev / fs cycles */
    double *pr = (double *)malloc(sizeof(double) * (n_a + 1)), *pi = (double *)malloc(sizeof(double) * (n_a + 1));
    double xprev = 0.0, yprev = 0.0, phi = 0.0;
    for (size_t n = 0; n < n_a; n++) {
        const double x = (double)audio[n] + sat_amp * cos(2.0 * M_PI * (double)((3 * n) % 8) / 8.0);   /* 6000/16000 = 3/8 */
        const double y = b[0] * x + b[1] * xprev - a[1] * yprev;
        xprev = x; yprev = y;
        phi += k_cycles * y;
        phi -= floor(phi);
        const double g = (mute && mute[n]) ? 0.0 : 1.0;
        pr[n] = g * cos(2.0 * M_PI * phi); pi[n] = g * sin(2.0 * M_PI * phi);
    }
    const int per = (ntaps + 7) / 8;
    double *E = (double *)malloc(sizeof(double) * 25 * (size_t)per);
    orc_arb25_taps(taps, ntaps, E);
    for (size_t i = 0; i < n_a; i++)
        for (int r = 0; r < 25; r++) {
            double sr = 0.0, si = 0.0;
            for (int k = 0; k < per && (size_t)k <= i; k++) { sr += E[r * per + k] * pr[i - k]; si += E[r * per + k] * pi[i - k]; }
            out[2 * (25 * i + (size_t)r)] = sr; out[2 * (25 * i + (size_t)r) + 1] = si;
        }
    free(pr); free(pi); free(E);
}
