/*
 * mm_timing.c -- ORACLE (test infrastructure).  The serial symbol-timing tail of the reference's receive graph:
 *   digital.clock_recovery_mm_ff(omega=10, gain_omega=0.02296875, mu=0, gain_mu=0.05, omega_relative_limit=0.005)
 *   -> digital.binary_slicer_fb                                      (grc/ampsbs.grc:1751-1813, 1712-1750)
 * Both are GNU Radio 3.7 blocks whose source is NOT under /root/reference; this restates their documented loop
 * (SURVEY.md section 8a row A12 and Appendix B).  PARITY UNPINNED: GNU Radio's 8-tap x 129-phase MMSE interpolator
 * table (gnuradio-runtime interpolator_taps.h, a numerically optimised table) is not available here, so the table
 * is derived from the same criterion in closed form (orc_mmse_table); everything downstream of the table is exact
 * fp32 with one fixed operation order, which the CUDA kernel (csrc/rx_kernels.cu: rx_mm_kernel) repeats.
 *
 * Per output half-symbol, all fp32, no fused operations except the explicit fmaf chain of the interpolator:
 *   imu   = (int)rint(mu * 128)
 *   s     = sum_{k=0..7} T[imu][k] * d[pos + k]             (k ascending: one multiply, then 7 fmaf)
 *   mm    = sgn(last) * s - sgn(s) * last                   sgn(x) = x < 0 ? -1 : +1
 *   last  = s
 *   omega = omega + gain_omega * mm
 *   omega = omega_mid + clip(omega - omega_mid, omega_lim)  clip(x, c) = 0.5 * (|x + c| - |x - c|)
 *   mu    = (mu + omega) + gain_mu * mm
 *   f     = floor(mu); pos += (int)f; mu -= f               (f outside [1, 64] or NaN -- garbage input only -- is
 *                                                            forced to 1 / 64 and mu to 0 so that the loop advances)
 *   symbol = s >= 0 ? 1 : 0                                 (binary_slicer_fb)
 * The stream starts at pos = 0 with mu = 0, omega = 10, last = 0.  A symbol is produced whenever d[pos .. pos+7]
 * is available, so the output does not depend on how the input is chunked.
 */
#include "amps_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* 8-tap interpolator minimising  integral_{-B}^{B} | e^{j w mu'} - sum_k h_k e^{j w (k - 3)} |^2 dw / 2pi  with
 * B = 0.25 cycles/sample (the bandwidth GNU Radio's table was generated for, recalled): the normal equations
 *   sum_j r(i - j) h_j = r(i - 3 - mu),  r(t) = sin(2 pi B t) / (pi t), r(0) = 2B
 * solved in double by Gaussian elimination with partial pivoting, rounded to float.  T[m*8 + k] weights d[pos+k]
 * for mu = m/128; T[0] selects d[pos+3], T[128] selects d[pos+4]. */
static double r_sinc(double t) {
    const double B = 0.25;
    if (fabs(t) < 1e-12) return 2.0 * B;
    return sin(2.0 * M_PI * B * t) / (M_PI * t);
}

void orc_mmse_table(float *T) {
    for (int m = 0; m <= 128; m++) {
        const double mu = (double)m / 128.0;
        double a[8][9];
        for (int i = 0; i < 8; i++) {
            for (int j = 0; j < 8; j++) a[i][j] = r_sinc((double)(i - j));
            a[i][8] = r_sinc((double)i - 3.0 - mu);
        }
        for (int c = 0; c < 8; c++) {
            int piv = c;
            for (int r = c + 1; r < 8; r++) if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
            if (piv != c) for (int k = 0; k < 9; k++) { double t = a[c][k]; a[c][k] = a[piv][k]; a[piv][k] = t; }
            for (int r = c + 1; r < 8; r++) {
                const double f = a[r][c] / a[c][c];
                for (int k = c; k < 9; k++) a[r][k] -= f * a[c][k];
            }
        }
        double h[8];
        for (int i = 7; i >= 0; i--) {
            double s = a[i][8];
            for (int k = i + 1; k < 8; k++) s -= a[i][k] * h[k];
            h[i] = s / a[i][i];
        }
        /* GNU Radio's header prints its taps with 6 significant digits ("%.5e"); round the same way so that the table
         * agrees with it wherever its optimiser converged (e.g. mu = 1/2: -6.77751e-03 3.94578e-02 -1.42658e-01
         * 6.09836e-01 ...), and snap the 1e-14 residue of the exact rows (mu = 0, 1) to zero */
        for (int k = 0; k < 8; k++) {
            char txt[32];
            snprintf(txt, sizeof txt, "%.5e", fabs(h[k]) < 1e-9 ? 0.0 : h[k]);
            T[m * 8 + k] = (float)strtod(txt, NULL);
        }
    }
}

void orc_mm_init(orc_mm_state *st) {
    memset(st, 0, sizeof *st);
    st->mu = 0.0f; st->omega = 10.0f; st->last = 0.0f; st->pos = 0;
}

static float sgn(float x) { return x < 0.0f ? -1.0f : 1.0f; }

/* d = the demodulated stream from its first sample, total = samples available so far.  Returns the number of
 * half-symbols written to sym (at most cap); call again with a larger total to continue. */
size_t orc_mm_process(orc_mm_state *st, const float *d, uint64_t total, const float *T, uint8_t *sym, size_t cap) {
    const float omega_mid = 10.0f, gain_omega = 0.02296875f, gain_mu = 0.05f;
    const float omega_lim = omega_mid * 0.005f;
    size_t n = 0;
    float mu = st->mu, omega = st->omega, last = st->last;
    uint64_t pos = st->pos;
    while (n < cap && pos + 8 <= total) {
        const int imu = (int)lrintf(mu * 128.0f);
        const float *t = T + 8 * imu, *x = d + pos;
        float s = t[0] * x[0];
        for (int k = 1; k < 8; k++) s = fmaf(t[k], x[k], s);
        const float mm = sgn(last) * s - sgn(s) * last;
        last = s;
        omega = omega + gain_omega * mm;
        const float dev = omega - omega_mid;
        omega = omega_mid + 0.5f * (fabsf(dev + omega_lim) - fabsf(dev - omega_lim));
        mu = (mu + omega) + gain_mu * mm;
        float f = floorf(mu);
        if (f >= 1.0f && f <= 64.0f) mu = mu - f;
        else { f = (f > 64.0f) ? 64.0f : 1.0f; mu = 0.0f; }
        pos += (uint64_t)(int)f;
        sym[n++] = (uint8_t)(s >= 0.0f ? 1 : 0);
    }
    st->mu = mu; st->omega = omega; st->last = last; st->pos = pos;
    return n;
}
