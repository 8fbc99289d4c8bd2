#include "design.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <utility>

namespace amps {

static const double kPi = 3.14159265358979323846;

std::vector<float> firdes_low_pass(double gain, double fs, double fc, double tw, Window win) {
    static const double atten[3] = {53.0, 44.0, 74.0};
    int ntaps = (int)(atten[win] * fs / (22.0 * tw));
    if ((ntaps & 1) == 0) ntaps++;
    const int M = (ntaps - 1) / 2;
    const double fwT0 = 2.0 * kPi * fc / fs;
    std::vector<float> taps((size_t)ntaps);
    for (int n = -M; n <= M; n++) {
        const int k = n + M;
        double wv;
        if (win == WIN_HAMMING) wv = 0.54 - 0.46 * std::cos(2.0 * kPi * k / (ntaps - 1));
        else if (win == WIN_HANN) wv = 0.5 - 0.5 * std::cos(2.0 * kPi * k / (ntaps - 1));
        else wv = 0.42 - 0.5 * std::cos(2.0 * kPi * k / (ntaps - 1)) + 0.08 * std::cos(4.0 * kPi * k / (ntaps - 1));
        const float wf = (float)wv;                      // the window is a float vector in GNU Radio
        if (n == 0) taps[k] = (float)(fwT0 / kPi * wf);
        else taps[k] = (float)(std::sin(n * fwT0) / (n * kPi) * wf);
    }
    double fmax = taps[M];
    for (int n = 1; n <= M; n++) fmax += 2.0 * taps[n + M];
    const double g = gain / fmax;
    for (int i = 0; i < ntaps; i++) taps[i] = (float)(taps[i] * g);
    return taps;
}

uint32_t nco_fcw(double center_freq, double samp_rate) {
    double turns = -center_freq / samp_rate;
    turns -= std::floor(turns);
    return (uint32_t)(uint64_t)std::llround(turns * 4294967296.0);
}

void nco_block_table(uint32_t fcw, int n, float *out) {
    for (int k = 0; k < n; k++) {
        const uint32_t psi = (uint32_t)((uint32_t)k * fcw);
        const double ang = 2.0 * kPi * ((double)psi / 4294967296.0);
        out[2 * k] = (float)std::cos(ang);
        out[2 * k + 1] = (float)std::sin(ang);
    }
}

void cic3_taps(int decim, std::vector<float> &taps) {
    std::vector<double> a((size_t)(2 * decim - 1), 0.0), c((size_t)(3 * decim - 2), 0.0);
    for (int i = 0; i < decim; i++) for (int j = 0; j < decim; j++) a[(size_t)(i + j)] += 1.0;
    for (int i = 0; i < 2 * decim - 1; i++) for (int j = 0; j < decim; j++) c[(size_t)(i + j)] += a[(size_t)i];
    const double norm = (double)decim * decim * decim;
    taps.resize(c.size());
    for (size_t i = 0; i < c.size(); i++) taps[i] = (float)(c[i] / norm);
}

// Normal equations of the band-limited (|f| < 1/4) least-squares interpolator:
//   sum_j r(i - j) h_j = r(i - 3 - mu),   r(t) = sin(pi t / 2) / (pi t),  r(0) = 1/2
// Gauss-Jordan in double; the result is then cut to the 6 significant digits GNU Radio's table is printed with.
static double band_autocorr(double t) { return std::fabs(t) < 1e-12 ? 0.5 : std::sin(0.5 * kPi * t) / (kPi * t); }

std::vector<float> mmse_interp_table() {
    const int N = 8, rows = 129;
    std::vector<float> table((size_t)rows * N);
    for (int m = 0; m < rows; m++) {
        const double mu = m / 128.0;
        double a[N][N + 1];
        for (int i = 0; i < N; i++) {
            for (int j = 0; j < N; j++) a[i][j] = band_autocorr(i - j);
            a[i][N] = band_autocorr(i - 3.0 - mu);
        }
        for (int c = 0; c < N; c++) {
            int best = c;
            for (int r = c + 1; r < N; r++)
                if (std::fabs(a[r][c]) > std::fabs(a[best][c])) best = r;
            for (int k = 0; k <= N; k++) std::swap(a[c][k], a[best][k]);
            const double inv = 1.0 / a[c][c];
            for (int k = c; k <= N; k++) a[c][k] *= inv;
            for (int r = 0; r < N; r++) {
                if (r == c) continue;
                const double f = a[r][c];
                for (int k = c; k <= N; k++) a[r][k] -= f * a[c][k];
            }
        }
        for (int k = 0; k < N; k++) {
            char txt[32];
            std::snprintf(txt, sizeof txt, "%.5e", std::fabs(a[k][N]) < 1e-9 ? 0.0 : a[k][N]);
            table[(size_t)m * N + k] = (float)std::strtod(txt, nullptr);
        }
    }
    return table;
}

void fm_preemph_taps(double fs, double tau, double fh, double b[2], double a[2]) {
    if (fh <= 0.0 || fh >= 0.5 * fs) fh = 0.925 * 0.5 * fs;
    // corner frequencies 1/tau and 2 pi fh, pre-warped, through the bilinear transform
    const double kl = -std::tan(1.0 / (2.0 * tau * fs)), kh = -std::tan(kPi * fh / fs);
    const double zero = (1.0 + kl) / (1.0 - kl), pole = (1.0 + kh) / (1.0 - kh), b0 = (1.0 - kl) / (1.0 - kh);
    const double gain = std::fabs(1.0 - pole) / (b0 * std::fabs(1.0 - zero));       // 0 dB at DC
    b[0] = gain * b0; b[1] = -gain * b0 * zero;
    a[0] = 1.0; a[1] = -pole;
}

std::vector<double> fm_preemph_impulse(double fs, double tau, double fh, int n) {
    double b[2], a[2];
    fm_preemph_taps(fs, tau, fh, b, a);
    std::vector<double> g((size_t)n, 0.0);
    double y = 0.0;
    for (int k = 0; k < n; k++) {
        y = (k == 0 ? b[0] : (k == 1 ? b[1] : 0.0)) - a[1] * y;
        g[(size_t)k] = y;
    }
    return g;
}

std::vector<float> arb25_taps(const std::vector<float> &taps, int &per) {
    const int n = (int)taps.size();
    per = (n + 7) / 8;
    std::vector<float> E((size_t)25 * per);
    auto h = [&](int i) { return i < n ? (double)taps[(size_t)i] : 0.0; };
    for (int r = 0; r < 25; r++) {
        const int arm = 8 * r / 25;
        const double frac = (8 * r % 25) / 25.0;                  // position between arm and arm + 1
        for (int k = 0; k < per; k++) E[(size_t)r * per + k] = (float)(h(arm + 8 * k) + frac * (h(arm + 8 * k + 1) - h(arm + 8 * k)));
    }
    return E;
}

}  // namespace amps
