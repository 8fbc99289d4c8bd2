// design.h -- host-side filter / NCO design for libamps_b200 (plain C++, no CUDA).
// Restates the documented behaviour of gr::filter::firdes::low_pass (GNU Radio 3.7, not in the
// reference tree; parameters from grc/ampsbs.grc:138-184, 2172, 2227) and defines the 10 MS/s
// front end.
#pragma once
#include <cstdint>
#include <vector>

namespace amps {
enum Window { WIN_HAMMING = 0, WIN_HANN = 1, WIN_BLACKMAN = 2 };
std::vector<float> firdes_low_pass(double gain, double fs, double fc, double tw, Window win);
uint32_t nco_fcw(double center_freq, double samp_rate);          // phase step for a shift by -center_freq
void nco_block_table(uint32_t fcw, int n, float *re_im_pairs);   // e^{j 2 pi (k*fcw mod 2^32) / 2^32}, k < n
void cic3_taps(int decim, std::vector<float> &taps);              // boxcar^3 / decim^3
// 129 x 8 MMSE interpolator of clock_recovery_mm_ff (bandwidth 1/4, 6 significant digits like GNU Radio's header);
// row m, column k weights sample pos+k for mu = m/128
std::vector<float> mmse_interp_table();
// analog.fm_preemph (GNU Radio 3.7.10+): y[n] = b[0] x[n] + b[1] x[n-1] - a[1] y[n-1], 0 dB at DC; fh <= 0 -> 0.925 fs/2
void fm_preemph_taps(double fs, double tau, double fh, double b[2], double a[2]);
// its impulse response g[0..n) (the pole is at |p| ~ 0.79 for 16 kS/s / 75 us: 192 terms reach 1e-19)
std::vector<double> fm_preemph_impulse(double fs, double tau, double fh, int n);
// pfb.arb_resampler_ccf(25, taps, 8 arms) as a plain x25 polyphase filter: E[r * per + k], per = ceil(ntaps / 8)
std::vector<float> arb25_taps(const std::vector<float> &taps, int &per);
}  // namespace amps
