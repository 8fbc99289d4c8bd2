"""An INDEPENDENT float64 implementation of the stock GNU Radio blocks on the hot path, written only with numpy / scipy.signal
library routines (upfirdn, cumsum, angle, exp) from the blocks' documented equations -- no code shared with
oracle/ or with the CUDA path, and a different structure on purpose:

  * freq_xlating_fir_filter_ccc is done the way GNU Radio does it -- band-pass taps c[k] = h[k] e^{j k theta}, decimating FIR,
    then the rotator e^{-j theta D m} -- whereas the oracle and the kernels rotate first and low-pass with the real taps;
  * quadrature_demod_cf is np.angle(y[m] conj(y[m-1]));
  * frequency_modulator_fc is exp(j cumsum(sens * x)); pfb.interpolator_ccf(R, taps) is upfirdn(taps, x, up=R);
  * the 10 MS/s extension stages (boxcar^3 decimator / interpolators) are upfirdn with np.convolve'd boxcars.

Only the tap DESIGN (firdes.low_pass) is taken from the product/oracle: it is pinned separately to GNU Radio's own QA vector
(tests/golden/kat_gnuradio_firdes.json).  Test infrastructure."""
import numpy as np
from scipy import signal


def nco_theta(center_freq, samp_rate):
    """2 pi fc / fs with fc / fs rounded to the product's 32-bit frequency word (|error| < 1.2e-10 cycles per sample = 47 uHz at
    400 kS/s: without it the two chains drift apart by 5e-5 rad over a 0.2 s buffer, which is all the comparison would show)."""
    fcw = int(np.round((-center_freq / samp_rate) % 1.0 * 2.0 ** 32)) % (1 << 32)
    return -2.0 * np.pi * fcw / 2.0 ** 32


def freq_xlating_fir_filter_ccc(x, taps, center_freq, samp_rate, decim):
    """GNU Radio's definition: y[m] = e^{-j theta D m} sum_k h[k] e^{j k theta} x[m D - k], theta = 2 pi fc / fs, zero history."""
    theta = nco_theta(center_freq, samp_rate)
    k = np.arange(len(taps))
    bp = np.asarray(taps, np.float64) * np.exp(1j * theta * k)
    y = signal.upfirdn(bp, np.asarray(x, np.complex128), up=1, down=decim)[: len(x) // decim]
    m = np.arange(len(y))
    return y * np.exp(-1j * theta * decim * m)


def quadrature_demod_cf(y, gain=1.0):
    prev = np.concatenate([[0.0 + 0.0j], y[:-1]])
    return gain * np.angle(y * np.conj(prev))


def boxcar3(n):
    b = np.ones(n)
    return np.convolve(np.convolve(b, b), b)


def rx_chain_400k(x, taps, center_freq=-160e3):
    """The reference's own receive graph at its own rate (grc/ampsbs.grc:1814-1872, 774-816)."""
    y = freq_xlating_fir_filter_ccc(x, taps, center_freq, 400e3, 2)
    return y, quadrature_demod_cf(y)


def rx_chain_10m(x, taps, center_freq=-160e3):
    """10 MS/s: translate, boxcar^3 / 25^3 decimating by 25 (output m covers samples ... 25 m + 24), then the reference's
    filter /2 at 400 kS/s and the demod."""
    n = np.arange(len(x))
    # the product's NCO is a 32-bit phase accumulator: the same frequency word, or the two drift apart by 1e-10 cycles/sample
    fcw = int(np.round((-center_freq / 10e6) % 1.0 * 2.0 ** 32))
    u = np.asarray(x, np.complex128) * np.exp(2j * np.pi * ((n * fcw) % (1 << 32)) / 2.0 ** 32)
    g = boxcar3(25) / 25.0 ** 3
    full = signal.upfirdn(g, u)                       # full[i] = sum_t g[t] u[i - t]
    v = full[24::25][: len(x) // 25]                  # v[m] = sum_t g[t] u[25 m + 24 - t]
    y = signal.upfirdn(np.asarray(taps, np.float64), v, up=1, down=2)[: len(v) // 2]
    return y, quadrature_demod_cf(y)


def frequency_modulator_fc(x, sensitivity):
    return np.exp(1j * np.cumsum(sensitivity * np.asarray(x, np.float64)))


def fwd_chain_10m(syms, taps, carrier_freq, scale=0.5, max_dev=8000.0, symrate=100e3):
    """char_to_float -> frequency_modulator_fc(2 pi max_dev / symrate) -> pfb.interpolator_ccf(4, taps) [-> mute] -> (x5 boxcar^3
    interpolator) -> mixer -> add -> (x5 boxcar^3 interpolator) -> x scale.  A symbol byte of 0 mutes the modulator output."""
    g5 = 5.0 * boxcar3(5) / 125.0
    total = None
    for s, t, fc in zip(syms, taps, carrier_freq):
        f = np.asarray(s).astype(np.int8).astype(np.float64)           # +1, -1 (0xFF), 0
        fm = frequency_modulator_fc(f, 2.0 * np.pi * max_dev / symrate) * (f != 0)
        a = signal.upfirdn(np.asarray(t, np.float64), fm, up=4)[: 4 * len(f)]
        b = signal.upfirdn(g5, a, up=5)[: 20 * len(f)]
        q = np.arange(len(b))
        fcw = int(np.round((fc / 10e6) % 1.0 * 2.0 ** 32))
        b = b * np.exp(2j * np.pi * ((5 * q * fcw) % (1 << 32)) / 2.0 ** 32)
        total = b if total is None else total + b
    return scale * signal.upfirdn(g5, total, up=5)[: 100 * len(syms[0])]
