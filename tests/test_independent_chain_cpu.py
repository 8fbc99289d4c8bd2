"""Float rows A5-A12 (stock GNU Radio blocks) held against something that is NOT ours:
  (1) an independent numpy/scipy.signal implementation of the blocks' documented equations (tests/scipy_chain.py), written the
      way GNU Radio structures them (band-pass taps + rotator for the frequency-translating FIR, upfirdn for the polyphase
      interpolator, np.angle for the demod) -- against the float64 oracle chains the GPU is held to (<= 1e-6 RMS);
  (2) known answers of GNU Radio's own QA suite for quadrature_demod_cf and for the M&M interpolator
      (tests/golden/kat_gnuradio_blocks.json; firdes has its own file)."""
import json
import os

import numpy as np

from gr_amps_b200 import synth
from tests import scipy_chain as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_gnuradio_blocks.json")


def rms(a):
    return float(np.sqrt(np.mean(np.abs(a) ** 2)))


def wrap(a):
    return np.angle(np.exp(1j * a))


def test_native_rate_chain_is_gnuradios_freq_xlating_fir_plus_quadrature_demod(oracle):
    """400 kS/s: the oracle rotates first and low-passes with the real taps; GNU Radio filters with band-pass taps and
    rotates afterwards.  Same numbers to 1e-9, on every config-4 carrier offset."""
    taps = oracle.lpf_taps()
    hs = synth.manchester(synth.recc_message_bits(synth.origination_words()))
    for g in (0, 3, 7):
        fc = -160e3 + 30e3 * g
        x = synth.fm_burst(hs, 55 * 1536, 800, samp_rate=400e3, center=fc, snr_db=20.0, seed=5 + g)
        y64, d64 = oracle.rx_chain400_f64(x, center=fc)
        ys, ds = S.rx_chain_400k(x, taps, fc)
        assert rms(y64 - ys) <= 1e-9 * rms(ys) and np.max(np.abs(y64 - ys)) <= 1e-9
        strong = np.abs(ys) > 1e-3                      # where there is a carrier the demod is well conditioned
        strong[1:] &= strong[:-1]
        assert np.max(np.abs(wrap(d64 - ds))[strong]) <= 1e-8
        # ... and the kernel-spec fp32 flavour (what the GPU reproduces bit for bit) stays within fp32 rounding of it
        y32, d32 = oracle.rx_chain400_f32(x, center=fc)
        assert rms(y32.astype(np.complex128) - ys) <= 1e-6
        assert np.max(np.abs(wrap(d32 - ds))[strong]) <= 2e-4
        sure = strong & (np.abs(ds) > 1e-3)
        assert np.array_equal((d32 >= 0)[sure], (ds >= 0)[sure])


def test_10ms_chain_against_scipy(oracle):
    taps = oracle.lpf_taps()
    for g in (0, 5):
        fc = -160e3 + 30e3 * g
        x, _, _ = synth.config2_period(n_total=55 * 38400, snr_db=20.0, seed=9 + g, center=fc)
        x = x[:12 * 38400]
        y64, d64 = oracle.rx_chain_f64(x, center=fc)
        ys, ds = S.rx_chain_10m(x, taps, fc)
        assert rms(y64 - ys) <= 1e-12 * max(rms(ys), 1.0)
        y32, d32 = oracle.rx_chain_f32(x, center=fc)
        assert rms(y32.astype(np.complex128) - ys) <= 1e-6      # the north-star tolerance, against an implementation that is not ours


def test_forward_chain_against_scipy(oracle):
    """char_to_float -> frequency_modulator_fc -> pfb.interpolator_ccf -> mixers -> add -> x0.5 as library calls."""
    nsym = 6000
    focc = oracle.Focc(100000, False).generate(nsym, chunk=1 << 20)
    v = oracle.Fvc(100000)
    v.push_words(oracle.word("orc_fvc_word1_general", 1, 0, 0, 1))
    out = bytearray()
    while len(out) < nsym:
        _, b, _ = v.work(min(8192, nsym - len(out)))
        out += b.tobytes()
    fvc = np.frombuffer(bytes(out), np.uint8).copy()
    fvc[2000:2600] = 0                                   # a muted stretch (mute_xx)
    syms = [focc, fvc, fvc.copy()]
    cf, tw = (0.0, 60e3, 90e3), (5e3, 3e3, 3e3)
    taps = [oracle.firdes_low_pass(1.0, 400e3, 10e3, t, 0) for t in tw]
    ref = oracle.fwd_chain_f64(syms, carrier_freq=cf, lpf_transition=tw, scale=0.5)
    ys = S.fwd_chain_10m(syms, taps, cf, scale=0.5)
    assert ref.shape == ys.shape and np.max(np.abs(ref - ys)) <= 2e-9
    # the modulator on its own: phase = running sum of sensitivity * symbol
    fm = S.frequency_modulator_fc(np.array([1, 1, -1, -1, -1, 1], float), 2 * np.pi * 8000 / 100e3)
    assert np.allclose(np.angle(fm), 0.5026548245743669 * np.array([1, 2, 1, 0, -1, 0]), atol=1e-12)


def test_gnuradio_qa_quadrature_demod(oracle):
    k = json.load(open(GOLD))["quadrature_demod_cf"]
    assert k["places"] == 5
    i = np.arange(200)
    x = np.exp(2j * np.pi * 1000.0 * i / 8000.0).astype(np.complex64)
    gain = 1.0 / (np.pi / 4)
    expected = np.array([0.0] + 199 * [1.0])
    import ctypes as C
    L = oracle.lib()
    L.orc_quad_demod.argtypes = [oracle.f32p, C.c_size_t, oracle.f32p, oracle.f64p]
    iq = np.ascontiguousarray(x.view(np.float32))
    d32, d64 = np.zeros(200, np.float32), np.zeros(200, np.float64)
    L.orc_quad_demod(oracle.ptr(iq, oracle.f32p), 200, oracle.ptr(d32, oracle.f32p), oracle.ptr(d64, oracle.f64p))
    for got in (gain * d64, gain * d32.astype(np.float64), S.quadrature_demod_cf(x.astype(np.complex128), gain)):
        assert np.max(np.abs(got - expected)) < 0.5e-5      # assertFloatTuplesAlmostEqual(..., 5)


def test_gnuradio_qa_clock_recovery_dc_gain(oracle):
    """GNU Radio's QA expects 0.99972 out of clock_recovery_mm for a constant 1.0 at mu = 0.5 ("doesn't quite get to 1.0"): the
    DC gain of the MMSE interpolator's middle row.  The table derived in closed form (csrc/design.cc, oracle/mm_timing.c)
    reproduces it -- a third anchor besides the two rows of interpolator_taps.h on record."""
    k = json.load(open(GOLD))["clock_recovery_mm_ff_dc"]
    t = oracle.mmse_table()
    assert round(float(t[64].astype(np.float64).sum()), k["places"]) == k["expected_tail"]
    assert np.array_equal(t[64], t[64][::-1])            # mu = 1/2: symmetric
