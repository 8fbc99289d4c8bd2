"""GPU parity on the eight BASELINE config-4 carriers: carrier g sits at -160 kHz + 30 kHz * g (the reference's
rx_offset, grc/ampsbs.grc:212-238, stepped along the 30 kHz AMPS raster), seed 0xA3B5 + g, its own MIN.  A different
center_freq changes the NCO control word, hence the in-block phasor tables w/wj and the block phasors W(b) of the front
kernel -- every one of them is held against the oracle here, bit for bit, through the C ABI: demodulated stream, burst
position, soft correlation, the 3374-byte blob, and the decode struct; the complex baseband within 1e-6 RMS of float64.
The forward path gets the same treatment with non-default carrier_freq (the +60 / +90 kHz mixers of
grc/ampsbs.grc:841,904 moved along the raster)."""
import numpy as np
import pytest

from gr_amps_b200 import multi, synth
from tests.helpers import bits_equal_f32, words_equal

pytestmark = pytest.mark.gpu

PASS = 38400
N1 = 55 * PASS


@pytest.fixture(scope="module")
def capi():
    from gr_amps_b200 import capi as c
    c.lib()
    return c


def check_carrier(capi, oracle, g, snr, sc16=False):
    car = multi.carrier(g)
    assert car.center_freq == -160e3 + 30e3 * g and car.seed == 0xA3B5 + g
    x, hs, _ = synth.config2_period(n_total=N1, snr_db=snr, seed=car.seed, center=car.center_freq, min10=car.min10)
    if sc16:
        scale = 1.0 / 8192.0
        q = np.clip(np.round(x.view(np.float32) / scale), -32768, 32767).astype(np.int16)
        x = (q.astype(np.float32) * np.float32(scale)).view(np.complex64)      # what the kernel converts to
        rx = capi.ReccIq(max_samples=N1, center_freq=car.center_freq, sc16=True, sc16_scale=scale, dump_baseband=True)
        bursts = rx.work(q)
    else:
        rx = capi.ReccIq(max_samples=N1, center_freq=car.center_freq, dump_baseband=True)
        bursts = rx.work(x)
    nd = N1 // 50
    y_orc, d_orc = oracle.rx_chain_f32(x, center=car.center_freq)
    assert bits_equal_f32(rx.read_demod(0, nd), d_orc), "demod stream differs on carrier %d" % g
    y_gpu = rx.read_baseband(0, nd)
    assert bits_equal_f32(y_gpu.view(np.float32), y_orc.view(np.float32))
    y64, _ = oracle.rx_chain_f64(x, center=car.center_freq)
    rms = np.sqrt(np.mean(np.abs(y_gpu.astype(np.complex128) - y64) ** 2))
    assert rms <= 1e-6, rms
    ob = oracle.rx_detect(d_orc)
    assert len(bursts) == len(ob) == 1
    b = bursts[0]
    assert b.demod_index == ob[0][0] and np.float32(b.corr) == np.float32(ob[0][1])
    assert np.array_equal(b.symbols_np(), ob[0][2]), "blob differs on carrier %d" % g
    assert np.array_equal(b.symbols_np(), hs[82:82 + 3374])
    assert words_equal(b.decoded, oracle.recc_decode(ob[0][2])) == []
    assert b.decoded.min == car.min10.encode() and list(b.decoded.valid) == [1] * 7 and b.decoded.kind == 4
    rx.close()


@pytest.mark.parametrize("g", range(8))
@pytest.mark.parametrize("snr", [None, 30.0, 15.0])
def test_config4_carrier_bit_exact(capi, oracle, g, snr):
    check_carrier(capi, oracle, g, snr)


@pytest.mark.parametrize("g", [1, 5, 7])
def test_config4_carrier_sc16(capi, oracle, g):
    check_carrier(capi, oracle, g, 20.0, sc16=True)


def test_carrier_rejects_its_neighbour(capi, oracle):
    """A handle tuned to carrier 3 must not decode carrier 4's burst (30 kHz away, > 70 dB down after the channel
    filter) -- and the oracle agrees on the whole demodulated stream."""
    c3, c4 = multi.carrier(3), multi.carrier(4)
    x, _, _ = synth.config2_period(n_total=N1, snr_db=30.0, seed=c4.seed, center=c4.center_freq, min10=c4.min10)
    rx = capi.ReccIq(max_samples=N1, center_freq=c3.center_freq)
    assert rx.work(x) == []
    _, d = oracle.rx_chain_f32(x, center=c3.center_freq)
    assert bits_equal_f32(rx.read_demod(0, N1 // 50), d) and oracle.rx_detect(d) == []
    rx.close()


def test_two_carriers_in_one_band(capi, oracle):
    """Carriers 0 and 6 transmitted into the SAME wideband buffer (the replicated-buffer case of SURVEY 8e): each handle
    recovers its own burst from the sum."""
    c0, c6 = multi.carrier(0), multi.carrier(6)
    xa, _, _ = synth.config2_period(n_total=N1, snr_db=None, center=c0.center_freq, min10=c0.min10)
    xb, _, _ = synth.config2_period(n_total=N1, snr_db=25.0, seed=c6.seed, center=c6.center_freq, min10=c6.min10, lead=31000)
    x = (xa + xb).astype(np.complex64)
    for c in (c0, c6):
        rx = capi.ReccIq(max_samples=N1, center_freq=c.center_freq)
        b = rx.work(x)
        _, d = oracle.rx_chain_f32(x, center=c.center_freq)
        ob = oracle.rx_detect(d)
        assert bits_equal_f32(rx.read_demod(0, N1 // 50), d)
        assert len(b) == len(ob) == 1 and b[0].demod_index == ob[0][0] and np.array_equal(b[0].symbols_np(), ob[0][2])
        assert b[0].decoded.min == c.min10.encode() and list(b[0].decoded.valid) == [1] * 7
        rx.close()


# ------------------------------------------------------------------ forward path along the raster
def fwd_symbols(oracle, nsym):
    focc = oracle.Focc(100000, False).generate(nsym, chunk=1 << 20)
    alert = oracle.word("orc_fvc_word1_general", 1, 0, 0, 1)
    v = oracle.Fvc(100000)
    v.push_words(alert)
    out = bytearray()
    while len(out) < nsym:
        _, b, _ = v.work(min(8192, nsym - len(out)))
        out += b.tobytes()
    fvc = np.frombuffer(bytes(out), np.uint8)
    return [focc, fvc, fvc.copy()]


@pytest.mark.parametrize("g", [1, 4, 7])
def test_forward_path_on_shifted_carriers(capi, oracle, g):
    """Config 4's forward half: FOCC + two FVC legs with the whole group moved by 30 kHz * g (and below zero for the
    FOCC leg), both input kinds, against the float64 chain (<= 1e-6 RMS) and against each other."""
    nbits = 2100
    syms = fwd_symbols(oracle, nbits * 10)
    cf = (30e3 * g - 120e3, 60e3 + 30e3 * g - 120e3, 90e3 + 30e3 * g - 120e3)
    tw = (5e3, 3e3, 3e3)
    ref = oracle.fwd_chain_f64(syms, carrier_freq=cf, lpf_transition=tw, scale=0.5)
    y = capi.Fwd(max_samples=nbits * 1000, carrier_freq=cf, lpf_transition=tw).work(syms)
    e = float(np.sqrt(np.mean(np.abs(y.astype(np.complex128) - ref) ** 2)))
    assert e <= 1e-6, e
    bits = [(np.asarray(s).reshape(-1, 10)[:, 5] == 1).astype(np.uint8) for s in syms]
    yb = capi.Fwd(max_samples=nbits * 1000, carrier_freq=cf, lpf_transition=tw).work_bits(bits)
    eb = float(np.sqrt(np.mean(np.abs(yb.astype(np.complex128) - ref) ** 2)))
    assert eb <= 1e-6, eb


def test_forward_then_reverse_loopback_on_a_carrier(capi, oracle):
    """A RECC-format burst pushed through the FORWARD modulator on carrier 2's offset, then through the receive chain
    tuned there: the GPU transmit path feeds the GPU receive path and the oracle sees the same blob."""
    c = multi.carrier(2)
    hs = synth.manchester(synth.recc_message_bits(synth.origination_words(min10=c.min10)))
    # half-symbols at 20 k/s -> forward-path symbol bytes at 100 kS/s (5 per half-symbol), +1 / 0xFF, silence (0) around
    # (four more half-symbols keep the carrier up while the receiver's filters still hold the last message symbol)
    hs_tx = np.concatenate([hs, np.array([1, 0, 1, 0], np.uint8)])
    body = np.repeat(np.where(hs_tx == 1, 1, 0xFF).astype(np.uint8), 5)
    nsym = N1 // 100
    sym = np.zeros(nsym, np.uint8)
    sym[200:200 + len(body)] = body
    fw = capi.Fwd(max_samples=N1, carrier_freq=(c.center_freq,), lpf_transition=(5e3,), out_scale=1.0)
    x = fw.work([sym])
    assert len(x) == N1
    rx = capi.ReccIq(max_samples=N1, center_freq=c.center_freq)
    b = rx.work(x)
    _, d = oracle.rx_chain_f32(x, center=c.center_freq)
    ob = oracle.rx_detect(d)
    assert bits_equal_f32(rx.read_demod(0, N1 // 50), d)
    assert len(b) == len(ob) == 1 and np.array_equal(b[0].symbols_np(), ob[0][2])
    assert np.array_equal(b[0].symbols_np(), hs[82:82 + 3374])
    assert b[0].decoded.min == c.min10.encode() and list(b[0].decoded.valid) == [1] * 7
    fw.close(); rx.close()
