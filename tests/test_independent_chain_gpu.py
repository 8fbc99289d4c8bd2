"""The CUDA path against the INDEPENDENT scipy implementation of the stock GNU Radio blocks (tests/scipy_chain.py) -- not
against anything of ours: complex baseband <= 1e-6 RMS (the north-star tolerance), demodulated stream within fp32 rounding,
hard decisions equal wherever the decision is not within rounding of zero, and the captured blob equal to the transmitted
half-symbols.  At 400 kS/s the chain under test is exactly the reference's graph (freq_xlating_fir_filter_ccc ->
quadrature_demod_cf, grc/ampsbs.grc:1814-1872, 774-816) with nothing of ours in front."""
import numpy as np
import pytest

from gr_amps_b200 import synth
from tests import scipy_chain as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from gr_amps_b200 import capi as c
    c.lib()
    return c


def rms(a):
    return float(np.sqrt(np.mean(np.abs(a) ** 2)))


def wrap(a):
    return np.angle(np.exp(1j * a))


@pytest.mark.parametrize("g", [0, 2, 7])
@pytest.mark.parametrize("snr", [None, 20.0])
def test_native_rate_against_gnuradio_structure(capi, oracle, g, snr):
    fc = -160e3 + 30e3 * g
    hs = synth.manchester(synth.recc_message_bits(synth.origination_words()))
    x = synth.fm_burst(hs, 55 * 1536, 800, samp_rate=400e3, center=fc, snr_db=snr, seed=70 + g)
    rx = capi.ReccIq(max_samples=len(x), samp_rate=400e3, center_freq=fc, dump_baseband=True)
    bursts = rx.work(x)
    ys, ds = S.rx_chain_400k(x, oracle.lpf_taps(), fc)
    y = rx.read_baseband(0, len(x) // 2).astype(np.complex128)
    d = rx.read_demod(0, len(x) // 2)
    assert rms(y - ys) <= 1e-6, rms(y - ys)
    strong = np.abs(ys) > 1e-3
    strong[1:] &= strong[:-1]
    assert np.max(np.abs(wrap(d - ds))[strong]) <= 2e-4
    sure = strong & (np.abs(ds) > 1e-3)
    assert np.array_equal((d >= 0)[sure], (ds >= 0)[sure])
    assert len(bursts) == 1 and np.array_equal(bursts[0].symbols_np(), hs[82:82 + 3374])
    # the sampling instants the GPU chose, sliced out of the INDEPENDENT demod stream, give the same blob
    pos = bursts[0].demod_index
    assert np.array_equal((ds[pos + 10 * (74 + np.arange(3374))] >= 0).astype(np.uint8), bursts[0].symbols_np())
    rx.close()


@pytest.mark.parametrize("g", [0, 5])
def test_10ms_rate_against_scipy(capi, oracle, g):
    fc = -160e3 + 30e3 * g
    n = 55 * 38400
    x, hs, _ = synth.config2_period(n_total=n, snr_db=15.0, seed=80 + g, center=fc)
    rx = capi.ReccIq(max_samples=n, center_freq=fc, dump_baseband=True)
    bursts = rx.work(x)
    ys, ds = S.rx_chain_10m(x, oracle.lpf_taps(), fc)
    y = rx.read_baseband(0, n // 50).astype(np.complex128)
    d = rx.read_demod(0, n // 50)
    assert rms(y - ys) <= 1e-6, rms(y - ys)
    strong = np.abs(ys) > 1e-2
    strong[1:] &= strong[:-1]
    sure = strong & (np.abs(ds) > 1e-3)
    assert np.array_equal((d >= 0)[sure], (ds >= 0)[sure])
    assert len(bursts) == 1 and np.array_equal(bursts[0].symbols_np(), hs[82:82 + 3374])
    pos = bursts[0].demod_index
    assert np.array_equal((ds[pos + 10 * (74 + np.arange(3374))] >= 0).astype(np.uint8), bursts[0].symbols_np())
    rx.close()


def test_forward_path_against_scipy(capi, oracle):
    nsym = 21000
    focc = oracle.Focc(100000, False).generate(nsym, chunk=1 << 20)
    v = oracle.Fvc(100000)
    v.push_words(oracle.word("orc_fvc_word1_general", 1, 0, 0, 1))
    out = bytearray()
    while len(out) < nsym:
        _, b, _ = v.work(min(8192, nsym - len(out)))
        out += b.tobytes()
    fvc = np.frombuffer(bytes(out), np.uint8).copy()
    syms = [focc, fvc, fvc.copy()]
    cf, tw = (0.0, 60e3, 90e3), (5e3, 3e3, 3e3)
    fw = capi.Fwd(max_samples=nsym * 100, carrier_freq=cf, lpf_transition=tw)
    taps = [fw.taps(c) for c in range(3)]
    ys = S.fwd_chain_10m(syms, taps, cf, scale=0.5)
    y = fw.work(syms).astype(np.complex128)
    assert rms(y - ys) <= 1e-6, rms(y - ys)
    bits = [(np.asarray(s).reshape(-1, 10)[:, 5] == 1).astype(np.uint8) for s in syms]
    yb = capi.Fwd(max_samples=nsym * 100, carrier_freq=cf, lpf_transition=tw).work_bits(bits).astype(np.complex128)
    assert rms(yb - ys) <= 1e-6, rms(yb - ys)
    fw.close()
