"""GPU parity of the fused RECC receive path (BASELINE config 2) against the CPU oracle.
Everything goes through the C ABI (gr_amps_b200.capi -> libamps_b200.so)."""
import numpy as np
import pytest

from gr_amps_b200 import synth
from tests.helpers import bits_equal_f32, words_equal

pytestmark = pytest.mark.gpu

PASS = 38400             # one full pass of the front kernel (24 units)
UNIT = 1600              # amps_recc_iq_granularity() at 10 MS/s: 32 demodulated samples, one word of hard decisions
N1 = 55 * PASS           # one config-2 period (2 112 000 samples)


@pytest.fixture(scope="module")
def capi():
    from gr_amps_b200 import capi as c
    c.lib()
    return c


@pytest.mark.parametrize("snr", [None, 30.0, 15.0])
def test_single_burst_bit_exact(capi, oracle, snr):
    x, hs, words = synth.config2_period(n_total=N1, snr_db=snr)
    rx = capi.ReccIq(max_samples=N1, dump_baseband=True, lpf_taps=oracle.lpf_taps())
    bursts = rx.work(x)
    nd = N1 // 50
    d_gpu = rx.read_demod(0, nd)
    y_gpu = rx.read_baseband(0, nd)
    y_orc, d_orc = oracle.rx_chain_f32(x)
    # hard decisions and the soft stream itself are bit-exact against the fp32 kernel-spec oracle
    assert np.array_equal(d_gpu >= 0, d_orc >= 0)
    assert bits_equal_f32(d_gpu, d_orc)
    assert bits_equal_f32(y_gpu.view(np.float32), y_orc.view(np.float32))
    # complex baseband within 1e-6 RMS of the float64 ideal chain (north_star tolerance)
    y64, _ = oracle.rx_chain_f64(x)
    rms = np.sqrt(np.mean(np.abs(y_gpu.astype(np.complex128) - y64) ** 2))
    assert rms <= 1e-6, rms
    ob = oracle.rx_detect(d_orc)
    assert len(bursts) == len(ob) == 1
    b = bursts[0]
    assert b.demod_index == ob[0][0] and b.sample_index == 50 * ob[0][0]
    assert np.float32(b.corr) == np.float32(ob[0][1])
    assert np.array_equal(b.symbols_np(), ob[0][2])
    assert np.array_equal(b.symbols_np(), hs[82:82 + 3374])          # what was transmitted
    assert words_equal(b.decoded, oracle.recc_decode(ob[0][2])) == []
    assert b.decoded.kind == 4 and b.decoded.min == b"2125551234" and b.decoded.dialed == b"18005551212"
    assert list(b.decoded.valid) == [1] * 7
    rx.close()


def test_default_taps_match_oracle(capi, oracle):
    rx = capi.ReccIq(max_samples=PASS)
    t = rx.taps()
    assert len(t) == 299 and np.array_equal(t, oracle.lpf_taps())
    rx.close()


def test_chunked_stream_equals_one_shot(capi, oracle):
    """Streaming invariance: arbitrary work() chunking gives the same bursts and the same demod stream."""
    x, hs, _ = synth.config2_period(n_total=N1, snr_db=20.0, seed=7)
    x = np.concatenate([x, x])          # two bursts
    one = capi.ReccIq(max_samples=len(x))
    b1 = one.work(x)
    d1 = one.read_demod(0, len(x) // 50)
    rng = np.random.default_rng(3)
    st = capi.ReccIq(max_samples=400000)
    got, pos = [], 0
    while pos < len(x):
        n = int(rng.integers(1, 400000))
        got += st.work(x[pos:pos + n])
        pos += n
    assert st.granularity == UNIT and st.stats()["demod_out"] == len(x) // UNIT * (UNIT // 50)
    assert len(b1) == 2 and len(got) == 2
    for a, b in zip(b1, got):
        assert a.demod_index == b.demod_index and bytes(a.symbols) == bytes(b.symbols)
        assert np.float32(a.corr) == np.float32(b.corr)
    y_orc, d_orc = oracle.rx_chain_f32(x)
    assert bits_equal_f32(d1, d_orc[:len(d1)])
    one.close(); st.close()


def test_device_resident_many_bursts(capi, oracle):
    """submit_dev on a device-resident buffer (the bench path): 8 periods, different noise each."""
    torch = pytest.importorskip("torch")
    periods = [synth.config2_period(n_total=N1, snr_db=18.0, seed=100 + i, min10="2125550%03d" % i)[0] for i in range(8)]
    x = np.concatenate(periods)
    t = torch.from_numpy(x.view(np.float32).copy()).cuda()
    rx = capi.ReccIq(max_samples=len(x))
    rx.submit_dev(t.data_ptr(), len(x), torch.cuda.current_stream().cuda_stream)
    bursts = rx.collect()
    _, d_orc = oracle.rx_chain_f32(x)
    ob = oracle.rx_detect(d_orc)
    assert len(bursts) == len(ob) == 8
    for i, (b, o) in enumerate(zip(bursts, ob)):
        assert b.demod_index == o[0] and np.array_equal(b.symbols_np(), o[2])
        assert words_equal(b.decoded, oracle.recc_decode(o[2])) == []
        assert b.decoded.min == ("2125550%03d" % i).encode()
    rx.close()


# ------------------------------------------------------------------ edge cases
def test_empty_and_tiny_calls(capi):
    rx = capi.ReccIq(max_samples=PASS * 4)
    assert rx.work(np.zeros(0, np.complex64)) == []
    for n in (1, 7, 49, 50, 1492):                       # less than one unit: carried, nothing produced yet
        assert rx.work(np.zeros(n, np.complex64)) == []
        assert rx.stats()["demod_out"] == 0
    for n in (1, 38399):
        assert rx.work(np.zeros(n, np.complex64)) == []
    st = rx.stats()
    assert st["demod_out"] == (1 + 7 + 49 + 50 + 1492 + 1 + 38399) // UNIT * (UNIT // 50)
    with pytest.raises(capi.AmpsError):
        rx.work(np.zeros(PASS * 4 + PASS + 1, np.complex64))     # more than max_samples
    import torch
    t = torch.zeros(2 * PASS + 2, dtype=torch.float32, device="cuda")
    rx2 = capi.ReccIq(max_samples=PASS * 4)
    with pytest.raises(capi.AmpsError):
        rx2.submit_dev(t.data_ptr(), PASS + 1, 0)        # an odd count: the byte length is not a multiple of 16
    with pytest.raises(capi.AmpsError):
        rx2.submit_dev(t.data_ptr() + 8, PASS, 0)        # not 16-byte aligned
    rx.close(); rx2.close()


def test_noise_only_and_silence_give_no_bursts(capi, oracle):
    rng = np.random.default_rng(5)
    n = 20 * PASS
    noise = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    for x in (noise, np.zeros(n, np.complex64)):
        rx = capi.ReccIq(max_samples=n)
        assert rx.work(x) == []
        _, d = oracle.rx_chain_f32(x)
        assert bits_equal_f32(rx.read_demod(0, n // 50), d)
        assert oracle.rx_detect(d) == []
        rx.close()


def test_truncated_burst_and_overlapping_triggers(capi, oracle):
    """(a) a burst cut off before its 3374 symbols are in: not reported (and reported once the rest arrives);
    (b) a second seizure precursor inside an already captured burst is ignored, a later one is captured."""
    x, hs, _ = synth.config2_period(n_total=N1, snr_db=25.0, seed=21)
    cut = 30 * PASS                                       # 1 152 000 samples: trigger in, capture incomplete
    rx = capi.ReccIq(max_samples=N1)
    assert rx.work(x[:cut]) == []
    got = rx.work(x[cut:])
    one = capi.ReccIq(max_samples=N1)
    ref = one.work(x)
    assert len(got) == len(ref) == 1 and got[0].demod_index == ref[0].demod_index and bytes(got[0].symbols) == bytes(ref[0].symbols)
    rx.close(); one.close()
    # (b): inject the 37-bit trigger pattern again 600 half-symbols into the message (inside the capture window)
    words = synth.origination_words()
    bits = synth.recc_message_bits(words).copy()
    trig_bits = np.array([1, 0] * 13 + synth.WORD_SYNC, np.uint8)
    bits[300:337] = trig_bits
    hs2 = synth.manchester(bits)
    xa = synth.fm_burst(hs2, N1, 20000, snr_db=25.0, seed=22)
    xb, _, _ = synth.config2_period(n_total=N1, snr_db=25.0, seed=23)
    xx = np.concatenate([xa, xb])
    rx = capi.ReccIq(max_samples=len(xx))
    b = rx.work(xx)
    _, d = oracle.rx_chain_f32(xx)
    ob = oracle.rx_detect(d)
    assert len(b) == len(ob) == 2                         # the embedded trigger did not produce a third record
    for g, o in zip(b, ob):
        assert g.demod_index == o[0] and np.array_equal(g.symbols_np(), o[2])
        assert words_equal(g.decoded, oracle.recc_decode(o[2])) == []
    assert b[1].decoded.min == b"2125551234" and list(b[1].decoded.valid) == [1] * 7
    rx.close()


def test_reset_restarts_the_stream(capi):
    x, _, _ = synth.config2_period(n_total=N1, snr_db=22.0, seed=31)
    rx = capi.ReccIq(max_samples=N1)
    a = rx.work(x)
    rx.reset()
    b = rx.work(x)
    assert len(a) == len(b) == 1 and a[0].demod_index == b[0].demod_index and bytes(a[0].symbols) == bytes(b[0].symbols)
    assert np.float32(a[0].corr) == np.float32(b[0].corr)
    rx.close()


def test_ring_overflow_is_reported_not_corrupting(capi):
    """More bursts than the host ring holds between collects: the newest max_bursts survive, in order."""
    torch = pytest.importorskip("torch")
    x, _, _ = synth.config2_period(n_total=N1, snr_db=None)
    xs = np.tile(x, 6)
    t = torch.from_numpy(xs.view(np.float32).copy()).cuda()
    rx = capi.ReccIq(max_samples=len(xs), max_bursts=4)
    rx.submit_dev(t.data_ptr(), len(xs), torch.cuda.current_stream().cuda_stream)
    got = rx.collect()
    assert len(got) == 4
    idx = [g.demod_index for g in got]
    assert idx == sorted(idx) and idx[-1] // (N1 // 50) == 5
    rx.close()


# ------------------------------------------------------------------ the reference's own rate: 400 kS/s
def burst_400k(seed, snr, words=None, n_total=55 * 1536, lead=800):
    words = words or synth.origination_words()
    hs = synth.manchester(synth.recc_message_bits(words))
    return synth.fm_burst(hs, n_total, lead, samp_rate=400e3, snr_db=snr, seed=seed), hs


@pytest.mark.parametrize("snr", [None, 20.0])
def test_native_400k_rate_bit_exact(capi, oracle, snr):
    """recc_iq at the reference's operating point (grc/ampsbs.grc:263): NCO + lpf_taps /2 + quadrature demod are then
    exactly freq_xlating_fir_filter_ccc + quadrature_demod_cf; bit-exact vs the fp32 oracle, 1e-6 RMS vs float64."""
    x, hs = burst_400k(5, snr)
    n = len(x)
    rx = capi.ReccIq(max_samples=n, samp_rate=400e3, dump_baseband=True)
    assert rx.granularity == 1536
    bursts = rx.work(x)
    y_orc, d_orc = oracle.rx_chain400_f32(x)
    assert bits_equal_f32(rx.read_demod(0, n // 2), d_orc)
    y_gpu = rx.read_baseband(0, n // 2)
    assert bits_equal_f32(y_gpu.view(np.float32), y_orc.view(np.float32))
    y64, _ = oracle.rx_chain400_f64(x)
    assert np.sqrt(np.mean(np.abs(y_gpu.astype(np.complex128) - y64) ** 2)) <= 1e-6
    ob = oracle.rx_detect(d_orc)
    assert len(bursts) == len(ob) == 1
    b = bursts[0]
    assert b.demod_index == ob[0][0] and b.sample_index == 2 * ob[0][0]
    assert np.array_equal(b.symbols_np(), ob[0][2]) and np.array_equal(b.symbols_np(), hs[82:82 + 3374])
    assert words_equal(b.decoded, oracle.recc_decode(ob[0][2])) == []
    rx.close()


def test_native_400k_streaming_small_chunks(capi, oracle):
    """The real-time shape of the reference graph: small scheduler buffers (here 1..4096 samples) at 400 kS/s."""
    parts = [burst_400k(60 + i, 18.0, w)[0] for i, w in enumerate(
        [synth.origination_words(min10="2125550111"), synth.page_response_words(min10="2125550112"), synth.registration_words(min10="2125550113")])]
    x = np.concatenate(parts + [np.zeros(1536, np.complex64)])
    rx = capi.ReccIq(max_samples=8192, samp_rate=400e3)
    rng = np.random.default_rng(4)
    got, pos = [], 0
    while pos < len(x):
        n = int(rng.integers(1, 4097))
        got += rx.work(x[pos:pos + n])
        pos += n
    _, d = oracle.rx_chain400_f32(x)
    ob = oracle.rx_detect(d)
    assert len(got) == len(ob) == 3
    for g, o in zip(got, ob):
        assert g.demod_index == o[0] and np.array_equal(g.symbols_np(), o[2])
    assert [g.decoded.kind for g in got] == [4, 2, 3]
    assert [g.decoded.min for g in got] == [b"2125550111", b"2125550112", b"2125550113"]
    rx.close()
