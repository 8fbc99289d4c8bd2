"""GPU parity of the M&M timing mode (AMPS_RX_TIMING_MM) of the fused receive path: the reference graph's own
serial tail clock_recovery_mm_ff -> binary_slicer_fb -> amps.recc (grc/ampsbs.grc:1751-1813, 1712-1750;
lib/recc_impl.cc:93-145) against oracle/mm_timing.c + oracle/recc_capture.c, bit-exact, through the C ABI."""
import numpy as np
import pytest

from gr_amps_b200 import synth
from tests.helpers import words_equal

pytestmark = pytest.mark.gpu

PASS = 38400
N1 = 55 * PASS
QUANTUM = 256            # bytes per emulated amps.recc work() call (include/amps_b200.h, AMPS_RX_TIMING_MM)


@pytest.fixture(scope="module")
def capi():
    from gr_amps_b200 import capi as c
    c.lib()
    return c


class OracleTail:
    """The serial tail on the oracle side, fed call by call with the demod count the GPU handle reports."""

    def __init__(self, oracle, d):
        self.o, self.d = oracle, np.ascontiguousarray(d, dtype=np.float32)
        self.mm, self.recc = oracle.MmTiming(), oracle.Recc()
        self.nsym = 0

    def advance(self, total_d):
        syms = self.mm.process(self.d, total_d)
        self.nsym += len(syms)
        before = len(self.recc.bursts)
        for i in range(0, len(syms), QUANTUM):
            self.recc.work(syms[i:i + QUANTUM])
        return self.recc.bursts[before:]


def three_bursts(snr, seed0=100):
    xs, hss = [], []
    for k in range(3):
        x, hs = synth.burst_period(synth.origination_words(), lead=20000 + 777 * k, snr_db=snr, seed=seed0 + k)
        xs.append(x); hss.append(hs)
    return np.concatenate(xs), hss


@pytest.mark.parametrize("snr", [None, 30.0, 15.0, 8.0])
def test_mm_tail_bit_exact_streaming(capi, oracle, snr):
    x, hss = three_bursts(snr)
    _, d = oracle.rx_chain_f32(x)
    tail = OracleTail(oracle, d)
    rx = capi.ReccIq(max_samples=600000, timing_mm=True)
    rng = np.random.default_rng(11)
    pos, n_total = 0, 0
    while pos < len(x):
        n = int(rng.integers(1, 600000))
        got = rx.work(x[pos:pos + n])
        pos += n
        want = tail.advance(rx.stats()["demod_out"])
        assert len(got) == len(want)                       # delivered by the same call as in the oracle's schedule
        for g, w in zip(got, want):
            assert np.array_equal(g.symbols_np(), w)
            assert words_equal(g.decoded, oracle.recc_decode(w)) == []
            assert g.corr == 0.0 and g.run_length == 0
        n_total += len(got)
    assert n_total == 3
    assert tail.nsym > 3 * 4200                            # the loop free-runs over noise too: ~1 symbol per 10 demod samples
    if snr is None or snr >= 15.0:
        assert all(list(oracle.recc_decode(b).valid) == [1] * 7 for b in tail.recc.bursts)
    if snr == 30.0:                                        # and at a comfortable SNR it is what was transmitted
        for b, hs in zip(tail.recc.bursts, hss):
            assert np.array_equal(b, hs[82:82 + 3374])
    rx.close()


def test_mm_device_path_and_reset(capi, oracle):
    import torch
    x, _ = three_bursts(20.0, seed0=7)
    _, d = oracle.rx_chain_f32(x)
    rx = capi.ReccIq(max_samples=len(x), timing_mm=True)
    xd = torch.from_numpy(x.view(np.float32).copy()).cuda()
    for rep in range(2):
        tail = OracleTail(oracle, d)
        half = (len(x) // PASS // 2) * PASS
        rx.submit_dev(xd.data_ptr(), half)
        rx.submit_dev(xd.data_ptr() + half * 8, len(x) - half)
        got = rx.collect()
        want = tail.advance(half // 50) + tail.advance(len(x) // 50)
        assert len(got) == len(want) == 3
        for g, w in zip(got, want):
            assert np.array_equal(g.symbols_np(), w)
            assert g.decoded.min == b"2125551234" and list(g.decoded.valid) == [1] * 7
        # nominal position (recovered half-symbol count x 10): the loop free-runs within +-0.5 % of the nominal rate,
        # so it stays within 1 % of where the feed-forward detector puts the trigger
        ff = oracle.rx_detect(d)
        assert [abs(int(g.demod_index) - p) < 0.01 * p + 200 for g, (p, _, _) in zip(got, ff)] == [True] * 3
        assert [int(g.sample_index) for g in got] == [50 * int(g.demod_index) for g in got]
        rx.reset()                                         # back to stream start: the second round repeats the first
    rx.close()


def test_mm_native_400k(capi, oracle):
    """At the reference's own 400 kS/s the whole chain is then the reference graph block for block."""
    words = synth.origination_words()
    hs = synth.manchester(synth.recc_message_bits(words))
    n = 1536 * 60
    x = synth.fm_burst(hs, n, 2000, samp_rate=400e3, snr_db=25.0, seed=5)
    x = np.concatenate([x, x])
    _, d = oracle.rx_chain400_f32(x)
    tail = OracleTail(oracle, d)
    rx = capi.ReccIq(max_samples=len(x), samp_rate=400e3, timing_mm=True)
    got = rx.work(x)
    want = tail.advance(rx.stats()["demod_out"])
    assert len(got) == len(want) == 2
    for g, w in zip(got, want):
        assert np.array_equal(g.symbols_np(), w) and np.array_equal(w, hs[82:82 + 3374])
        assert words_equal(g.decoded, oracle.recc_decode(w)) == []
    rx.close()


def test_mm_and_feed_forward_agree_on_clean_bursts(capi, oracle):
    """Both timing methods recover the transmitted half-symbols at 30 dB."""
    x, hss = three_bursts(30.0, seed0=40)
    a = capi.ReccIq(max_samples=len(x), timing_mm=True)
    b = capi.ReccIq(max_samples=len(x))
    ga, gb = a.work(x), b.work(x)
    assert len(ga) == len(gb) == 3
    for p, q, hs in zip(ga, gb, hss):
        assert np.array_equal(p.symbols_np(), hs[82:82 + 3374]) and np.array_equal(q.symbols_np(), hs[82:82 + 3374])
    a.close(); b.close()


def test_mm_tail_batched_channels(capi, oracle):
    """The M&M tails of several channels in ONE pair of launches (one CTA per channel walks its recurrence, side by side): three
    carriers in one wideband buffer, uploaded once; every channel delivers, call by call, what the oracle's tail delivers."""
    from gr_amps_b200 import multi
    ks = [0, 3, 6]
    parts = []
    for k in ks:
        c = multi.carrier(k)
        x, _ = synth.burst_period(synth.origination_words(min10=c.min10), lead=20000 + 5000 * k, snr_db=None, center=c.center_freq)
        parts.append(x)
    rng = np.random.default_rng(5)
    x = (sum(parts) + 0.03 * (rng.standard_normal(N1) + 1j * rng.standard_normal(N1))).astype(np.complex64)
    hs = [capi.ReccIq(max_samples=N1, center_freq=multi.carrier(k).center_freq, timing_mm=True) for k in ks]
    b = capi.ReccIqBatch(hs)
    tails = []
    for k in ks:
        _, d = oracle.rx_chain_f32(x, center=multi.carrier(k).center_freq)
        tails.append(OracleTail(oracle, d))
    pos, total = 0, [0, 0, 0]
    for n in (300001, 700000, 555555, N1):
        n = min(n, N1 - pos)
        got = b.work_shared(x[pos:pos + n])
        pos += n
        for i, h in enumerate(hs):
            want = tails[i].advance(h.stats()["demod_out"])
            mine = [g for ch, g in got if ch == i]
            assert len(mine) == len(want)
            for g, w in zip(mine, want):
                assert np.array_equal(g.symbols_np(), w) and words_equal(g.decoded, oracle.recc_decode(w)) == []
            total[i] += len(mine)
        if pos >= N1:
            break
    assert total == [1, 1, 1]
    assert b.stats()["kernel_launches"] == 4 * b.stats()["calls"]          # front, M&M, amps.recc, capture -- whatever K is
    b.close()
    for h in hs:
        h.close()
