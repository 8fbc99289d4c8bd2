"""GPU parity of the pieces BASELINE configs 4 and 5 lean on: the device-resident path at ANY buffer length (2^14 ... samples,
the remainder of a 1600-sample unit carried on the device), the front kernel's work split and in-kernel trigger search under
every grid size, and the batched entry points (K channels per launch, one
uploaded buffer feeding several carriers).  Everything is compared with the oracle bit for bit through the C ABI."""
import os

import numpy as np
import pytest

from gr_amps_b200 import multi, synth
from tests.helpers import bits_equal_f32, words_equal

pytestmark = pytest.mark.gpu

PASS = 38400
UNIT = 1600
N1 = 55 * PASS


@pytest.fixture(scope="module")
def capi():
    from gr_amps_b200 import capi as c
    c.lib()
    return c


@pytest.fixture(scope="module")
def torch():
    return pytest.importorskip("torch")


@pytest.fixture(scope="module")
def two_bursts(oracle):
    parts = [synth.config2_period(n_total=N1, snr_db=20.0, seed=40 + i, min10="212555%04d" % (4000 + i))[0] for i in range(2)]
    x = np.concatenate(parts)
    _, d = oracle.rx_chain_f32(x)
    return x, d, oracle.rx_detect(d)


def same_bursts(got, ob, oracle):
    assert len(got) == len(ob)
    for g, o in zip(got, ob):
        assert g.demod_index == o[0] and np.float32(g.corr) == np.float32(o[1])
        assert np.array_equal(g.symbols_np(), o[2])
        assert words_equal(g.decoded, oracle.recc_decode(o[2])) == []


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("logn", [14, 15, 17, 20])
def test_power_of_two_buffers_device_path(capi, torch, oracle, two_bursts, logn, fused):
    """BASELINE config 5's buffer sizes through amps_recc_iq_submit_dev: 2^k is not a multiple of the 1600-sample unit, the
    remainder is carried on the device; the demodulated stream and the bursts are those of the one-shot oracle."""
    x, d_orc, ob = two_bursts
    n = 1 << logn
    t = torch.from_numpy(x.view(np.float32).copy()).cuda()
    rx = capi.ReccIq(max_samples=n, fused_search=fused)
    got, pos = [], 0
    while pos + n <= len(x):
        rx.submit_dev(t.data_ptr() + 8 * pos, n, torch.cuda.current_stream().cuda_stream)
        pos += n
        if (pos // n) % 64 == 0:
            got += rx.collect()
    got += rx.collect()
    st = rx.stats()
    assert st["samples_in"] == pos // UNIT * UNIT and st["demod_out"] == pos // UNIT * 32
    nd = st["demod_out"]
    keep = min(nd, 2 * (n // 50) + 30000)                 # what the ring is guaranteed to hold
    assert bits_equal_f32(rx.read_demod(nd - keep, keep), d_orc[nd - keep:nd])
    want = oracle.rx_detect(d_orc[:nd])                   # bursts whose capture is complete in the samples fed
    same_bursts(got, want, oracle)
    assert len(want) >= 1
    rx.close()


def test_ragged_device_calls_and_mixed_sizes(capi, torch, oracle, two_bursts):
    x, d_orc, ob = two_bursts
    t = torch.from_numpy(x.view(np.float32).copy()).cuda()
    rng = np.random.default_rng(11)
    rx = capi.ReccIq(max_samples=300000)
    got, pos = [], 0
    while pos < len(x):
        n = min(int(rng.integers(1, 150000)) * 2, len(x) - pos)          # even counts (16-byte granule)
        rx.submit_dev(t.data_ptr() + 8 * pos, n, torch.cuda.current_stream().cuda_stream)
        got += rx.collect()
        pos += n
    assert rx.stats()["demod_out"] == len(x) // 50
    same_bursts(got, ob, oracle)
    rx.submit_dev(t.data_ptr(), 802, torch.cuda.current_stream().cuda_stream)
    with pytest.raises(capi.AmpsError):
        rx.work(np.zeros(10, np.complex64))            # device-path samples are pending: no mixing of the two paths
    rx.close()


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("grid", [1, 2, 7, 37, 148, 296, 0])
def test_result_does_not_depend_on_the_work_split(capi, oracle, two_bursts, grid, fused, monkeypatch):
    """AMPS_RX_GRID caps the front kernel's grid: segments of every length and, with AMPS_RX_FUSED_SEARCH, the trigger-search
    boundaries between CTAs in different places (a burst's trigger straddling two, or -- with short segments -- up to ten
    CTAs).  Without the flag the search is a launch of its own: same bursts."""
    x, d_orc, ob = two_bursts
    if grid:
        monkeypatch.setenv("AMPS_RX_GRID", str(grid))
    monkeypatch.setenv("AMPS_RX_FUSED", "1" if fused else "0")
    rx = capi.ReccIq(max_samples=len(x))
    got = rx.work(x[:3 * N1 // 2 + 1234]) + rx.work(x[3 * N1 // 2 + 1234:])
    assert bits_equal_f32(rx.read_demod(0, len(x) // 50), d_orc)
    same_bursts(got, ob, oracle)
    rx.close()


@pytest.mark.parametrize("fused", [True, False])
def test_small_calls_with_tiny_segments(capi, oracle, two_bursts, fused):
    """Calls of a few units on the full grid: every CTA owns one short tile and (fused search) the trigger search of each
    segment leans on up to nine CTAs in front of it."""
    x, d_orc, ob = two_bursts
    x = x[:N1 + 30 * PASS]
    rx = capi.ReccIq(max_samples=1 << 20, fused_search=fused)
    got, pos = [], 0
    sizes = [UNIT, 3 * UNIT, 5 * UNIT + 7, 50 * UNIT, 301 * UNIT + 333, 1 << 20]
    k = 0
    while pos < len(x):
        n = min(sizes[k % len(sizes)], len(x) - pos)
        got += rx.work(x[pos:pos + n])
        pos += n
        k += 1
    nd = rx.stats()["demod_out"]
    assert nd == len(x) // UNIT * 32
    keep = min(nd, 60000)
    assert bits_equal_f32(rx.read_demod(nd - keep, keep), d_orc[nd - keep:nd])
    same_bursts(got, oracle.rx_detect(d_orc[:nd]), oracle)
    rx.close()


# ------------------------------------------------------------------ batched launches
def carriers_period(ks, snr=20.0, n_total=N1):
    out = []
    for k in ks:
        c = multi.carrier(k)
        x, _, _ = synth.config2_period(n_total=n_total, snr_db=snr, seed=c.seed, center=c.center_freq, min10=c.min10, lead=20000 + 3000 * (k % 5))
        out.append((c, x))
    return out


@pytest.mark.parametrize("fused", [True, False])
def test_batch_of_independent_channels(capi, torch, oracle, fused):
    """Eight carriers, eight different device buffers, ONE front launch (+ one search launch unless it is fused into the
    front kernel) + ONE capture launch per call; ragged lengths."""
    cs = carriers_period(range(8))
    hs = [capi.ReccIq(max_samples=N1, center_freq=c.center_freq, fused_search=fused) for c, _ in cs]
    b = capi.ReccIqBatch(hs)
    ts = [torch.from_numpy(x.view(np.float32).copy()).cuda() for _, x in cs]
    cuts = [0, 700000, 700000 + 2 * 9999, N1]
    stream = torch.cuda.current_stream().cuda_stream
    for a, e in zip(cuts[:-1], cuts[1:]):
        b.submit_dev([t.data_ptr() + 8 * a for t in ts], e - a, stream)
    assert b.stats() == dict(calls=3, kernel_launches=6 if fused else 9)
    for h, (c, x) in zip(hs, cs):
        got = h.collect()
        _, d = oracle.rx_chain_f32(x, center=c.center_freq)
        assert bits_equal_f32(h.read_demod(0, N1 // 50), d)
        same_bursts(got, oracle.rx_detect(d), oracle)
        assert len(got) == 1 and got[0].decoded.min == c.min10.encode()
    with pytest.raises(capi.AmpsError):
        hs[0].work(np.zeros(UNIT, np.complex64))       # a batched handle is driven through its batch only
    b.close()
    for h in hs:
        h.close()


def test_batch_larger_than_one_launch_and_uneven_lengths(capi, torch, oracle):
    """70 channels (more than one launch's 64) on 3 distinct buffers, channel i fed in calls of its own length."""
    cs = carriers_period([0, 3, 6], n_total=N1)
    K = 70
    hs = [capi.ReccIq(max_samples=N1, center_freq=cs[i % 3][0].center_freq, fused_search=(K % 2 == 1)) for i in range(K)]
    b = capi.ReccIqBatch(hs, time_kernels=True)
    ts = [torch.from_numpy(x.view(np.float32).copy()).cuda() for _, x in cs]
    stream = torch.cuda.current_stream().cuda_stream
    step = [2 * (40000 + 1111 * (i % 7)) for i in range(K)]
    pos = [0] * K
    while any(p < N1 for p in pos):
        ns = [min(step[i], N1 - pos[i]) for i in range(K)]
        b.submit_dev([ts[i % 3].data_ptr() + 8 * pos[i] for i in range(K)], ns, stream)
        pos = [p + n for p, n in zip(pos, ns)]
    assert len(b.front_times_ms()) > 0
    ref = []
    for c, x in cs:
        _, d = oracle.rx_chain_f32(x, center=c.center_freq)
        ref.append((d, oracle.rx_detect(d)))
    for i, h in enumerate(hs):
        d, ob = ref[i % 3]
        assert h.stats()["demod_out"] == N1 // 50
        assert bits_equal_f32(h.read_demod(0, N1 // 50), d), "channel %d" % i
        same_bursts(h.collect(), ob, oracle)
    b.close()
    for h in hs:
        h.close()


@pytest.mark.parametrize("sc16", [False, True])
def test_one_uploaded_buffer_feeds_several_carriers(capi, oracle, sc16):
    """Carriers 0, 2, 5, 7 transmitted into ONE wideband buffer; amps_recc_iq_batch_work_shared uploads it once and every
    channel recovers its own burst -- each identical to a stand-alone handle fed the same buffer, and to the oracle."""
    ks = [0, 2, 5, 7]
    cs = carriers_period(ks, snr=None)
    rng = np.random.default_rng(77)
    x = sum(xc for _, xc in cs)
    x = (x + 0.05 * (rng.standard_normal(N1) + 1j * rng.standard_normal(N1))).astype(np.complex64)
    kw = {}
    feed = x
    if sc16:
        scale = 1.0 / 8192.0
        q = np.clip(np.round(x.view(np.float32) / scale), -32768, 32767).astype(np.int16)
        x = (q.astype(np.float32) * np.float32(scale)).view(np.complex64)
        feed = q
        kw = dict(sc16=True, sc16_scale=scale)
    hs = [capi.ReccIq(max_samples=N1, center_freq=c.center_freq, **kw) for c, _ in cs]
    b = capi.ReccIqBatch(hs)
    got = []
    per = 2 if sc16 else 1
    cuts = [0, 123457, 1000001, N1]
    for a, e in zip(cuts[:-1], cuts[1:]):
        got += b.work_shared(feed[per * a:per * e])
    assert sorted(ch for ch, _ in got) == [0, 1, 2, 3]
    for i, (c, _) in enumerate(cs):
        _, d = oracle.rx_chain_f32(x, center=c.center_freq)
        ob = oracle.rx_detect(d)
        mine = [bb for ch, bb in got if ch == i]
        assert bits_equal_f32(hs[i].read_demod(0, N1 // 50), d)
        same_bursts(mine, ob, oracle)
        assert mine[0].decoded.min == c.min10.encode() and list(mine[0].decoded.valid) == [1] * 7
    b.close()
    for h in hs:
        h.close()


def test_batch_argument_checks(capi, torch):
    a = capi.ReccIq(max_samples=PASS)
    m = capi.ReccIq(max_samples=PASS, timing_mm=True)
    n4 = capi.ReccIq(max_samples=1536, samp_rate=400e3)
    s = capi.ReccIq(max_samples=PASS, sc16=True)
    f = capi.ReccIq(max_samples=PASS, fused_search=True)
    for bad in ([a, m], [a, n4], [a, s], [a, a], [a, f], []):
        with pytest.raises((capi.AmpsError, ValueError)):
            capi.ReccIqBatch(bad)
    b = capi.ReccIqBatch([a])
    with pytest.raises(capi.AmpsError):
        capi.ReccIqBatch([a])                          # already in a batch
    t = torch.zeros(2 * PASS, dtype=torch.float32, device="cuda")
    with pytest.raises(capi.AmpsError):
        b.submit_dev([t.data_ptr()], 2 * PASS, 0)      # more than max_samples
    with pytest.raises(capi.AmpsError):
        b.submit_dev([t.data_ptr()], 1601, 0)          # odd count
    b.submit_dev([t.data_ptr()], 0, 0)
    b.close()
    assert a.work(np.zeros(UNIT, np.complex64)) == []  # usable on its own again
    for h in (a, m, n4, s, f):
        h.close()


@pytest.mark.parametrize("fused", [True, False])
def test_large_call_many_bursts(capi, torch, oracle, fused):
    """22 periods = 46.5 M samples in one submit (about 33 tiles per CTA on the full grid), against the one-shot oracle."""
    periods = [synth.config2_period(n_total=N1, snr_db=20.0, seed=500 + i, min10="212555%04d" % (5000 + i))[0] for i in range(3)]
    x = np.concatenate([periods[i % 3] for i in range(22)])
    t = torch.from_numpy(x.view(np.float32).copy()).cuda()
    rx = capi.ReccIq(max_samples=len(x), fused_search=fused)
    rx.submit_dev(t.data_ptr(), len(x), torch.cuda.current_stream().cuda_stream)
    got = rx.collect()
    _, d = oracle.rx_chain_f32(x)
    ob = oracle.rx_detect(d, max_bursts=64)
    assert bits_equal_f32(rx.read_demod(0, len(x) // 50), d)
    assert len(ob) == 22
    same_bursts(got, ob, oracle)
    rx.submit_dev(t.data_ptr(), len(x), torch.cuda.current_stream().cuda_stream)
    assert len(rx.collect()) == 22
    rx.close()
