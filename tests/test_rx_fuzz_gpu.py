"""Randomised GPU parity of the receive path: random call lengths (1 ... 700 000 samples, even for the device path), random
front-kernel grids, search fused into the front kernel or as its own launch, fc32 / sc16, host / device-resident / batched
entry points -- the demodulated stream and the bursts must be the one-shot oracle's whatever the cut."""
import numpy as np
import pytest

from gr_amps_b200 import multi, synth
from tests.helpers import bits_equal_f32, words_equal

pytestmark = pytest.mark.gpu
N1 = 55 * 38400


@pytest.fixture(scope="module")
def capi():
    from gr_amps_b200 import capi as c
    c.lib()
    return c


@pytest.fixture(scope="module")
def stream(oracle):
    """three bursts: a normal one, one whose trigger is repeated inside the message, one close to the previous"""
    xs = []
    for i, lead in enumerate((20000, 150000, 3000)):
        words = synth.origination_words(min10="21255577%02d" % i)
        bits = synth.recc_message_bits(words).copy()
        if i == 1:
            bits[300:337] = np.array([1, 0] * 13 + synth.WORD_SYNC, np.uint8)
        hs = synth.manchester(bits)
        xs.append(synth.fm_burst(hs, N1 - 100000 * (i == 2), lead, snr_db=[25.0, 18.0, None][i], seed=900 + i))
    x = np.concatenate(xs)
    x = x[:len(x) // 1600 * 1600]
    _, d = oracle.rx_chain_f32(x)
    return x, d, oracle.rx_detect(d)


def check(got, ob, oracle):
    assert len(got) == len(ob)
    for g, o in zip(got, ob):
        assert g.demod_index == o[0] and np.float32(g.corr) == np.float32(o[1]) and np.array_equal(g.symbols_np(), o[2])
        assert words_equal(g.decoded, oracle.recc_decode(o[2])) == []


@pytest.mark.parametrize("seed", range(8))
def test_random_cuts_grids_and_modes(capi, oracle, stream, seed, monkeypatch):
    torch = pytest.importorskip("torch")
    x, d_orc, ob = stream
    rng = np.random.default_rng(1000 + seed)
    fused = bool(seed & 1)
    device_path = bool(seed & 2)
    grid = int(rng.choice([0, 1, 3, 11, 50, 149, 296]))
    if grid:
        monkeypatch.setenv("AMPS_RX_GRID", str(grid))
    big = int(rng.choice([5000, 70000, 700000]))
    rx = capi.ReccIq(max_samples=big + 2, fused_search=fused)
    t = torch.from_numpy(x.view(np.float32).copy()).cuda() if device_path else None
    got, pos = [], 0
    while pos < len(x):
        n = min(int(rng.integers(1, big)), len(x) - pos)
        if device_path:
            n = max(2, n - (n & 1)) if len(x) - pos >= 2 else len(x) - pos
            rx.submit_dev(t.data_ptr() + 8 * pos, n, torch.cuda.current_stream().cuda_stream)
            if rng.integers(0, 4) == 0:
                got += rx.collect()
        else:
            got += rx.work(x[pos:pos + n])
        pos += n
    got += rx.collect()
    nd = rx.stats()["demod_out"]
    assert nd == len(x) // 50
    keep = min(nd, 2 * (big // 50) + 30000)
    assert bits_equal_f32(rx.read_demod(nd - keep, keep), d_orc[nd - keep:nd])
    check(got, ob, oracle)
    rx.close()


@pytest.mark.parametrize("seed", range(3))
def test_random_batches(capi, oracle, seed):
    """K channels with their own carriers and their own call lengths, sc16 or fc32, fused or not, one launch set per call."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(2000 + seed)
    K = int(rng.choice([2, 5, 9]))
    sc16 = bool(seed & 1)
    fused = bool(seed & 2)
    scale = 1.0 / 8192.0
    chans = []
    for k in range(K):
        c = multi.carrier(int(rng.integers(0, 8)))
        x, _, _ = synth.config2_period(n_total=N1, snr_db=float(rng.choice([15.0, 25.0])), seed=3000 + 10 * seed + k, center=c.center_freq,
                                       min10=c.min10, lead=int(rng.integers(2000, 300000)))
        if sc16:
            q = np.clip(np.round(x.view(np.float32) / scale), -32768, 32767).astype(np.int16)
            x = (q.astype(np.float32) * np.float32(scale)).view(np.complex64)
            dev = torch.from_numpy(q.copy()).cuda()
        else:
            dev = torch.from_numpy(x.view(np.float32).copy()).cuda()
        _, d = oracle.rx_chain_f32(x, center=c.center_freq)
        chans.append((c, dev, d, oracle.rx_detect(d)))
    hs = [capi.ReccIq(max_samples=N1, center_freq=c.center_freq, sc16=sc16, sc16_scale=scale if sc16 else 0.0, fused_search=fused) for c, _, _, _ in chans]
    b = capi.ReccIqBatch(hs)
    isz = 4 if sc16 else 8
    step = 4 if sc16 else 2
    pos = [0] * K
    st = torch.cuda.current_stream().cuda_stream
    while any(p < N1 for p in pos):
        ns = []
        for k in range(K):
            n = min(int(rng.integers(0, 400000)) // step * step, N1 - pos[k])
            ns.append(n)
        b.submit_dev([chans[k][1].data_ptr() + isz * pos[k] for k in range(K)], ns, st)
        pos = [p + n for p, n in zip(pos, ns)]
    for k, h in enumerate(hs):
        assert h.stats()["demod_out"] == N1 // 50
        assert bits_equal_f32(h.read_demod(0, N1 // 50), chans[k][2]), "channel %d" % k
        check(h.collect(), chans[k][3], oracle)
    b.close()
    for h in hs:
        h.close()


@pytest.mark.parametrize("collect_between", [False, True])
def test_mixed_call_sizes_pipelined(capi, oracle, collect_between):
    """Device-resident calls of very different lengths back to back without a synchronisation in between: calls short enough
    that the search kernel captures by itself (< 1.7 M samples) alternate with calls whose capture is a launch of its own on
    its own stream -- the records must come out complete, in stream order and equal to the one-shot oracle's."""
    torch = pytest.importorskip("torch")
    xs = []
    for i, lead in enumerate((20000, 380000, 300000, 40000, 150000, 250000)):
        c = multi.carrier(0)
        x, _, _ = synth.config2_period(n_total=N1, snr_db=[None, 30.0, 15.0][i % 3], seed=4100 + i, center=c.center_freq,
                                       min10="21255588%02d" % i, lead=lead)
        xs.append(x)
    x = np.concatenate(xs)
    _, d = oracle.rx_chain_f32(x)
    ob = oracle.rx_detect(d)
    assert len(ob) == 6
    sizes = [2200000, 64000, 1900000, 300000, 2400000, 1600, 1800000, 2, 2500000, 120000, 1730000, 500000]
    rx = capi.ReccIq(max_samples=2500000, max_bursts=64)
    t = torch.from_numpy(x.view(np.float32).copy()).cuda()
    st = torch.cuda.current_stream().cuda_stream
    got, pos, k = [], 0, 0
    while pos < len(x):
        n = min(sizes[k % len(sizes)], len(x) - pos)
        rx.submit_dev(t.data_ptr() + 8 * pos, n, st)
        pos += n
        k += 1
        if collect_between and k % 3 == 0:
            got += rx.collect()
    got += rx.collect()
    assert rx.stats()["demod_out"] == len(x) // 50
    check(got, ob, oracle)
    assert [g.demod_index for g in got] == sorted(g.demod_index for g in got)
    rx.close()


@pytest.mark.parametrize("sc16", [False, True])
def test_host_path_uploads_big_calls_in_pieces(capi, oracle, stream, sc16, monkeypatch):
    """amps_recc_iq_work on a call of at least two upload pieces (AMPS_RX_PIECE, 2^24 samples by default, 300 000 here): the
    pieces go up on a copy stream while the kernels of the piece before run -- same stream, same bursts, ragged call lengths."""
    x, d_orc, ob = stream
    scale = 1.0 / 8192.0
    monkeypatch.setenv("AMPS_RX_PIECE", "300000")
    if sc16:
        q = np.clip(np.round(x.view(np.float32) / scale), -32768, 32767).astype(np.int16)
        xq = (q.astype(np.float32) * np.float32(scale)).view(np.complex64)
        _, d_orc = oracle.rx_chain_f32(xq)
        ob = oracle.rx_detect(d_orc)
        rx = capi.ReccIq(max_samples=len(x), sc16=True, sc16_scale=scale)
        cuts = [0, 2500001 * 2, 2500001 * 2 + 4 * 150001, 2 * len(x)]          # (int16 counts: whole I,Q pairs)
        got = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            got += rx.work(q[a:b])
    else:
        rx = capi.ReccIq(max_samples=len(x))
        cuts = [0, 2500001, 2500001 + 599999, 2500001 + 599999 + 17, len(x)]    # big (8 pieces), just under two pieces, tiny, big
        got = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            got += rx.work(x[a:b])
    nd = rx.stats()["demod_out"]
    assert nd == len(x) // 50
    assert bits_equal_f32(rx.read_demod(nd - 60000, 60000), d_orc[nd - 60000:nd])
    check(got, ob, oracle)
    assert rx.stats()["kernel_launches"] > 3 * 8
    rx.close()


def test_device_calls_then_host_calls_without_collect(capi, oracle, stream):
    """submit_dev() calls still in flight on the side streams when the caller switches to amps_recc_iq_work(): the host-path
    call is ordered behind them (whole units only, so that nothing is pending in the device-side carry)."""
    torch = pytest.importorskip("torch")
    x, d_orc, ob = stream
    rx = capi.ReccIq(max_samples=2400000)
    t = torch.from_numpy(x.view(np.float32).copy()).cuda()
    st = torch.cuda.current_stream().cuda_stream
    pos = 0
    for n in (1600 * 1400, 1600 * 40, 1600 * 1100):          # one with its own capture launch, two that capture by themselves
        rx.submit_dev(t.data_ptr() + 8 * pos, n, st)
        pos += n
    got = []
    while pos < len(x):
        n = min(777777, len(x) - pos)
        got += rx.work(x[pos:pos + n])
        pos += n
    got += rx.collect()
    assert rx.stats()["demod_out"] == len(x) // 50
    check(sorted(got, key=lambda g: g.demod_index), ob, oracle)
    rx.close()
