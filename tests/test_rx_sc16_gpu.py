"""sc16 input (AMPS_RX_INPUT_SC16): interleaved int16 I,Q -- the USRP's wire format -- converted in the front kernel.
Bar: bit-identical to feeding the fc32 path (and the fp32 oracle) the converted floats x = (float)int16 * scale."""
import numpy as np
import pytest

from gr_amps_b200 import synth
from tests.helpers import bits_equal_f32, words_equal

pytestmark = pytest.mark.gpu
PASS = 38400
N1 = 55 * PASS


@pytest.fixture(scope="module")
def capi():
    from gr_amps_b200 import capi as c
    c.lib()
    return c


def quantize(x: np.ndarray, scale: float):
    """complex64 -> (int16 I,Q interleaved, the float32 stream those integers stand for)."""
    f = x.view(np.float32)
    s = np.clip(np.rint(f / np.float32(scale)), -32768, 32767).astype(np.int16)
    back = (s.astype(np.float32) * np.float32(scale)).view(np.complex64)
    return s, back


@pytest.mark.parametrize("snr,scale", [(None, 1.0 / 32768.0), (20.0, 1.0 / 32768.0), (15.0, 1.0 / 8192.0), (20.0, 3.0e-5)])
def test_sc16_is_the_fc32_path_on_the_converted_floats(capi, oracle, snr, scale):
    x, hs, _ = synth.config2_period(n_total=N1, snr_db=snr)
    s16, xf = quantize(x, scale)
    rx = capi.ReccIq(max_samples=N1, sc16=True, sc16_scale=scale if scale != 1.0 / 32768.0 else 0.0)
    bursts = rx.work(s16)
    d_gpu = rx.read_demod(0, N1 // 50)
    _, d_orc = oracle.rx_chain_f32(xf)
    assert bits_equal_f32(d_gpu, d_orc)
    ob = oracle.rx_detect(d_orc)
    assert len(bursts) == len(ob) == 1
    assert bursts[0].demod_index == ob[0][0] and np.float32(bursts[0].corr) == np.float32(ob[0][1])
    assert np.array_equal(bursts[0].symbols_np(), ob[0][2])
    assert np.array_equal(bursts[0].symbols_np(), hs[82:82 + 3374])
    assert words_equal(bursts[0].decoded, oracle.recc_decode(ob[0][2])) == []
    # and the float path given those floats agrees bit for bit
    fx = capi.ReccIq(max_samples=N1)
    bf = fx.work(xf)
    assert bits_equal_f32(fx.read_demod(0, N1 // 50), d_gpu)
    assert len(bf) == 1 and bytes(bf[0].symbols) == bytes(bursts[0].symbols)
    rx.close(); fx.close()


def test_sc16_streaming_and_device_path(capi, oracle):
    torch = pytest.importorskip("torch")
    x, _, _ = synth.config2_period(n_total=N1, snr_db=20.0, seed=11)
    x = np.concatenate([x, x, x])
    s16, xf = quantize(x, 1.0 / 32768.0)
    _, d_orc = oracle.rx_chain_f32(xf)
    # arbitrary chunking through the host call (carry of partial passes in the 4-byte format)
    st = capi.ReccIq(max_samples=500000, sc16=True)
    rng = np.random.default_rng(5)
    got, pos = [], 0
    while pos < len(x):
        n = int(rng.integers(1, 500000))
        got += st.work(s16[2 * pos:2 * (pos + n)])
        pos += n
    assert len(got) == 3
    nd = len(x) // 50
    assert bits_equal_f32(st.read_demod(nd - 5000, 5000), d_orc[nd - 5000:nd])
    # device-resident
    dv = capi.ReccIq(max_samples=len(x), sc16=True)
    t = torch.from_numpy(s16).cuda()
    dv.submit_dev(t.data_ptr(), len(x), torch.cuda.current_stream().cuda_stream)
    b = dv.collect()
    assert [r.demod_index for r in b] == [r.demod_index for r in got]
    assert all(bytes(p.symbols) == bytes(q.symbols) for p, q in zip(b, got))
    assert bits_equal_f32(dv.read_demod(0, len(x) // 50), d_orc)
    st.close(); dv.close()


def test_sc16_native_400k_rate(capi, oracle):
    """The same switch on the reference's own 400 kS/s operating point."""
    n = 1536 * 64
    hsb = synth.manchester(synth.recc_message_bits(synth.origination_words()))
    x = synth.fm_burst(hsb, n, 2000, samp_rate=400e3, snr_db=25.0, seed=3)
    s16, xf = quantize(x, 1.0 / 32768.0)
    rx = capi.ReccIq(max_samples=n, samp_rate=400e3, sc16=True)
    b = rx.work(s16)
    _, d_orc = oracle.rx_chain400_f32(xf)
    assert bits_equal_f32(rx.read_demod(0, n // 2), d_orc)
    assert len(b) == 1 and np.array_equal(b[0].symbols_np(), hsb[82:82 + 3374])
    rx.close()


def test_format_mismatch_is_refused(capi):
    a = capi.ReccIq(max_samples=PASS, sc16=True)
    b = capi.ReccIq(max_samples=PASS)
    z16 = np.zeros(2 * PASS, np.int16)
    zf = np.zeros(PASS, np.complex64)
    with pytest.raises(TypeError):
        a.work(zf)
    import ctypes as C
    L = capi.lib()
    assert L.amps_recc_iq_work(a.h, zf.ctypes.data_as(C.c_void_p), PASS, C.cast(None, capi.BURST_CB), None) != 0
    assert L.amps_recc_iq_work_sc16(b.h, z16.ctypes.data_as(C.c_void_p), PASS, C.cast(None, capi.BURST_CB), None) != 0
    a.close(); b.close()
