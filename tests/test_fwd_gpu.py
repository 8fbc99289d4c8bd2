"""GPU parity of the fused forward path (BASELINE config 3: FOCC @0 Hz + FVC @+60 kHz + FVC @+90 kHz,
x0.5, 10 MS/s) against the float64 oracle chain: <= 1e-6 RMS on the complex baseband."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from gr_amps_b200 import capi as c
    c.lib()
    return c


def config3_symbols(oracle, nsym):
    focc = oracle.Focc(100000, False).generate(nsym, chunk=1 << 20)
    alert = oracle.word("orc_fvc_word1_general", 1, 0, 0, 1)
    streams = [focc]
    for _ in range(2):
        v = oracle.Fvc(100000)
        v.push_words(alert)
        out = bytearray()
        while len(out) < nsym:
            r, b, _ = v.work(min(8192, nsym - len(out)))
            out += b.tobytes()
        streams.append(np.frombuffer(bytes(out), np.uint8))
    return streams


def rms(a):
    return float(np.sqrt(np.mean(np.abs(a) ** 2)))


def test_config3_three_carriers_vs_float64(capi, oracle):
    nsym = 41944                                   # 4 194 400 output samples (>= 2^22)
    syms = config3_symbols(oracle, nsym)
    fw = capi.Fwd(max_samples=nsym * 100)
    for c, tw in enumerate((5e3, 3e3, 3e3)):
        assert np.array_equal(fw.taps(c), oracle.firdes_low_pass(1.0, 400e3, 10e3, tw, 0))
    assert len(fw.taps(0)) == 193 and len(fw.taps(1)) == 321
    y = fw.work(syms)
    ref = oracle.fwd_chain_f64(syms)
    assert y.shape == ref.shape == (nsym * 100,)
    err = rms(y.astype(np.complex128) - ref)
    assert err <= 1e-6, err
    # the x4 pfb interpolator has no gain compensation: each carrier comes out at ~0.2; three of them x 0.5
    assert 0.1 < rms(ref) < 0.3
    fw.close()


def test_streaming_chunks_equal_one_shot(capi, oracle):
    nsym = 20000
    syms = config3_symbols(oracle, nsym)
    one = capi.Fwd(max_samples=nsym * 100).work(syms)
    fw = capi.Fwd(max_samples=nsym * 100)
    rng = np.random.default_rng(2)
    parts, pos = [], 0
    while pos < nsym:
        n = int(rng.integers(1, 5000))
        parts.append(fw.work([s[pos:pos + n] for s in syms]))
        pos += n
    got = np.concatenate(parts)
    assert got.shape == one.shape
    assert np.array_equal(got.view(np.float32), one.view(np.float32))     # bit-identical, whatever the chunking
    fw.reset()
    again = fw.work(syms)
    assert np.array_equal(again.view(np.float32), one.view(np.float32))


def test_single_carrier_spectrum_and_mute(capi, oracle):
    nsym = 8192
    focc = oracle.Focc(100000, False).generate(nsym, chunk=1 << 20)
    fw = capi.Fwd(max_samples=nsym * 100, carrier_freq=(0.0,), lpf_transition=(5e3,), out_scale=1.0)
    y = fw.work([focc])
    ref = oracle.fwd_chain_f64([focc], carrier_freq=(0.0,), lpf_transition=(5e3,), scale=1.0)
    assert rms(y.astype(np.complex128) - ref) <= 1e-6
    steady = y[60000:]
    assert 0.15 < rms(steady) < 0.27                # ~1/4 (uncompensated x4 interpolator) x in-band fraction of the FSK spectrum
    # instantaneous frequency stays within the +-8 kHz deviation (plus filter ringing)
    f = np.angle(steady[1:] * np.conj(steady[:-1])) * 10e6 / (2 * np.pi)
    assert np.max(np.abs(f)) < 12e3
    # a muted stream (bytes 0) produces silence once the filters have drained
    fw.reset()
    z = fw.work([np.zeros(nsym, np.uint8)])
    assert np.all(z == 0)


def test_device_resident_submit(capi, oracle):
    torch = pytest.importorskip("torch")
    nsym = 63 * 400 + 17
    syms = config3_symbols(oracle, nsym)
    dsym = [torch.from_numpy(s.copy()).cuda() for s in syms]
    out = torch.empty(2 * nsym * 100, dtype=torch.float32, device="cuda")
    fw = capi.Fwd(max_samples=nsym * 100)
    fw.submit_dev([d.data_ptr() for d in dsym], nsym, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = oracle.fwd_chain_f64(syms)
    assert rms(out.cpu().numpy().view(np.complex64).astype(np.complex128) - ref) <= 1e-6


# ------------------------------------------------------------------ Manchester-bit fast path
def bits_of(sym_bytes, sps=5):
    """Half-symbol bytes (+1 = 0x01, -1 = 0xFF), 2*sps per bit -> data bits (bit 1 = (low, high))."""
    hs = np.asarray(sym_bytes).reshape(-1, 2 * sps)
    return (hs[:, sps] == 1).astype(np.uint8)


def test_bit_fast_path_matches_float64_chain(capi, oracle):
    nbits = 4200
    syms = config3_symbols(oracle, nbits * 10)
    bits = [bits_of(s) for s in syms]
    bits[2][1000:1400] = 0xFF                                    # a muted stretch on the third carrier
    syms[2] = syms[2].copy()
    syms[2][10000:14000] = 0
    fw = capi.Fwd(max_samples=nbits * 1000)
    y = fw.work_bits(bits)
    ref = oracle.fwd_chain_f64(syms)
    assert y.shape == ref.shape
    assert rms(y.astype(np.complex128) - ref) <= 1e-6
    # and it agrees with the general half-symbol path to fp32 rounding
    g = capi.Fwd(max_samples=nbits * 1000).work(syms)
    assert rms(y.astype(np.complex128) - g.astype(np.complex128)) <= 1e-6
    with pytest.raises(capi.AmpsError):
        fw.work(syms)                                            # no mixing of input kinds without reset()
    fw.reset()
    assert np.array_equal(fw.work(syms).view(np.float32), g.view(np.float32))


def test_bit_fast_path_streaming(capi, oracle):
    nbits = 3000
    syms = config3_symbols(oracle, nbits * 10)
    bits = [bits_of(s) for s in syms]
    one = capi.Fwd(max_samples=nbits * 1000).work_bits(bits)
    fw = capi.Fwd(max_samples=nbits * 1000)
    rng = np.random.default_rng(9)
    parts, pos = [], 0
    while pos < nbits:
        n = int(rng.integers(1, 400))
        parts.append(fw.work_bits([b[pos:pos + n] for b in bits]))
        pos += n
    got = np.concatenate(parts)
    assert np.array_equal(got.view(np.float32), one.view(np.float32))


@pytest.mark.parametrize("ncar", [1, 2])
def test_fewer_carriers_both_input_kinds(capi, oracle, ncar):
    """One and two carriers (the polyphase taps are then shared by three / six warps, the bit path's ballots run on idle
    warps too): half-symbol input and data-bit input against the float64 chain, and against each other."""
    nbits = 2100
    syms = config3_symbols(oracle, nbits * 10)[:ncar]
    cf, tw = (0.0, 60e3)[:ncar], (5e3, 3e3)[:ncar]
    ref = oracle.fwd_chain_f64(syms, carrier_freq=cf, lpf_transition=tw, scale=0.5)
    g = capi.Fwd(max_samples=nbits * 1000, carrier_freq=cf, lpf_transition=tw).work(syms)
    assert rms(g.astype(np.complex128) - ref) <= 1e-6
    b = capi.Fwd(max_samples=nbits * 1000, carrier_freq=cf, lpf_transition=tw).work_bits([bits_of(s) for s in syms])
    assert rms(b.astype(np.complex128) - ref) <= 1e-6
    assert rms(b.astype(np.complex128) - g.astype(np.complex128)) <= 1e-6
