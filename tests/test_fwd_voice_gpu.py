"""GPU parity of the voice legs of the forward path: audio @16 kS/s + SAT -> nbfm_tx ->
[mute] -> x25 arb resampler, added to the +60 kHz carrier (with the FVC data) and alone on the +90 kHz carrier
(grc/ampsbs.grc:715-773, 943-1005, 1994-2119, 4494-4500, 4632-4638) against the float64 oracle: <= 1e-6 RMS."""
import numpy as np
import pytest

from tests.test_fwd_gpu import config3_symbols, rms

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from gr_amps_b200 import capi as c
    c.lib()
    return c


def speech_like(n, seed=3):
    """Band-limited test audio @16 kS/s, peak ~0.5 (stands in for boot16k.wav, which the reference does not ship)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    x = sum(a * np.sin(2 * np.pi * f * t + p) for a, f, p in
            zip(rng.uniform(0.03, 0.12, 8), rng.uniform(300, 3000, 8), rng.uniform(0, 6.28, 8)))
    return (x * (0.6 + 0.4 * np.sin(2 * np.pi * 3.1 * t))).astype(np.float32)


def voice_symbols(oracle, nsym):
    s = config3_symbols(oracle, nsym)
    return [s[0], s[1], np.zeros(nsym, np.uint8)]          # +90 kHz carries audio only in the reference graph


def test_voice_legs_vs_float64(capi, oracle):
    nsym = 41950                                            # 4 195 000 output samples (>= 2^22), 6712 audio samples
    syms = voice_symbols(oracle, nsym)
    audio = speech_like(nsym * 4 // 25)
    fw = capi.Fwd(max_samples=nsym * 100)
    fw.enable_voice()
    y = fw.work_voice(syms, audio)
    v = oracle.voice_tx_f64(audio)
    ref = oracle.fwd_chain_voice_f64(syms, [None, v, v])
    err = rms(y.astype(np.complex128) - ref)
    assert err <= 1e-6, err
    # the legs are really there: without them the result differs by the two voice carriers (0.375 each, x0.5)
    plain = oracle.fwd_chain_f64(syms)
    assert 0.2 < rms(ref - plain) < 0.35
    # the resampler (gain 3 over 8 arms) passes the FM carrier at ~3/8, minus the sidebands its 4.8 kHz cut-off removes
    assert 0.3 < rms(v[25 * 200:]) < 0.4
    fw.close()


def test_voice_streaming_and_mute(capi, oracle):
    nsym = 20000
    syms = voice_symbols(oracle, nsym)
    audio = speech_like(nsym * 4 // 25, seed=9)
    one = capi.Fwd(max_samples=nsym * 100)
    one.enable_voice()
    y1 = one.work_voice(syms, audio)
    # arbitrary chunking (whole audio samples) is bit-identical; audio_mute on some calls follows the oracle's per-sample mute
    fw = capi.Fwd(max_samples=nsym * 100)
    fw.enable_voice()
    rng = np.random.default_rng(5)
    parts, pos = [], 0
    while pos < nsym:
        n = 25 * int(rng.integers(1, 200))
        n = min(n, nsym - pos)
        parts.append(fw.work_voice([s[pos:pos + n] for s in syms], audio[pos * 4 // 25:(pos + n) * 4 // 25]))
        pos += n
    assert np.array_equal(np.concatenate(parts).view(np.float32), y1.view(np.float32))
    fw.reset()
    mute = np.zeros(len(audio), np.uint8)
    parts, pos, k = [], 0, 0
    while pos < nsym:
        n = min(2500, nsym - pos)
        m = (k % 3) == 1
        mute[pos * 4 // 25:(pos + n) * 4 // 25] = m
        parts.append(fw.work_voice([s[pos:pos + n] for s in syms], audio[pos * 4 // 25:(pos + n) * 4 // 25], audio_mute=m))
        pos += n; k += 1
    got = np.concatenate(parts)
    ref = oracle.fwd_chain_voice_f64(syms, [None, oracle.voice_tx_f64(audio, mute=mute), oracle.voice_tx_f64(audio)])
    assert rms(got.astype(np.complex128) - ref) <= 1e-6
    assert rms(got - y1) > 0.01                              # muting changed the +60 kHz leg
    one.close(); fw.close()


def test_voice_api_errors(capi, oracle):
    fw = capi.Fwd(max_samples=100000)
    syms = voice_symbols(oracle, 1000)
    with pytest.raises(capi.AmpsError):
        fw.work_voice(syms, np.zeros(160, np.float32))       # voice not enabled
    fw.enable_voice()
    with pytest.raises(capi.AmpsError):
        fw.enable_voice()                                    # twice
    with pytest.raises(capi.AmpsError):
        fw.work(syms)                                        # plain entry point on a voice handle
    with pytest.raises(capi.AmpsError):
        fw.work_voice([s[:990] for s in syms], np.zeros(990 * 4 // 25, np.float32))    # not a multiple of 25
    y = fw.work_voice(syms, np.zeros(160, np.float32))
    assert y.shape == (100000,) and np.isfinite(y.view(np.float32)).all()
    fw.close()
    two = capi.Fwd(max_samples=100000, carrier_freq=(0.0, 60e3), lpf_transition=(5e3, 3e3))
    with pytest.raises(capi.AmpsError):
        two.enable_voice(carrier_gated=1, carrier_open=2)    # no carrier 2
    two.enable_voice(carrier_gated=1, carrier_open=-1)
    two.close()
